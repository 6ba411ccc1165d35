"""Stand-ins for the caller-side callables of the jellyfish sampler, shared by tests/golden/make_golden_jellyfish_sampler.py
(run against the unmodified reference) and tests/test_jellyfish_sampler.py: a deterministic `bd_updater(bd [N,3,H,W],
dtheta [N]) -> [N,3,H,W]` and a `design_fn(x [B,F,4,H,W], bd_0_expand) -> dJ/dx` evaluated with autograd, in the shape of
inference/inference_2d_jellyfish.py:85-114, :276-279 (the real ones are the surrogate nets of SURVEY.md 8(f) rank 1)."""
import torch


class BdUpdater(torch.nn.Module):
    def forward(self, bd, dtheta):
        d = dtheta.reshape(-1, 1, 1, 1).to(bd.dtype)
        shifted = torch.roll(bd, shifts=1, dims=-1)
        return torch.tanh(bd * (1 + d) + 0.25 * d * shifted)


bd_updater = BdUpdater()


def design_fn(x, bd_0_expand):
    state, theta_expand = x[:, :, :3], x[:, :, 3]
    theta = theta_expand.mean((-1, -2))
    w = torch.arange(theta.shape[1], 0, -1, dtype=x.dtype, device=x.device)
    force = (state[:, :, 2] * bd_0_expand[:, :, 0]).mean((-1, -2)) * torch.sin(3 * theta)
    J = -(force * w).mean(1) + 10.0 * (theta[:, 1:] - theta[:, :-1]).square().sum(1) + 0.1 * state[:, :, :2].square().mean((1, 2, 3, 4))
    return torch.autograd.grad(J, x, grad_outputs=torch.ones_like(J))[0]
