"""Golden traces of the jellyfish sampler from the UNMODIFIED reference
diffusion/diffusion_2d_jellyfish.py with reference Unet3D_with_Conv3D nets (build container only).  `bd_updater` and
`design_fn` are the caller's callables in the reference too; the stand-ins below are defined identically in
tests/test_jellyfish_sampler.py.  Recorded per sampling step: the state fed to the models, the noise drawn, the state after
the step (pred_states / pred_theta) and, for the loop, the final result."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import param_gen, ref_import  # noqa: E402
from tests.jellyfish_standins import bd_updater, design_fn  # noqa: E402

ref_import._prepare()
import importlib  # noqa: E402

jm = importlib.import_module("diffusion.diffusion_2d_jellyfish")
um = ref_import.unet3d_module()
B, FR, S = 2, 4, 16


def build(out_dim, seed):
    net = um.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=7, out_dim=out_dim)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items() if not k.endswith("rotary_emb.freqs")}
    net.load_state_dict(param_gen.make_params(shapes, seed), strict=False)
    return net.eval()


def run(name, guidance, sampling_timesteps=None, eta=0.0, T=5, design=True, B=B, FR=FR, S=S, store_trace=True, cond_steps=1):
    mj, mw = build(4, 41), build(1, 42)
    d = jm.GaussianDiffusion([mj, mw], image_size=S, frames=FR, cond_steps=cond_steps, timesteps=T, sampling_timesteps=sampling_timesteps,
                             loss_type='l2', objective='pred_noise', standard_fixed_ratio=0.05, coeff_ratio_J=0.3,
                             coeff_ratio_w=0.4, eval_2ddpm=True, w_prob_exp=0.7, ddim_sampling_eta=eta, device='cpu')
    g = torch.Generator().manual_seed(9)
    state_0 = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    bd_0 = torch.cat([(torch.rand(B, 1, S, S, generator=g) > 0.7).float(), torch.rand(B, 2, S, S, generator=g) - 0.5], 1)
    thetas_0 = torch.rand(B, generator=g) * 0.7 + 0.2
    trace, draws = {}, []

    def rec_noise(shape, device):
        n = torch.randn(shape)
        draws.append(n.numpy().copy())
        return n

    d.sample_noise = rec_noise
    orig_rl = torch.randn_like

    def rec_rl(t_, **k):
        n = orig_rl(t_, **k)
        draws.append(n.numpy().copy())
        return n

    orig_mp = d.model_predictions

    def rec_mp(x, t, *a, **k):
        trace[f"x{len([q for q in trace if q.startswith('x')])}"] = x.detach().clone().numpy()
        return orig_mp(x, t, *a, **k)

    d.model_predictions = rec_mp
    torch.randn_like = rec_rl
    torch.manual_seed(3)
    try:
        states, theta = d.sample(design_fn=design_fn if design else None, design_guidance=guidance, cond=[state_0, bd_0],
                                 thetas_0=thetas_0, bd_updater=bd_updater)
    finally:
        torch.randn_like = orig_rl
    for i, n in enumerate(draws):
        trace[f"z{i}"] = n
    if not store_trace:      # the test regenerates the noise from the CPU generator (torch.manual_seed(3), same draw order)
        trace = {}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), state_0=state_0.numpy(), bd_0=bd_0.numpy(), thetas_0=thetas_0.numpy(),
                        states=states.detach().numpy(), theta=theta.detach().numpy(), **trace)
    print(name, sorted(trace), tuple(states.shape), tuple(theta.shape), float(states.abs().mean()))


if len(sys.argv) > 1 and sys.argv[1] == "repaint":
    # cond_steps == 0: the unconditional model with repaint conditioning (jf.py:865-873); the draws include the q_sample noise
    run("jelly_ddpm_repaint", "standard-alpha", cond_steps=0)
    sys.exit(0)
run("jelly_ddpm_alpha", "standard-alpha")
run("jelly_ddpm_standard", "standard")
run("jelly_ddpm_noguide", "standard", design=False)
# the reference's DDIM path hard-codes [B, 20, 4, 64, 64] (jellyfish.py:726): full-size frames, one sample, two effective steps
run("jelly_ddim_alpha", "standard-alpha", sampling_timesteps=3, eta=1.0, T=6, B=1, FR=20, S=64, store_trace=False)
run("jelly_ddim_standard", "standard", sampling_timesteps=3, eta=0.0, T=6, B=1, FR=20, S=64, store_trace=False)
