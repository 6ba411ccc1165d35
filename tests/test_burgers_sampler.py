"""Burgers two-model guided DDPM sampler against the trace of the UNMODIFIED reference
diffusion/diffusion_1d_burgers.py (tests/golden/make_golden_burgers_sampler.py).

Teacher-forced per step: each p_sample gets the reference's own input and noise, so the comparison isolates one step.
The step divides the eps error by sqrt(abar_t): the bound below is  tol_eps * (c1 * sqrt(1/abar - 1) + 1)  with the
U-Net tolerance of the precision mode (3xTF32 2e-4 of the eps scale).  The whole loop is asserted in 3xTF32 only."""
import os

import numpy as np
import pytest
import torch

from diffphycon_b200 import diffusion_1d_burgers as dd
from diffphycon_b200.burgers_unet import Unet2D
from oracle import param_gen

KW_UW = dict(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=1)
KW_W = dict(dim=32, dim_mults=(1, 2, 4), channels=2, resnet_block_groups=1)
T = 6


def build(kw, seed, precision):
    net = Unet2D(**kw)
    net.load_state_dict(param_gen.make_params({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed), strict=True)
    net.precision = precision
    return net


VARIANTS = {   # golden name -> (models, guidance_u0, diffusion kwargs), as in make_golden_burgers_sampler.py
    "burgers_sampler": ("both", True, dict(eval_two_models=True, prior_beta=1.5)),
    "burgers_sampler_single_ut": ("uw", False, {}),
    "burgers_sampler_model_w": ("w", True, dict(is_model_w=True, prior_beta=0.7)),
    # self recurrence (burgers.py:472-482, :535-578): two passes per diffusion step, each followed by recurrent_sample
    "burgers_sampler_recurrent": ("both", True, dict(eval_two_models=True, prior_beta=1.5, recurrence=True, recurrence_k=2)),
}


def make(z, device, precision="3xtf32", variant="burgers_sampler"):
    which, _, dkw = VARIANTS[variant]
    uw, w = build(KW_UW, 31, precision), build(KW_W, 32, precision)
    d = dd.GaussianDiffusion({"both": (uw, w), "uw": uw, "w": w}[which], seq_length=(16, 128), timesteps=T,
                             auto_normalize=False, use_conv2d=True, temporal=True, is_condition_u0=True,
                             is_condition_uT=True, **dkw).to(device)
    target = torch.from_numpy(z["target"]).to(device)

    def loss_fn(x):
        return (x[:, 0, 10, :] - target).square().mean(-1) + 0.05 * x[:, 1, :10, :].square().mean((-1, -2))

    kw = dict(nablaJ=dd.get_nablaJ(loss_fn), J_scheduler=lambda t: 0.5 * dd.cosine_beta_J_schedule(t),
              w_scheduler=dd.sigmoid_schedule_flip, u_init=torch.from_numpy(z["u_init"]).to(device),
              u_final=torch.from_numpy(z["u_final"]).to(device))
    return d, kw


def check_steps(z, device, precision, tol_eps):
    d, kw = make(z, device, precision)
    d.guidance_u0 = True
    zi = 0
    for t in reversed(range(T)):
        x = torch.from_numpy(z[f"x{t}"]).to(device)
        if t > 0:
            n = torch.from_numpy(z[f"z{zi}"]).to(device)
            zi += 1
            d.sample_noise = lambda shape, dev, n=n: n
        pred, xs, _ = d.p_sample(x, t, clip_denoised=True, **kw)
        s = d._sched()
        amp = float(s['posterior_mean_coef1'][t]) * float(s['sqrt_recipm1_alphas_cumprod'][t]) + 1.0
        ref, rxs = torch.from_numpy(z[f"pred{t}"]), torch.from_numpy(z[f"xstart{t}"])
        scale = max(1.0, ref.abs().max().item())
        assert (pred.cpu() - ref).abs().max().item() <= tol_eps * amp * scale, (t, (pred.cpu() - ref).abs().max().item())
        assert (xs.cpu() - rxs).abs().max().item() <= tol_eps * (float(s['sqrt_recipm1_alphas_cumprod'][t]) + 1.0), t


def check_loop(z, device, variant="burgers_sampler"):
    d, kw = make(z, device, variant=variant)
    nz = len([k for k in z.files if k.startswith("z")])
    noises = [torch.from_numpy(z[f"x{T - 1}"]).to(device)] + [torch.from_numpy(z[f"z{i}"]).to(device) for i in range(nz)]
    # x{T-1} is the initial noise with the conditions already written; set_condition rewrites the same rows
    it = iter(noises)
    d.sample_noise = lambda shape, dev: next(it).clone()
    y = d.sample(batch_size=2, clip_denoised=True, guidance_u0=VARIANTS[variant][1], **kw)
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def _emulate(monkeypatch):
    from diffphycon_b200 import unet3d
    import diffphycon_b200.burgers_unet as bu
    from tests import cpu_emulator
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(unet3d, "_require_cuda", lambda x: None)
    monkeypatch.setattr(bu, "_require_cuda", lambda x: None)


def test_burgers_sampler_host_logic_steps(golden_dir, monkeypatch):
    _emulate(monkeypatch)
    check_steps(np.load(os.path.join(golden_dir, "burgers_sampler.npz")), "cpu", "3xtf32", 2e-4)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_burgers_sampler_host_logic_loop(variant, golden_dir, monkeypatch):
    _emulate(monkeypatch)
    check_loop(np.load(os.path.join(golden_dir, variant + ".npz")), "cpu", variant)


def test_burgers_sampler_accepts_reference_style_nablaJ(golden_dir, monkeypatch):
    """A nablaJ built the reference's way (requires_grad_ + autograd.grad, no enable_grad) must work: the loop is not
    under no_grad (burgers.py:525 has no decorator)."""
    _emulate(monkeypatch)
    z = np.load(os.path.join(golden_dir, "burgers_sampler.npz"))
    d, kw = make(z, "cpu")

    def ref_style(x):
        x.requires_grad_(True)
        J = x.square().mean((-1, -2, -3))
        return torch.autograd.grad(J, x, grad_outputs=torch.ones_like(J))[0]

    kw["nablaJ"] = ref_style
    x = torch.from_numpy(z[f"x{T - 1}"])
    pred, _, _ = d.p_sample(x, T - 1, clip_denoised=True, **kw)
    assert torch.isfinite(pred).all()


def test_unimplemented_options_raise():
    net = Unet2D(**KW_UW)
    with pytest.raises(NotImplementedError):
        dd.GaussianDiffusion(net, seq_length=(16, 128), temporal=False, use_conv2d=False)
    d = dd.GaussianDiffusion(net, seq_length=(16, 128), temporal=True, use_conv2d=True, timesteps=10, sampling_timesteps=5)
    with pytest.raises(NotImplementedError):
        d.sample(batch_size=1)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-4), ("tf32", 1e-2)])
def test_burgers_sampler_steps_gpu(precision, tol, golden_dir):
    check_steps(np.load(os.path.join(golden_dir, "burgers_sampler.npz")), "cuda", precision, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_burgers_sampler_loop_gpu(variant, golden_dir):
    check_loop(np.load(os.path.join(golden_dir, variant + ".npz")), "cuda", variant)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_burgers_cuda_graph_networks_equal_eager(variant, golden_dir):
    """use_cuda_graph replays the captured network forwards (fixed state / time buffers) instead of launching ~270 kernels per
    step: same kernels and arithmetic, so the sampled trajectories are bit-identical to the eager loop, and one graph launch per
    network evaluation is counted."""
    from diffphycon_b200 import _lib
    z = np.load(os.path.join(golden_dir, variant + ".npz"))
    nz = len([k for k in z.files if k.startswith("z")])
    out = []
    for graph in (False, True):
        d, kw = make(z, "cuda", precision="tf32", variant=variant)
        d.use_cuda_graph = graph
        noises = [torch.from_numpy(z[f"x{T - 1}"]).cuda()] + [torch.from_numpy(z[f"z{i}"]).cuda() for i in range(nz)]
        it = iter(noises)
        d.sample_noise = lambda shape, dev: next(it).clone()
        n0 = _lib.LaunchCounter.graph_launches
        out.append(d.sample(batch_size=2, clip_denoised=True, guidance_u0=VARIANTS[variant][1], **kw))
        if graph:
            assert _lib.LaunchCounter.graph_launches - n0 == T * VARIANTS[variant][2].get("recurrence_k", 1)
    assert torch.equal(out[0], out[1]), (out[0] - out[1]).abs().max().item()
