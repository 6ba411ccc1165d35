"""GPU parity of the sampling loops against the reference's recorded loops (tests/golden/sampler_loop_*.npz).
Noise is drawn on the CPU with the golden run's seed and moved to the device, so both sides consume the identical
noise stream; the networks run in the fp32-class "3xtf32" mode (see tests/test_host_logic.py for why whole-loop parity
is only meaningful there).  Tolerance: 2e-3 of the output scale."""
import os

import numpy as np
import pytest
import torch

import diffphycon_b200 as dpc
from oracle import unet3d_oracle as uo

pytestmark = pytest.mark.gpu


def sampler(precision="3xtf32", **kw):
    cj = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6)
    cw = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2)
    mj = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=2)
    mj.load_state_dict(uo.make_params(cj, 11))
    mw.load_state_dict(uo.make_params(cw, 12))
    mj.precision = mw.precision = precision
    return dpc.GaussianDiffusion([mj, mw], image_size=16, frames=4, eval_2ddpm=True, **kw).cuda()


def cpu_noise(monkeypatch):
    real = torch.randn
    monkeypatch.setattr(torch, "randn", lambda *a, device=None, **k: real(*a, **k).to(device) if device is not None else real(*a, **k))
    monkeypatch.setattr(torch, "randn_like", lambda t, **k: real(t.shape).to(t.device))


def test_ddpm_loop_matches_reference(golden_dir, monkeypatch):
    z = np.load(os.path.join(golden_dir, "sampler_loop_ddpm4.npz"))
    d = sampler(timesteps=4, sampling_timesteps=4, standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
    cpu_noise(monkeypatch)
    torch.manual_seed(42)
    y = d.sample(batch_size=2, design_fn=dpc.StockSmokeGuidance(), design_guidance="standard",
                 init=torch.from_numpy(z["init"])).cpu()
    ref = torch.from_numpy(z["y"])
    assert (y - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_ddim_steps_match_reference_trace(golden_dir):
    """Teacher-forced DDIM steps on the reference's recorded per-step states and noise (1000 -> 3 steps, eta = 1).
    A DDIM step maps an eps perturbation d to  d * (sqrt(1/abar_t - 1) * sqrt(abar_next) + c): up to ~6e2 on this
    schedule, so the bound is that amplification times the fp32-class U-Net tolerance (1e-4), plus 1e-5."""
    from oracle import smoke_sampler_oracle as so
    z = np.load(os.path.join(golden_dir, "sampler_loop_ddim3.npz"))
    d = sampler(timesteps=1000, sampling_timesteps=3, ddim_sampling_eta=1.0, standard_fixed_ratio=1e5, coeff_ratio=0.0,
                w_prob_exp=0.97)
    sch = d._sched()
    init = torch.from_numpy(z["init"]).cuda()
    pairs = so.ddim_times(1000, 3)
    for i, (time, time_next) in enumerate(pairs):
        x = torch.from_numpy(z[f"x{i}"]).cuda()
        last = time_next < 0
        noise = None if last else torch.from_numpy(z[f"z{i}"]).cuda()
        out = d.ddim_step(x, time, time_next, design_fn=dpc.StockSmokeGuidance(), design_guidance="standard", init=init,
                          noise=noise).cpu()
        ref = torch.from_numpy(z["y"] if last else z[f"x{i + 1}"])
        srm1 = float(sch["sqrt_recipm1_alphas_cumprod"][time])
        amp = srm1 if last else srm1 * float(sch["alphas_cumprod"][time_next].sqrt()) + 1.0
        err = (out - ref).abs().max().item()
        assert err <= 1e-4 * amp + 1e-5, (i, err, amp)


def test_sampling_properties_tf32_mode():
    """Size-independent properties of the default (TF32) path: the initial condition is re-imposed exactly, every value is
    finite, fixed seeds reproduce bit-identical trajectories, and samples do not depend on their batch neighbours."""
    d = sampler(precision="tf32", timesteps=1000, sampling_timesteps=5, ddim_sampling_eta=1.0, standard_fixed_ratio=1e5,
                coeff_ratio=0.0, w_prob_exp=0.97)
    g = torch.Generator().manual_seed(1)
    init = torch.rand(3, 16, 16, generator=g).cuda() / 2
    torch.manual_seed(7)
    y = d.sample(batch_size=3, design_fn=dpc.StockSmokeGuidance(), init=init)
    torch.manual_seed(7)
    y2 = d.sample(batch_size=3, design_fn=dpc.StockSmokeGuidance(), init=init)
    assert torch.isfinite(y).all() and torch.equal(y, y2)
    assert y.shape == (3, 4, 6, 16, 16)
    assert y.abs().max().item() <= 1.0 + 1e-6  # last DDIM step returns the clipped x_start
    dd = sampler(precision="tf32", timesteps=6, sampling_timesteps=6, standard_fixed_ratio=1e5, coeff_ratio=0.0,
                 w_prob_exp=0.97)
    torch.manual_seed(8)
    yp = dd.sample(batch_size=3, design_fn=dpc.StockSmokeGuidance(), init=init)
    assert torch.equal(yp[:, 0, 0], init) and torch.isfinite(yp).all()


@pytest.mark.parametrize("two_streams", [False, True])
@pytest.mark.parametrize("mode", ["ddpm", "ddim"])
def test_cuda_graph_loop_equals_eager_loop(mode, two_streams):
    """Step orchestration: the loop replayed from ONE captured CUDA graph (device-side step counter, time /
    coefficient tables, optionally the two U-Nets on two streams) reproduces the eager loop bit for bit on the same seed —
    same kernels, same arithmetic, same torch.randn stream."""
    kw = dict(timesteps=6, sampling_timesteps=6) if mode == "ddpm" else dict(timesteps=1000, sampling_timesteps=5, ddim_sampling_eta=1.0)
    d = sampler(precision="tf32", standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97, **kw)
    g = torch.Generator().manual_seed(1)
    init = torch.rand(3, 16, 16, generator=g).cuda() / 2
    fn = dpc.StockSmokeGuidance(w_energy=0.2)
    torch.manual_seed(7)
    ref = d.sample(batch_size=3, design_fn=fn, init=init)
    d.use_cuda_graph, d.two_streams = True, two_streams
    from diffphycon_b200 import _lib
    n0 = _lib.LaunchCounter.graph_launches
    torch.manual_seed(7)
    y = d.sample(batch_size=3, design_fn=fn, init=init)
    assert _lib.LaunchCounter.graph_launches - n0 == (6 if mode == "ddpm" else 5)
    assert torch.equal(y, ref), (y - ref).abs().max().item()
    torch.manual_seed(9)                      # the cached graph serves the next call (other seed, other init)
    init2 = init.flip(0).contiguous()
    y2 = d.sample(batch_size=3, design_fn=fn, init=init2)
    d.use_cuda_graph = False
    torch.manual_seed(9)
    ref2 = d.sample(batch_size=3, design_fn=fn, init=init2)
    assert torch.equal(y2, ref2)
