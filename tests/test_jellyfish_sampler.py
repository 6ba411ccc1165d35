"""Jellyfish sampler against traces of the UNMODIFIED reference diffusion/diffusion_2d_jellyfish.py
(tests/golden/make_golden_jellyfish_sampler.py), with the caller-side stand-ins of tests/jellyfish_standins.py.

DDPM: teacher-forced per step (the reference's own state and noise go in), so one step is isolated; the bound is the U-Net
tolerance of the precision mode times the step's amplification  c1 * sqrt(1/abar - 1) + 1  (+ the guidance's sensitivity,
folded into the factor 4).  DDIM: the reference hard-codes [B,20,4,64,64] there, so the golden is one full-size sample with
two effective steps, noise regenerated from the same CPU generator seed; compared relative to the output scale (x_start is
not clamped on that path and reaches O(1e3))."""
import os

import numpy as np
import pytest
import torch

from diffphycon_b200 import diffusion_2d_jellyfish as dj
from diffphycon_b200.unet3d import Unet3D_with_Conv3D
from oracle import param_gen
from tests.jellyfish_standins import bd_updater, design_fn

DDPM = {"jelly_ddpm_alpha": ("standard-alpha", True), "jelly_ddpm_standard": ("standard", True),
        "jelly_ddpm_noguide": ("standard", False)}
DDIM = {"jelly_ddim_alpha": ("standard-alpha", 1.0), "jelly_ddim_standard": ("standard", 0.0)}


def build(out_dim, seed, precision):
    net = Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=7, out_dim=out_dim)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items() if not k.endswith("rotary_emb.freqs")}
    net.load_state_dict(param_gen.make_params(shapes, seed), strict=False)     # RoPE frequencies keep their defined values
    net.precision = precision
    return net


def make(device, precision, S, FR, T, sampling_timesteps=None, eta=0.0, cond_steps=1):
    d = dj.GaussianDiffusion([build(4, 41, precision), build(1, 42, precision)], image_size=S, frames=FR, cond_steps=cond_steps,
                             timesteps=T, sampling_timesteps=sampling_timesteps, loss_type='l2', objective='pred_noise',
                             standard_fixed_ratio=0.05, coeff_ratio_J=0.3, coeff_ratio_w=0.4, eval_2ddpm=True, w_prob_exp=0.7,
                             ddim_sampling_eta=eta, device='cpu')
    return d.to(device)


def check_ddpm_steps(z, name, device, precision, tol):
    guidance, design = DDPM[name]
    T, S, FR = 5, 16, 4
    d = make(device, precision, S, FR, T)
    dev = lambda a: torch.from_numpy(a).to(device)
    init = iter([dev(z["z0"]), dev(z["z1"]), dev(z["z2"])])
    d.sample_noise = lambda shape, dv: next(init)
    st = d._begin((2, FR, 3, S, S), [dev(z["state_0"]), dev(z["bd_0"])], dev(z["thetas_0"]), bd_updater)
    assert torch.equal(st.x.cpu(), torch.from_numpy(z["x0"]))
    s = d._sched()
    for i, t in enumerate(reversed(range(T))):
        x = dev(z[f"x{i}"])
        st.x.copy_(x)
        st.x_w[:, :, 3:] = x[:, :, 3:]
        if t > 0:
            n = dev(z[f"z{3 + i}"])
            d.sample_noise = lambda shape, dv, n=n: n
        d._ddpm_step(st, t, design_fn if design else None, guidance)
        amp = 4.0 * (float(s['posterior_mean_coef1'][t]) * float(s['sqrt_recipm1_alphas_cumprod'][t]) + 1.0)
        if i + 1 < T:
            ref = torch.from_numpy(z[f"x{i + 1}"])
            err = (st.x.cpu() - ref).abs().max().item()
            assert err <= tol * amp * max(1.0, ref.abs().max().item()), (t, err)
        else:
            ref_s, ref_t = torch.from_numpy(z["states"]), torch.from_numpy(z["theta"])
            assert (st.x[:, :, :3].cpu() - ref_s).abs().max().item() <= tol * amp * max(1.0, ref_s.abs().max().item())
            assert (st.theta_mean.cpu() - ref_t).abs().max().item() <= tol * amp


def check_ddpm_loop(z, name, device, cond_steps=1):
    guidance, design = DDPM.get(name, ("standard-alpha", True))
    d = make(device, "3xtf32", 16, 4, 5, cond_steps=cond_steps)
    nz = len([k for k in z.files if k.startswith("z")])
    it = iter([torch.from_numpy(z[f"z{i}"]).to(device) for i in range(nz)])
    d.sample_noise = lambda shape, dv: next(it)
    dev = lambda a: torch.from_numpy(a).to(device)
    states, theta = d.sample(design_fn=design_fn if design else None, design_guidance=guidance,
                             cond=[dev(z["state_0"]), dev(z["bd_0"])], thetas_0=dev(z["thetas_0"]), bd_updater=bd_updater)
    ref_s, ref_t = torch.from_numpy(z["states"]), torch.from_numpy(z["theta"])
    assert states.shape == ref_s.shape and theta.shape == ref_t.shape
    assert (states.cpu() - ref_s).abs().max().item() <= 5e-3 * max(1.0, ref_s.abs().max().item())
    assert (theta.cpu() - ref_t).abs().max().item() <= 5e-3


def check_ddim(z, name, device, precision, tol):
    guidance, eta = DDIM[name]
    d = make(device, precision, 64, 20, 6, sampling_timesteps=3, eta=eta)
    torch.manual_seed(3)
    d.sample_noise = lambda shape, dv: torch.randn(list(shape)).to(device)     # the reference run's CPU noise stream
    dev = lambda a: torch.from_numpy(a).to(device)
    states, theta = d.sample(design_fn=design_fn, design_guidance=guidance, cond=[dev(z["state_0"]), dev(z["bd_0"])],
                             thetas_0=dev(z["thetas_0"]), bd_updater=bd_updater)
    ref_s, ref_t = torch.from_numpy(z["states"]), torch.from_numpy(z["theta"])
    scale = ref_s.abs().max().item()
    assert (states.cpu() - ref_s).abs().max().item() <= tol * scale, (states.cpu() - ref_s).abs().max().item() / scale
    assert (theta.cpu() - ref_t).abs().max().item() <= tol * max(1.0, ref_t.abs().max().item())


def _emulate(monkeypatch):
    from diffphycon_b200 import unet3d
    from tests import cpu_emulator
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(unet3d, "_require_cuda", lambda x: None)


@pytest.mark.parametrize("name", list(DDPM))
def test_jellyfish_ddpm_host_logic_steps(name, golden_dir, monkeypatch):
    _emulate(monkeypatch)
    check_ddpm_steps(np.load(os.path.join(golden_dir, name + ".npz")), name, "cpu", "3xtf32", 2e-4)


def test_jellyfish_ddpm_host_logic_loop(golden_dir, monkeypatch):
    _emulate(monkeypatch)
    check_ddpm_loop(np.load(os.path.join(golden_dir, "jelly_ddpm_alpha.npz")), "jelly_ddpm_alpha", "cpu")


def test_jellyfish_repaint_host_logic_loop(golden_dir, monkeypatch):
    """cond_steps == 0: unconditional model + repaint conditioning (jf.py:865-873), whole loop against the reference trace (the
    recorded draws include the q_sample noise of the three re-imposed conditions, in the reference's order)."""
    _emulate(monkeypatch)
    check_ddpm_loop(np.load(os.path.join(golden_dir, "jelly_ddpm_repaint.npz")), "jelly_ddpm_repaint", "cpu", cond_steps=0)


def test_jellyfish_ddim_host_logic(golden_dir, monkeypatch):
    _emulate(monkeypatch)
    check_ddim(np.load(os.path.join(golden_dir, "jelly_ddim_alpha.npz")), "jelly_ddim_alpha", "cpu", "3xtf32", 2e-4)


def test_unimplemented_options_raise():
    net = Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=7, out_dim=4)
    with pytest.raises(NotImplementedError):
        dj.GaussianDiffusion(net, image_size=16, only_vis_pressure=True)
    d = dj.GaussianDiffusion([net, net], image_size=16, frames=4, cond_steps=1, timesteps=4, eval_2ddpm=True)
    with pytest.raises(NotImplementedError):       # the reference's p_sample returns None for it (jf.py:789)
        d._ddpm_step(None, 0, None, "recurrence")


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-4), ("tf32", 1e-2)])
@pytest.mark.parametrize("name", list(DDPM))
def test_jellyfish_ddpm_steps_gpu(name, precision, tol, golden_dir):
    check_ddpm_steps(np.load(os.path.join(golden_dir, name + ".npz")), name, "cuda", precision, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(DDPM))
def test_jellyfish_ddpm_loop_gpu(name, golden_dir):
    check_ddpm_loop(np.load(os.path.join(golden_dir, name + ".npz")), name, "cuda")


@pytest.mark.gpu
def test_jellyfish_repaint_loop_gpu(golden_dir):
    check_ddpm_loop(np.load(os.path.join(golden_dir, "jelly_ddpm_repaint.npz")), "jelly_ddpm_repaint", "cuda", cond_steps=0)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("3xtf32", 2e-4), ("tf32", 1e-2)])
@pytest.mark.parametrize("name", list(DDIM))
def test_jellyfish_ddim_gpu(name, precision, tol, golden_dir):
    check_ddim(np.load(os.path.join(golden_dir, name + ".npz")), name, "cuda", precision, tol)
