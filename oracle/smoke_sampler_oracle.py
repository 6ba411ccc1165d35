"""CPU restatement (plain PyTorch fp32) of the smoke `GaussianDiffusion` sampling path.

TEST INFRASTRUCTURE ONLY — never imported by the product path.  Follows
/root/reference/diffusion/diffusion_2d_smoke.py (cited as `smoke.py:line`) and the stock guidance function of
/root/reference/inference/inference_2d_smoke.py:30-44.  Pinned by tests/golden/make_golden.py against the UNMODIFIED
reference `GaussianDiffusion.p_sample` / `ddim_sample` executed in the build container (teacher-forced single steps
with recorded noise) — see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

SMOKE_RESCALER = (2.0, 18.0, 20.0, 16.0, 20.0, 1.0)  # dataset/data_2d.py:167


# ---- schedules (smoke.py:414-448) ---------------------------------------------------------------------------

def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


def make_schedule(timesteps: int = 1000, beta_schedule: str = "sigmoid") -> Dict[str, torch.Tensor]:
    """The fp32 buffers GaussianDiffusion.__init__ registers (smoke.py:507-552); computed in fp64 then cast."""
    fn = {"linear": linear_beta_schedule, "cosine": cosine_beta_schedule, "sigmoid": sigmoid_beta_schedule}
    if beta_schedule not in fn:
        raise ValueError(f"unknown beta schedule {beta_schedule}")
    betas = fn[beta_schedule](timesteps)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.nn.functional.pad(ac[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    bufs = dict(
        betas=betas,
        alphas_cumprod=ac,
        alphas_cumprod_prev=ac_prev,
        sqrt_alphas_cumprod=torch.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=torch.sqrt(1.0 - ac),
        log_one_minus_alphas_cumprod=torch.log(1.0 - ac),
        sqrt_recip_alphas_cumprod=torch.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=torch.sqrt(1.0 / ac - 1),
        posterior_variance=post_var,
        posterior_log_variance_clipped=torch.log(post_var.clamp(min=1e-20)),
        posterior_mean_coef1=betas * torch.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
    )
    return {k: v.to(torch.float32) for k, v in bufs.items()}


# ---- guidance (inference_2d_smoke.py:30-44) -----------------------------------------------------------------

def guidance_fn(x: torch.Tensor, rescaler: torch.Tensor, w_energy: float = 0.0) -> torch.Tensor:
    """Gradient of J w.r.t. the RESCALED tensor x*R (no chain-rule factor R)."""
    with torch.enable_grad():
        x = x.detach().clone().requires_grad_()
        xr = x * rescaler
        guidance_success = xr[:, -1, -1].mean((-1, -2)).sum()
        guidance_energy = xr[:, :, 3:5].square().mean((1, 2, 3, 4)).sum()
        guidance = -guidance_success + w_energy * guidance_energy
        (g,) = torch.autograd.grad(guidance, xr, grad_outputs=torch.ones_like(guidance))
    return g.detach()


def _ext(a, t, like):
    return a[t].reshape(-1, *((1,) * (like.dim() - 1)))


def model_predictions(sched, x, t, eps_joint, eps_w, design_fn, *, design_guidance="standard",
                      standard_fixed_ratio=0.01, coeff_ratio=0.1, w_prob_exp=1.0,
                      clip_x_start=False, rederive_pred_noise=False):
    """smoke.py:610-656 given the two network outputs. eps_w: [B,F,2,H,W] scattered into channels 3:5."""
    pred_noise_w = torch.zeros_like(eps_joint)
    pred_noise_w[:, :, 3:5] = eps_w
    clip = (lambda v: v.clamp(-1.0, 1.0)) if clip_x_start else (lambda v: v)
    sr = _ext(sched["sqrt_recip_alphas_cumprod"], t, x)
    srm1 = _ext(sched["sqrt_recipm1_alphas_cumprod"], t, x)
    x_start = clip(sr * x - srm1 * eps_joint)
    g = design_fn(x_start)
    if design_guidance == "standard":
        grad_final = standard_fixed_ratio * g + (w_prob_exp - 1) * pred_noise_w
    elif design_guidance == "standard-alpha":
        eta = _ext(coeff_ratio * sched["betas"].flip(0), t, x)
        grad_final = eta * g + (w_prob_exp - 1) * pred_noise_w
    else:
        raise RuntimeError(design_guidance)
    pred_noise = eps_joint + grad_final
    x_start = clip(sr * x - srm1 * pred_noise)
    if clip_x_start and rederive_pred_noise:
        pred_noise = (sr * x - x_start) / srm1
    return pred_noise, x_start


def p_sample_step(sched, x, t_int: int, eps_joint, eps_w, noise, init, design_fn, **kw):
    """One iteration of p_sample_loop (smoke.py:671-699, :717-720) given network outputs and the noise draw."""
    b = x.shape[0]
    t = torch.full((b,), t_int, dtype=torch.long)
    _, x_start = model_predictions(sched, x, t, eps_joint, eps_w, design_fn, **kw)
    x_start = x_start.clamp(-1.0, 1.0)
    mean = _ext(sched["posterior_mean_coef1"], t, x) * x_start + _ext(sched["posterior_mean_coef2"], t, x) * x
    logvar = _ext(sched["posterior_log_variance_clipped"], t, x)
    z = noise if t_int > 0 else 0
    pred = mean + (0.5 * logvar).exp() * z
    pred[:, 0, 0] = init
    return pred, x_start


def ddim_times(total_timesteps: int, sampling_timesteps: int):
    """smoke.py:729-731."""
    times = torch.linspace(-1, total_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_step(sched, x, time: int, time_next: int, eps_joint, eps_w, noise, init, design_fn, eta: float, **kw):
    """One iteration of ddim_sample (smoke.py:739-775)."""
    b = x.shape[0]
    t = torch.full((b,), time, dtype=torch.long)
    pred_noise, x_start = model_predictions(sched, x, t, eps_joint, eps_w, design_fn,
                                            clip_x_start=True, rederive_pred_noise=True, **kw)
    if time_next < 0:
        return x_start, x_start
    alpha = sched["alphas_cumprod"][time]
    alpha_next = sched["alphas_cumprod"][time_next]
    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = (1 - alpha_next - sigma ** 2).sqrt()
    img = x_start * alpha_next.sqrt() + c * pred_noise + sigma * noise
    img[:, 0, 0] = init
    return img, x_start


@torch.no_grad()
def p_sample_loop(sched, nets: Callable, shape, init, design_fn, num_timesteps: int, generator=None, **kw):
    """smoke.py:702-723 with `nets(x, t_long) -> (eps_joint, eps_w)`; noise drawn with torch.randn in the reference's
    order (initial state first, then one draw per step with t > 0)."""
    x = torch.randn(shape, generator=generator)
    x[:, 0, 0] = init
    for t in reversed(range(num_timesteps)):
        tt = torch.full((shape[0],), t, dtype=torch.long)
        ej, ew = nets(x, tt)
        z = torch.randn(shape, generator=generator) if t > 0 else None
        x, _ = p_sample_step(sched, x, t, ej, ew, z, init, design_fn, **kw)
    return x
