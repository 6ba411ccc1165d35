"""B200-native `GaussianDiffusion` for the jellyfish task — drop-in for diffusion/diffusion_2d_jellyfish.py:529-978 (cited
as jf.py:line) on the sampling path inference/inference_2d_jellyfish.py drives: two `Unet3D_with_Conv3D` forwards per step
(joint model on x = [state 3, boundary 3, theta 1], prior model on [state_cond 3, boundary 3, theta 1]), posterior on the
four diffused channels, post-hoc guidance  pred -= eta_J * g - eta_w * eps_w,  boundary update through `bd_updater` and the
re-imposed frame-0 / frame-(-1) conditions.  Same constructor and registered buffers, same `sample / p_sample_loop /
ddim_sample` call surface, same noise call order.

The elementwise / reduction work of a step is two fused kernels plus one that writes the boundary channels
(include/dpc_b200.h: dpc_jelly_x_start, dpc_jelly_step, dpc_jelly_write_bd).  `design_fn` and `bd_updater` are the caller's
PyTorch callables, as in the reference (the surrogate 2-D nets behind them are SURVEY.md 8(f) rank 1, not on this engine yet).

Reproduced quirks: the DDPM path hands `use_guidance_in_model_predictions` to `clip_x_start`
positionally (jf.py:760 vs :703), so in-model guidance is never enabled from p_sample; the post-hoc guidance broadcasts
eps_w over all four channels (jf.py:800-804) whereas the DDIM path pads it onto the theta channel only (jf.py:728-733);
`ddim_sample` returns the result of the second-to-last pair (jf.py:913-916 `continue` before `final_result` is refreshed),
so the last model evaluation does not influence the output and is skipped here.
cond_steps == 0 (unconditional model): the conditions are re-imposed as noisy conditions after every DDPM step (repaint, jf.py:865-873).
Not implemented (raise NotImplementedError): only_vis_pressure (hard-coded 64x64 shapes in the reference), objectives other than
pred_noise, 'recurrence' guidance (the reference's p_sample returns None for it, jf.py:789)."""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .trainer_shim import JellyfishTrainer as Trainer  # noqa: F401  (load-only stand-in, see trainer_shim.py)
from .diffusion_2d_smoke import cosine_beta_schedule, linear_beta_schedule, sigmoid_beta_schedule
from .jellyfish_nets import ForceUnet, JellyfishGuidance, Unet  # noqa: F401  (jf.py:276-481: same import surface as the reference module)

ModelPrediction = namedtuple('ModelPrediction', ['pred_noise', 'pred_noise_w', 'pred_x_start'])


def _f32(v) -> float:
    return float(torch.as_tensor(v, dtype=torch.float64).to(torch.float32))


class GaussianDiffusion(nn.Module):
    """Constructor: jf.py:530-559."""

    def __init__(self, model, *, image_size, frames=20, cond_steps=0, timesteps=1000, sampling_timesteps=None, loss_type='l1',
                 objective='pred_noise', beta_schedule='sigmoid', schedule_fn_kwargs=dict(), ddim_sampling_eta=0.,
                 auto_normalize=True, min_snr_loss_weight=False, min_snr_gamma=5, backward_steps=5, backward_lr=0.01,
                 standard_fixed_ratio=0.01, forward_fixed_ratio=0.01, coeff_ratio_J=0.3, coeff_ratio_w=0.3,
                 only_vis_pressure=False, eval_2ddpm=False, w_prob_exp=1.0, use_guidance_in_model_predictions=False,
                 return_all_timesteps=True, device=None):
        super().__init__()
        if only_vis_pressure:
            raise NotImplementedError("only_vis_pressure is not implemented (the reference hard-codes 64x64 there, jf.py:722)")
        if objective != 'pred_noise':
            raise NotImplementedError("only objective='pred_noise' is implemented on the sampling path")
        if eval_2ddpm:
            self.model_states, self.model_thetas = model
            self.channels = self.model_states.channels
            self.self_condition = self.model_states.self_condition
        else:
            self.model = model
            self.channels = self.model.channels
            self.self_condition = self.model.self_condition
        self.eval_2ddpm = eval_2ddpm
        self.frames, self.cond_steps, self.image_size, self.objective = frames, cond_steps, image_size, objective
        self.backward_steps, self.backward_lr = backward_steps, backward_lr
        self.standard_fixed_ratio, self.forward_fixed_ratio = standard_fixed_ratio, forward_fixed_ratio
        self.coeff_ratio_J, self.coeff_ratio_w = coeff_ratio_J, coeff_ratio_w
        self.only_vis_pressure = False
        self.w_prob_exp = w_prob_exp
        self.use_guidance_in_model_predictions = use_guidance_in_model_predictions
        self.return_all_timesteps = return_all_timesteps
        fns = {'linear': linear_beta_schedule, 'cosine': cosine_beta_schedule, 'sigmoid': sigmoid_beta_schedule}
        if beta_schedule not in fns:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        betas = fns[beta_schedule](timesteps, **schedule_fn_kwargs).to(device)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.sampling_timesteps = timesteps if sampling_timesteps is None else sampling_timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta

        def register_buffer(name, val):
            self.register_buffer(name, val.to(torch.float32))

        register_buffer('betas', betas)
        register_buffer('alphas_cumprod', alphas_cumprod)
        register_buffer('alphas_cumprod_prev', alphas_cumprod_prev)
        register_buffer('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        register_buffer('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        register_buffer('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        register_buffer('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        register_buffer('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        register_buffer('posterior_variance', posterior_variance)
        register_buffer('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        register_buffer('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        register_buffer('posterior_mean_coef2', (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        snr = alphas_cumprod / (1 - alphas_cumprod)
        clipped = snr.clone()
        if min_snr_loss_weight:
            clipped.clamp_(max=min_snr_gamma)
        register_buffer('loss_weight', clipped / snr)
        self.progress = False
        self._host_sched = None

    def _sched(self):
        if self._host_sched is None:
            names = ('betas', 'alphas_cumprod', 'sqrt_alphas_cumprod', 'sqrt_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod',
                     'sqrt_recipm1_alphas_cumprod', 'posterior_log_variance_clipped', 'posterior_mean_coef1', 'posterior_mean_coef2')
            self._host_sched = {n: getattr(self, n).detach().float().cpu() for n in names}
        return self._host_sched

    def _apply(self, fn, *a, **k):
        self._host_sched = None
        return super()._apply(fn, *a, **k)

    def sample_noise(self, shape, device):
        return torch.randn(shape, device=device)

    # ---- per-step pieces ------------------------------------------------------------------------------------------
    def _models(self):
        if not self.eval_2ddpm:
            raise RuntimeError("jellyfish sampling requires eval_2ddpm=True with [model_states, model_thetas] (jf.py:704-706)")
        return self.model_states, self.model_thetas

    def _guidance_scalars(self, t: int, design_guidance: str, ddim: bool):
        """(ga, gb) of  ga * g - gb * eps_w  — jf.py:734-741 (in model, DDIM) and :797-802 (post hoc, DDPM)."""
        s = self._sched()
        if design_guidance == "standard":
            ga = _f32(self.standard_fixed_ratio)
            gb = -_f32(self.w_prob_exp - 1) if ddim else ga
        elif design_guidance == "standard-alpha":
            ga = float((self.coeff_ratio_J * s['betas'].flip(0))[t])
            gb = float((self.coeff_ratio_w * s['betas'].flip(0))[t])
        else:
            raise RuntimeError(f"unknown design_guidance {design_guidance!r}")   # bare `raise` in the reference
        return ga, gb

    def _design_gradient(self, design_fn, x_start, bd_0_expand):
        with torch.enable_grad():                                              # jf.py:792-794
            x_clone = x_start.clone().detach().requires_grad_()
            g = design_fn(x_clone, bd_0_expand)
        return g.detach().float().contiguous()

    class _State:
        """Buffers of one sampling run (the engine owns and ping-pongs them; SURVEY.md 8(b) ownership row)."""

    def _begin(self, shape, cond, thetas_0, bd_updater):
        b, f, c, h, w = shape
        device = self.betas.device
        assert cond is not None
        st = self._State()
        st.state_0 = cond[0].to(device).float().contiguous()
        st.bd_0 = cond[1].to(device).float().contiguous()
        noise_state = self.sample_noise([b, f, 3, h, w], device)               # call order of jf.py:826-832
        noise_bd = self.sample_noise([b, f, 3, h, w], device)
        st.thetas_0 = thetas_0.to(device).float().contiguous()
        noisy_thetas = self.sample_noise([b, f, 1, h, w], device)
        th = st.thetas_0.reshape(b, 1, 1, 1, 1).expand(-1, 1, 1, h, w)
        st.bd_0_expand = st.bd_0.unsqueeze(1).expand(-1, self.frames, -1, -1, -1)
        if isinstance(bd_updater, nn.Module):
            bd_updater.to(device)
            bd_updater.eval()
        cs = self.cond_steps
        if cs > 0:                                                              # conditional model (jf.py:838-842)
            noise_state[:, :cs] = st.state_0.unsqueeze(1)
            noise_bd[:, :cs] = st.bd_0.unsqueeze(1)
            noisy_thetas[:, :cs] = th
            noisy_thetas[:, -cs:] = th
        st.x = torch.cat([noise_state, noise_bd, noisy_thetas], dim=2).contiguous()
        st.x_next = torch.empty_like(st.x)
        st.x_w = st.x.clone()
        st.x_w[:, :, :3] = st.state_0.unsqueeze(1)                              # state_cond, jf.py:842-843
        st.bd_flat = st.bd_0_expand.reshape(b * f, 3, h, w).contiguous()
        st.x_start = torch.empty(b, f, 4, h, w, device=device)
        st.dtheta = torch.empty(b, f, device=device)
        st.theta_mean = torch.empty(b, f, device=device)
        st.bd_updater = bd_updater
        return st

    def _finish_step(self, st, eps_j, eps_w, g, noise, ga, gb, c1, c2, sigma, ddim):
        b, f = st.x.shape[:2]
        _lib.jelly_step(st.x, st.x_start, eps_j, eps_w, g, noise, st.state_0, st.thetas_0, st.x_next, st.x_w, st.dtheta,
                        st.theta_mean, ga, gb, c1, c2, sigma, ddim, self.cond_steps)
        pred_bd = st.bd_updater(st.bd_flat, st.dtheta.reshape(b * f))           # update_bd, jf.py:809-817
        pred_bd = pred_bd.detach().float().reshape(b * f, 3, *st.x.shape[-2:]).contiguous()
        _lib.jelly_write_bd(pred_bd, st.bd_0, st.x_next, st.x_w, self.cond_steps)
        st.x, st.x_next = st.x_next, st.x

    def _repaint(self, st, t: int):
        """cond_steps == 0 (unconditional model, jf.py:865-873): the conditions are re-imposed as NOISY conditions, q_sample(., t)
        with fresh noise, on frame 0 (state, boundary, theta) and on the last frame (theta).  Same draw order as the reference
        (state, boundary, theta) and the same two rounded products + rounded sum (dpc_renoise)."""
        s = self._sched()
        a, bb = float(s['sqrt_alphas_cumprod'][t]), float(s['sqrt_one_minus_alphas_cumprod'][t])
        dev = st.x.device
        b, _, _, h, w = st.x.shape

        def q_sample(x0):
            x0 = x0.contiguous()
            z = self.sample_noise(list(x0.shape), dev).contiguous()
            out = torch.empty_like(x0)
            _lib.renoise(x0, z, a, bb, out)
            return out

        s0, b0 = q_sample(st.state_0), q_sample(st.bd_0)
        th = q_sample(st.thetas_0.reshape(b, 1, 1, 1, 1).expand(-1, 1, 1, h, w))[:, 0, 0]   # [B,H,W]
        for buf in (st.x, st.x_w):
            buf[:, 0, 3:6] = b0
            buf[:, 0, 6] = th
            buf[:, -1, 6] = th
        st.x[:, 0, 0:3] = s0
        tm = th.mean(dim=2).mean(dim=1)                                         # jf.py:875: mean over W, then over H
        st.theta_mean[:, 0] = tm
        st.theta_mean[:, -1] = tm

    def _eps(self, st, t: int):
        mj, mw = self._models()
        tt = torch.full((st.x.shape[0],), t, device=st.x.device, dtype=torch.long)
        return mj(st.x, tt), mw(st.x_w, tt)

    # ---- DDPM (jf.py:776-806, :819-881) ---------------------------------------------------------------------------------
    def _ddpm_step(self, st, t: int, design_fn, design_guidance, clip_denoised=True):
        if "recurrence" in design_guidance:
            raise NotImplementedError("'recurrence' guidance is not implemented")
        s = self._sched()
        eps_j, eps_w = self._eps(st, t)
        clip = bool(clip_denoised) or bool(self.use_guidance_in_model_predictions)   # quirk 4
        _lib.jelly_x_start(st.x, eps_j, st.x_start, float(s['sqrt_recip_alphas_cumprod'][t]),
                           float(s['sqrt_recipm1_alphas_cumprod'][t]), clip)
        noise = self.sample_noise(st.x_start.shape, st.x.device) if t > 0 else None
        g, ga, gb = None, 0.0, 0.0
        if not self.use_guidance_in_model_predictions and design_fn is not None:
            if not design_guidance.startswith("standard"):
                raise RuntimeError(f"unknown design_guidance {design_guidance!r}")   # NameError in the reference (jf.py:804)
            ga, gb = self._guidance_scalars(t, design_guidance, ddim=False)
            g = self._design_gradient(design_fn, st.x_start, st.bd_0_expand)
        sigma = float((0.5 * s['posterior_log_variance_clipped'][t]).exp())
        self._finish_step(st, None, eps_w, g, noise, ga, gb, float(s['posterior_mean_coef1'][t]),
                          float(s['posterior_mean_coef2'][t]), sigma, False)
        if self.cond_steps <= 0:
            self._repaint(st, t)

    @torch.no_grad()
    def p_sample_loop(self, shape, design_fn=None, design_guidance="standard", return_all_timesteps=None, cond=None,
                      thetas_0=None, bd_updater=None, device=None):
        with _lib.on_device(self.betas.device):
            st = self._begin(shape, cond, thetas_0, bd_updater)
            steps = reversed(range(0, self.num_timesteps))
            if self.progress:
                from tqdm.auto import tqdm
                steps = tqdm(steps, desc='sampling loop time step', total=self.num_timesteps)
            for t in steps:
                self._ddpm_step(st, t, design_fn, design_guidance)
            return [st.x[:, :, :3], st.theta_mean.clone()]

    # ---- DDIM (jf.py:883-966) -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def ddim_sample(self, shape, design_fn=None, design_guidance="standard", return_all_timesteps=None, cond=None,
                    thetas_0=None, bd_updater=None, device=None):
        if return_all_timesteps:
            raise NotImplementedError("return_all_timesteps stacks tensors with lists in the reference (jf.py:964) and fails")
        assert design_fn is not None, "the DDIM path calls design_fn unconditionally (jf.py:735)"
        with _lib.on_device(self.betas.device):
            eta = self.ddim_sampling_eta
            times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
            times = list(reversed(times.int().tolist()))
            pairs = list(zip(times[:-1], times[1:]))
            st = self._begin(shape, cond, thetas_0, bd_updater)
            s = self._sched()
            if self.progress:
                from tqdm.auto import tqdm
                pairs = tqdm(pairs, desc='sampling loop time step')
            for time, time_next in pairs:
                if time_next < 0:
                    continue    # the reference evaluates the models once more but returns the previous pair's result
                eps_j, eps_w = self._eps(st, time)
                _lib.jelly_x_start(st.x, eps_j, st.x_start, float(s['sqrt_recip_alphas_cumprod'][time]),
                                   float(s['sqrt_recipm1_alphas_cumprod'][time]), False)
                g = self._design_gradient(design_fn, st.x_start, st.bd_0_expand)
                ga, gb = self._guidance_scalars(time, design_guidance, ddim=True)
                alpha, alpha_next = s['alphas_cumprod'][time], s['alphas_cumprod'][time_next]
                sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                c = (1 - alpha_next - sigma ** 2).sqrt()
                noise = self.sample_noise(st.x_start.shape, st.x.device)
                self._finish_step(st, eps_j, eps_w, g, noise, ga, gb, float(alpha_next.sqrt()), float(c), float(sigma), True)
            return [st.x[:, :, :3], st.theta_mean.clone()]

    @torch.no_grad()
    def sample(self, batch_size=16, design_fn=None, design_guidance="standard", return_all_timesteps=False, cond=None,
               thetas_0=None, bd_updater=None, device=None):
        """jf.py:968-978: returns [pred_states [B,F,3,H,W], pred_theta [B,F]]."""
        image_size, channels, frames = self.image_size, self.channels // 2, self.frames
        sample_fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        batch_size = cond[0].shape[0]
        return sample_fn((batch_size, frames, channels, image_size, image_size), design_fn, design_guidance,
                         return_all_timesteps=return_all_timesteps, cond=cond, thetas_0=thetas_0, bd_updater=bd_updater,
                         device=device)

    def q_sample(self, x_start, t, noise=None):
        """jf.py:1000-1006 (training-side helper kept for callers; plain tensor ops on the caller's device)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        ex = lambda a: a.gather(-1, t).reshape(t.shape[0], *((1,) * (x_start.dim() - 1)))
        return ex(self.sqrt_alphas_cumprod) * x_start + ex(self.sqrt_one_minus_alphas_cumprod) * noise
