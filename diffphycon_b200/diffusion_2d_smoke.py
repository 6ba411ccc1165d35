"""B200-native `GaussianDiffusion` for the 2-D smoke task — drop-in for diffusion/diffusion_2d_smoke.py:451-806
(cited below as smoke.py:line).  Same constructor, same registered buffers (so `Trainer.load` checkpoints load), same
`sample / p_sample_loop / ddim_sample / p_sample` call surface.  The two U-Net forwards run on the sm_100a kernels
of `Unet3D_with_Conv3D`; guidance, prior re-weighting, x0 clamp, posterior / DDIM update and the re-imposed initial
condition are ONE fused elementwise kernel per step (dpc_ddpm_guided_step / dpc_ddim_guided_step).

Noise is drawn with `torch.randn` in the reference's call order (smoke.py:668-669, :707, :684, :736, :768) so fixed
seeds reproduce the reference's noise stream on the same device.
"""
from __future__ import annotations

import math
from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .trainer_shim import SmokeTrainer as Trainer  # noqa: F401  (load-only stand-in, see trainer_shim.py)

ModelPrediction = namedtuple('ModelPrediction', ['pred_noise', 'pred_x_start'])

SMOKE_RESCALER = (2.0, 18.0, 20.0, 16.0, 20.0, 1.0)  # dataset/data_2d.py:167


# ---- beta schedules (smoke.py:414-448), float64 ---------------------------------------------------------------
def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1, clamp_min=1e-5):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


class StockSmokeGuidance:
    """The stock smoke design objective of inference/inference_2d_smoke.py:30-44:
        J = -sum_b mean_{h,w}(x*R)[b,-1,-1] + w_energy * sum_b mean((x*R)[b,:,3:5]^2),   returns dJ/d(x*R).
    Passing an instance as `design_fn` lets the sampler evaluate the gradient in closed form inside the fused step
    kernel.  Calling it evaluates the same objective with autograd (used to cross-check)."""

    def __init__(self, rescaler=SMOKE_RESCALER, w_energy: float = 0.0):
        r = torch.as_tensor(rescaler, dtype=torch.float32).reshape(-1)
        assert r.numel() == 6
        self.rescaler = tuple(float(v) for v in r)
        self.w_energy = float(w_energy)

    def __call__(self, x, low=None, init=None, init_u=None):
        R = torch.tensor(self.rescaler, dtype=x.dtype, device=x.device).reshape(1, 1, 6, 1, 1)
        with torch.enable_grad():
            xr = x * R
            succ = xr[:, -1, -1].mean((-1, -2)).sum()
            energy = xr[:, :, 3:5].square().mean((1, 2, 3, 4)).sum()
            J = -succ + self.w_energy * energy
            (g,) = torch.autograd.grad(J, xr, grad_outputs=torch.ones_like(J))
        return g


class GaussianDiffusion(nn.Module):
    """Constructor: smoke.py:452-472.

    Scope: the SAMPLING surface of the reference class (schedules / registered buffers, sample, p_sample_loop, ddim_sample,
    p_sample).  The training-side methods (forward, p_losses, q_sample; smoke.py:791-839) are not provided: calling the module
    raises, so a reference training script fails at once instead of at backward()."""

    def forward(self, *a, **k):
        raise NotImplementedError("diffphycon_b200.GaussianDiffusion is a sampling engine (sample / p_sample_loop / ddim_sample); "
                                  "training (forward / p_losses / q_sample, smoke.py:791-839) stays with the reference class")

    def __init__(self, model, *, image_size, frames, timesteps=1000, sampling_timesteps=None, loss_type='l1',
                 objective='pred_noise', beta_schedule='sigmoid', schedule_fn_kwargs=dict(), ddim_sampling_eta=0.,
                 min_snr_loss_weight=False, min_snr_gamma=5, standard_fixed_ratio=0.01, coeff_ratio=0.1,
                 eval_2ddpm=False, w_prob_exp=1.0, device=None):
        super().__init__()
        if eval_2ddpm:
            self.model_joint, self.model_thetas = model
            self.channels = self.model_joint.channels
            self.self_condition = self.model_joint.self_condition
        else:
            self.model = model
            self.channels = self.model.channels
            self.self_condition = self.model.self_condition
        self.is_w_model = self.channels == 2
        self.image_size = image_size
        self.frames = frames
        self.objective = objective
        self.standard_fixed_ratio = standard_fixed_ratio
        self.coeff_ratio = coeff_ratio
        self.eval_2ddpm = eval_2ddpm
        self.w_prob_exp = w_prob_exp
        assert objective in {'pred_noise', 'pred_x0', 'pred_v', 'pred_optimal_design'}
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
        if objective in ('pred_noise', 'pred_optimal_design'):
            register_buffer('loss_weight', clipped / snr)
        elif objective == 'pred_x0':
            register_buffer('loss_weight', clipped)
        elif objective == 'pred_v':
            register_buffer('loss_weight', clipped / (snr + 1))
        self.progress = False      # tqdm bar like the reference's (smoke.py:717) when True
        self._host_sched = None    # CPU copies of the schedule buffers: per-step scalars without device syncs
        # Step orchestration: with use_cuda_graph one denoising step (time-table lookup, both U-Nets, noise
        # draw, fused update) is captured ONCE in a CUDA graph and replayed for the whole schedule — one graph launch per step
        # instead of ~270 kernel launches through Python/ctypes; two_streams runs the two U-Nets on two captured streams
        # (None = automatic: batches of <= 16 trajectories, where one network alone does not fill the GPU's tail waves).
        # Only the stock guidance (StockSmokeGuidance) is graph-capturable; other design_fn callables use the eager loop.
        self.use_cuda_graph = False
        self.two_streams = None
        self._graphs = {}

    # ---- host-side scalar schedule ------------------------------------------------------------------------------
    def _sched(self):
        if self._host_sched is None:
            names = ('betas', 'alphas_cumprod', 'sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod',
                     'posterior_log_variance_clipped', 'posterior_mean_coef1', 'posterior_mean_coef2')
            self._host_sched = {n: getattr(self, n).detach().float().cpu() for n in names}
        return self._host_sched

    def _apply(self, fn, *a, **k):
        self._host_sched = None
        return super()._apply(fn, *a, **k)

    def _coefs(self, t: int, design_fn, design_guidance: str) -> _lib.StepCoefs:
        s = self._sched()
        c = _lib.StepCoefs()
        c.sqrt_recip_alphas_cumprod = float(s['sqrt_recip_alphas_cumprod'][t])
        c.sqrt_recipm1_alphas_cumprod = float(s['sqrt_recipm1_alphas_cumprod'][t])
        if design_guidance == "standard":
            c.guidance_coef = float(torch.tensor(self.standard_fixed_ratio, dtype=torch.float32))
        elif design_guidance == "standard-alpha":
            # coeff_ratio * betas.flip(0), gathered at t (smoke.py:632-633): fp32 tensor times python scalar
            c.guidance_coef = float((self.coeff_ratio * s['betas'].flip(0))[t])
        else:
            raise RuntimeError(f"unknown design_guidance {design_guidance!r}")  # the reference has a bare `raise`
        c.prior_coef = float(torch.tensor(self.w_prob_exp - 1, dtype=torch.float32))
        if isinstance(design_fn, StockSmokeGuidance):
            c.w_energy = float(torch.tensor(design_fn.w_energy, dtype=torch.float32))
            for i in range(6):
                c.rescaler[i] = design_fn.rescaler[i]
        return c

    # ---- network evaluation ---------------------------------------------------------------------------------------
    def _eps(self, x, t: int):
        """model_joint(x, t), model_thetas(x[:, :, 3:5], t) — smoke.py:611-613 (quirk: `self.model` is never used)."""
        if not self.eval_2ddpm:
            raise RuntimeError("smoke sampling requires eval_2ddpm=True with [model_joint, model_w] (smoke.py:611-613)")
        b = x.shape[0]
        tt = torch.full((b,), t, device=x.device, dtype=torch.long)
        eps_j = self.model_joint(x, tt)
        mw = self.model_thetas
        eps_w = torch.empty(b, x.shape[1], mw.out_dim, x.shape[3], x.shape[4], dtype=torch.float32, device=x.device)
        mw.forward_slice(x, 3, tt, eps_w)
        return eps_j, eps_w

    def _user_gradient(self, x, eps_j, c: _lib.StepCoefs, clip: bool, design_fn, low, init, init_u):
        """Generic design_fn callable: evaluate it on x_start under autograd, like smoke.py:620-627."""
        x_start = torch.empty_like(x)
        _lib.predict_x_start(x, eps_j, c.sqrt_recip_alphas_cumprod, c.sqrt_recipm1_alphas_cumprod, clip, x_start)
        with torch.enable_grad():
            x_clone = x_start.clone().detach().requires_grad_()
            g = design_fn(x_clone, low=low, init=init, init_u=init_u)
        return g.detach().float().contiguous()

    def sample_noise(self, shape, device):
        return torch.randn(shape, device=device)

    # ---- DDPM -----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    @_lib.device_guarded
    def p_sample(self, shape, x, t: int, x_self_cond=None, clip_denoised=True, design_fn=None,
                 design_guidance="standard", low=None, init=None, init_u=None, _impose_init=False):
        """smoke.py:671-699.  Returns (x_{t-1}, x_start)."""
        assert clip_denoised, "the reference always samples with clip_denoised=True"
        b, f, ch, h, w = x.shape
        x = x.contiguous()
        c = self._coefs(t, design_fn, design_guidance)
        s = self._sched()
        c.posterior_mean_coef1 = float(s['posterior_mean_coef1'][t])
        c.posterior_mean_coef2 = float(s['posterior_mean_coef2'][t])
        c.sigma = float((0.5 * s['posterior_log_variance_clipped'][t]).exp())
        c.add_noise = 1 if t > 0 else 0
        eps_j, eps_w = self._eps(x, t)
        g = None
        if not isinstance(design_fn, StockSmokeGuidance):
            g = self._user_gradient(x, eps_j, c, False, design_fn, low, init, init_u)
        noise = self.sample_noise(x.shape, x.device) if t > 0 else None
        pred = torch.empty_like(x)
        x_start = torch.empty_like(x)
        _lib.guided_step(False, x, eps_j, eps_w, noise, init if _impose_init else None, g, c, pred, x_start, b, f, h, w)
        return pred, x_start

    @torch.no_grad()
    def p_sample_loop(self, shape, design_fn=None, design_guidance="standard", return_all_timesteps=None, init=None,
                      init_u=None, control=None, low=None, device=None):
        """smoke.py:702-723."""
        b, f, c, h, w = shape
        device = self.betas.device
        x = self.sample_noise([b, f, c, h, w], device)
        assert init is not None
        init = init.to(device).float().contiguous()
        x[:, 0, 0] = init
        if self.use_cuda_graph and isinstance(design_fn, StockSmokeGuidance) and not self.progress and x.is_cuda:
            return self._graph_loop(x, init, design_fn, design_guidance, ddim=False)
        steps = reversed(range(0, self.num_timesteps))
        if self.progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps, desc='sampling loop time step', total=self.num_timesteps)
        for t in steps:
            # x[:, 0, 0] = init after every step (smoke.py:720) is fused into the step kernel
            x, _ = self.p_sample(shape, x, t, None, design_fn=design_fn, design_guidance=design_guidance, low=low,
                                 init=init, init_u=init_u, _impose_init=True)
        return x

    # ---- DDIM -----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def ddim_sample(self, shape, design_fn=None, design_guidance="standard", init=None, init_u=None, control=None,
                    low=None, device=None):
        """smoke.py:725-779."""
        batch, device = shape[0], self.betas.device
        total, steps_n, eta = self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta
        times = torch.linspace(-1, total - 1, steps=steps_n + 1)
        times = list(reversed(times.int().tolist()))
        time_pairs = list(zip(times[:-1], times[1:]))
        img = self.sample_noise(list(shape), device)
        init = init.to(device).float().contiguous()
        img[:, 0, 0] = init
        if self.use_cuda_graph and isinstance(design_fn, StockSmokeGuidance) and not self.progress and img.is_cuda:
            return self._graph_loop(img, init, design_fn, design_guidance, ddim=True)
        it = time_pairs
        if self.progress:
            from tqdm.auto import tqdm
            it = tqdm(time_pairs, desc='sampling loop time step')
        for time, time_next in it:
            img = self.ddim_step(img, time, time_next, design_fn=design_fn, design_guidance=design_guidance, init=init,
                                 init_u=init_u, low=low)
        return img

    @torch.no_grad()
    @_lib.device_guarded
    def ddim_step(self, img, time: int, time_next: int, design_fn=None, design_guidance="standard", init=None,
                  init_u=None, low=None, noise=None):
        """One iteration of the DDIM loop, smoke.py:739-775 (model_predictions with clip_x_start=True,
        rederive_pred_noise=True).  `noise` defaults to torch.randn_like(img) drawn where the reference draws it."""
        b, f, ch, h, w = img.shape
        s = self._sched()
        eta = self.ddim_sampling_eta
        c = self._coefs(time, design_fn, design_guidance)
        eps_j, eps_w = self._eps(img, time)
        g = None
        if not isinstance(design_fn, StockSmokeGuidance):
            g = self._user_gradient(img, eps_j, c, True, design_fn, low, init, init_u)
        if time_next < 0:
            c.last = 1
            noise = None
        else:
            alpha = s['alphas_cumprod'][time]
            alpha_next = s['alphas_cumprod'][time_next]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            cc = (1 - alpha_next - sigma ** 2).sqrt()
            c.sqrt_alpha_next, c.c, c.ddim_sigma, c.last = float(alpha_next.sqrt()), float(cc), float(sigma), 0
            if noise is None:
                noise = self.sample_noise(list(img.shape), img.device)
        out = torch.empty_like(img)
        _lib.guided_step(True, img.contiguous(), eps_j, eps_w, noise, init, g, c, out, None, b, f, h, w)
        return out

    # ---- CUDA-graph step orchestration ---------------------------------------------------------------------------------
    def _step_tables(self, design_fn, design_guidance, ddim):
        """Per-step (time, coefficients) of the whole schedule, in sampling order, exactly as p_sample / ddim_step set them."""
        s = self._sched()
        ts, cs = [], []
        if not ddim:
            for t in reversed(range(0, self.num_timesteps)):
                c = self._coefs(t, design_fn, design_guidance)
                c.posterior_mean_coef1 = float(s['posterior_mean_coef1'][t])
                c.posterior_mean_coef2 = float(s['posterior_mean_coef2'][t])
                c.sigma = float((0.5 * s['posterior_log_variance_clipped'][t]).exp())
                c.add_noise = 1 if t > 0 else 0
                ts.append(t)
                cs.append(c)
        else:
            times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
            times = list(reversed(times.int().tolist()))
            eta = self.ddim_sampling_eta
            for time, time_next in zip(times[:-1], times[1:]):
                c = self._coefs(time, design_fn, design_guidance)
                if time_next < 0:
                    c.last = 1
                else:
                    alpha, alpha_next = s['alphas_cumprod'][time], s['alphas_cumprod'][time_next]
                    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                    cc = (1 - alpha_next - sigma ** 2).sqrt()
                    c.sqrt_alpha_next, c.c, c.ddim_sigma, c.last = float(alpha_next.sqrt()), float(cc), float(sigma), 0
                ts.append(time)
                cs.append(c)
        return ts, cs

    def _graph_loop(self, x, init, design_fn, design_guidance, ddim):
        """The whole sampling loop as replays of one captured step.  Same arithmetic, same kernels and the same torch.randn
        draws as the eager loop (the final step draws a noise tensor it does not use)."""
        import ctypes
        dev = x.device
        shape = tuple(x.shape)
        b = shape[0]
        two = self.two_streams if self.two_streams is not None else b <= 16
        ts, cs = self._step_tables(design_fn, design_guidance, ddim)
        n = len(ts)
        csize = ctypes.sizeof(_lib.StepCoefs)
        raw = b"".join(bytes(c) for c in cs)
        key = (shape, str(dev), ddim, two, design_guidance, hash(raw), tuple(ts), getattr(self, "_noise_key", None),
               self.model_joint._param_key(), self.model_thetas._param_key())
        g = self._graphs.get(key)
        if g is None:
            self._graphs.clear()       # one captured schedule at a time: a graph pins its activation buffers
            rng = torch.cuda.get_rng_state(dev)     # warm-up and capture draw noise: leave the caller's RNG stream untouched
            g = _GraphedStep(self, shape, dev, ddim, two, ts, raw, csize, cs[0])
            torch.cuda.set_rng_state(rng, dev)
            self._graphs[key] = g
        return g.run(x, init, n)

    @torch.no_grad()
    def sample(self, batch_size=16, design_fn=None, design_guidance="standard", init=None, init_u=None, control=None,
               low=None, device=None):
        """smoke.py:781-789."""
        sample_fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        assert batch_size == init.shape[0]
        size = (batch_size, self.frames, self.channels, self.image_size, self.image_size)
        return sample_fn(size, design_fn, design_guidance, init=init, init_u=init_u, control=control, low=low,
                         device=device)


class _GraphedStep:
    """One denoising step captured in a CUDA graph: dpc_sampler_prepare (device-side step counter -> time tensor and step
    coefficients), joint and prior U-Net forwards (optionally on two streams), torch.randn, dpc_guided_step_dev in place."""

    def __init__(self, diff, shape, dev, ddim, two_streams, ts, raw_coefs, csize, coefs_host):
        self.diff, self.shape, self.ddim, self.two = diff, shape, ddim, two_streams
        b, f, c, h, w = shape
        self.n = len(ts)
        self.x = torch.randn(shape, device=dev)
        self.init = torch.zeros(b, h, w, device=dev)
        self.tt = torch.zeros(b, dtype=torch.long, device=dev)
        self.sidx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cur = torch.zeros(csize, dtype=torch.uint8, device=dev)
        self.t_table = torch.tensor(ts, dtype=torch.long, device=dev)
        self.c_table = torch.frombuffer(bytearray(raw_coefs), dtype=torch.uint8).to(dev)
        self.eps_w = torch.empty(b, f, diff.model_thetas.out_dim, h, w, device=dev)
        self.coefs_host = coefs_host
        self.stream = torch.cuda.Stream(device=dev)
        self.side = torch.cuda.Stream(device=dev) if two_streams else None
        cur_stream = torch.cuda.current_stream(dev)
        self.stream.wait_stream(cur_stream)
        with torch.cuda.stream(self.stream):          # warm-up on the capture stream: packs weights, fills the buffer pools
            for _ in range(2):
                self._body()
        cur_stream.wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        c0 = _lib.LaunchCounter.count
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body()
        self.kernels_per_replay = _lib.LaunchCounter.count - c0     # launches of this binding captured in the graph

    def _body(self):
        d = self.diff
        b, f, c, h, w = self.shape
        _lib.sampler_prepare(self.t_table, self.c_table, self.sidx, self.n, self.tt, b, self.cur)
        if self.side is not None:
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                d.model_thetas.forward_slice(self.x, 3, self.tt, self.eps_w)
            eps_j = d.model_joint(self.x, self.tt)
            main.wait_stream(self.side)
        else:
            eps_j = d.model_joint(self.x, self.tt)
            d.model_thetas.forward_slice(self.x, 3, self.tt, self.eps_w)
        noise = d.sample_noise(list(self.shape), self.x.device)    # captured: torch.randn (or the sharded global-noise slice)
        _lib.guided_step_dev(self.ddim, self.x, eps_j, self.eps_w, noise, self.init, self.coefs_host, self.cur, self.x, None,
                             b, f, h, w)

    def run(self, x0, init, nsteps):
        self.x.copy_(x0)
        self.init.copy_(init)
        self.sidx.zero_()
        for _ in range(nsteps):
            self.graph.replay()
            _lib.LaunchCounter.graph_launches += 1
            _lib.LaunchCounter.count += self.kernels_per_replay
        return self.x.clone()
