"""TEST INFRASTRUCTURE: a CPU emulation of the C-ABI entry points of libdpc_b200.so, written with torch ops.

It lets the CPU test suite run the HOST logic of diffphycon_b200 (weight packing, tap tables, transposed-conv parity
classes, buffer orchestration, scale/shift offsets, sampler coefficients) against the golden vectors without a GPU.
It is installed by monkeypatching `diffphycon_b200._lib` inside tests only; the product never imports this file and
has no CPU path.  Each emulator follows the contract documented in include/dpc_b200.h.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from oracle import unet3d_oracle as uo

HEADS_DIM = 32


def _view(ptr, numel, dtype=np.float32):
    if not ptr:
        return None
    ctype = {np.float32: ctypes.c_float, np.float64: ctypes.c_double, np.int32: ctypes.c_int32}[dtype]
    arr = np.ctypeslib.as_array((ctype * numel).from_address(ptr))
    return torch.from_numpy(arr)


TC_ACCEPTS = False      # tests set this to emulate a dpc_conv3d_tcgen05 that serves the shape (returns True, GroupNorm(8) stats only)


def conv(p, tcgen05=False, tc_only=False):
    ran_tc = bool(tcgen05 and TC_ACCEPTS and (not p.gn_stats or p.gn_groups == 8))
    if tc_only and not ran_tc:
        return False                                          # declined (-2): nothing is launched
    B, Fi, Hi, Wi, C1, C2 = p.B, p.Fi, p.Hi, p.Wi, p.C1, p.C2
    Fo, Ho, Wo = p.Fo, p.Ho, p.Wo
    cin = C1 + C2
    x = _view(p.x1, B * Fi * Hi * Wi * C1).reshape(B, Fi, Hi, Wi, C1)
    if C2:
        x2 = _view(p.x2, B * Fi * Hi * Wi * C2).reshape(B, Fi, Hi, Wi, C2)
        x = torch.cat([x, x2], dim=-1)
    w = _view(p.w, p.Npad * p.Kpad).reshape(p.Npad, p.Kpad)
    taps = _view(p.taps, p.ntaps * 4, np.int32).reshape(p.ntaps, 4)
    acc = torch.zeros(B, Fo, Ho, Wo, p.Cout, dtype=torch.float64)
    fo = torch.arange(Fo)
    ho = torch.arange(Ho)
    wo = torch.arange(Wo)
    for t in range(p.ntaps):
        dt, dh, dw, delta = [int(v) for v in taps[t]]
        assert delta == (dt * Hi + dh) * Wi + dw
        fi = fo * p.st + dt - p.pt
        hi = ho * p.sh + dh - p.ph
        wi = wo * p.sw + dw - p.pw
        vf, vh, vw = (fi >= 0) & (fi < Fi), (hi >= 0) & (hi < Hi), (wi >= 0) & (wi < Wi)
        g = x[:, fi.clamp(0, Fi - 1)][:, :, hi.clamp(0, Hi - 1)][:, :, :, wi.clamp(0, Wi - 1)]
        mask = (vf[:, None, None] & vh[None, :, None] & vw[None, None, :]).to(g.dtype)
        g = g * mask[None, :, :, :, None]
        wt = w[: p.Cout, t * cin:(t + 1) * cin].double()
        acc += g.double() @ wt.t()
    if p.bias:
        acc += _view(p.bias, p.Cout).double()
    Hf, Wf = p.Hfull, p.Wfull
    hs = slice(p.oh_off, Hf, p.oh_mul)
    ws = slice(p.ow_off, Wf, p.ow_mul)
    if p.out_layout == 0:
        y = _view(p.y, B * Fo * Hf * Wf * p.Cout).reshape(B, Fo, Hf, Wf, p.Cout)
        if p.residual:
            r = _view(p.residual, B * Fo * Hf * Wf * p.Cout).reshape(B, Fo, Hf, Wf, p.Cout)[:, :, hs, ws].double()
            if p.res_scale:                                   # folded GroupNorm-apply + SiLU of the residual branch
                a = _view(p.res_scale, B * p.Cout).reshape(B, 1, 1, 1, p.Cout).double()
                d = _view(p.res_shift, B * p.Cout).reshape(B, 1, 1, 1, p.Cout).double()
                r = torch.nn.functional.silu(r * a + d)
            acc += r
        y[:, :, hs, ws] = acc.float()
    else:
        y = _view(p.y, B * Fo * Hf * Wf * p.Cout).reshape(B, Fo, p.Cout, Hf, Wf)
        y[:, :, :, hs, ws] = acc.float().permute(0, 1, 4, 2, 3)
    if p.gn_stats:
        G = p.gn_groups
        st = _view(p.gn_stats, B * G * 2, np.float64).reshape(B, G, 2)
        v = acc.float().double().reshape(B, -1, G, p.Cout // G)
        st[:, :, 0] += v.sum(dim=(1, 3))
        st[:, :, 1] += (v * v).sum(dim=(1, 3))
    return ran_tc


def groupnorm_silu(y, stats, gamma, beta, scale_shift, ss_stride, ss_off, residual, out, B, rps, Cn, groups, eps=1e-5):
    v = y[: B * rps * Cn].reshape(B, rps, groups, Cn // groups)
    st = stats[: B * groups * 2].reshape(B, groups, 2)
    n = rps * (Cn // groups)
    mean = st[:, :, 0] / n
    var = (st[:, :, 1] / n - mean * mean).clamp(min=0)
    rstd = (1.0 / torch.sqrt(var + eps)).float()
    t = (v - mean.float()[:, None, :, None]) * rstd[:, None, :, None]
    t = t.reshape(B, rps, Cn) * gamma + beta
    if scale_shift is not None:
        ss = scale_shift[: B * ss_stride].reshape(B, ss_stride)
        t = t * (ss[:, None, ss_off:ss_off + Cn] + 1) + ss[:, None, ss_off + Cn:ss_off + 2 * Cn]
    t = torch.nn.functional.silu(t)
    if residual is not None:
        t = t + residual[: B * rps * Cn].reshape(B, rps, Cn)
    out[: B * rps * Cn] = t.reshape(-1)


def gn_stats_merge(stats_in, stats_out, B, groups_in, groups_out):
    a = stats_in[: B * groups_in * 2].reshape(B, groups_out, groups_in // groups_out, 2)
    stats_out[: B * groups_out * 2] += a.sum(dim=2).reshape(-1)


def gn_fold(stats, gamma, beta, scale, shift, B, rps, Cn, groups, eps=1e-5):
    st = stats[: B * groups * 2].reshape(B, groups, 2).double()
    n = rps * (Cn // groups)
    mean = st[:, :, 0] / n
    var = (st[:, :, 1] / n - mean * mean).clamp(min=0)
    rstd = (1.0 / (var + eps).sqrt()).repeat_interleave(Cn // groups, dim=1)
    mean = mean.repeat_interleave(Cn // groups, dim=1)
    a = rstd * gamma.double()[None, :]
    scale[: B * Cn] = a.float().reshape(-1)
    shift[: B * Cn] = (beta.double()[None, :] - mean * a).float().reshape(-1)


def layernorm_channels(x, gamma, out, rows, Cn, eps=1e-5, residual=None, use_rsqrt=False):
    v = x[: rows * Cn].reshape(rows, Cn)
    mean = v.mean(dim=1, keepdim=True)
    var = v.var(dim=1, unbiased=False, keepdim=True)
    o = (v - mean) * (var + eps).rsqrt() * gamma if use_rsqrt else (v - mean) / (var + eps).sqrt() * gamma
    if residual is not None:
        o = o + residual[: rows * Cn].reshape(rows, Cn)
    out[: rows * Cn] = o.reshape(-1)


def upsample_nearest2x(x, out, BF, H, W, Cn):
    v = x[: BF * H * W * Cn].reshape(BF, H, W, Cn)
    o = v.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    out[: o.numel()] = o.reshape(-1)


def pack_input(x, out, B, F, Ctot, c0, Cin, H, W, Cpad):
    v = x.reshape(B, F, Ctot, H, W)[:, :, c0:c0 + Cin].permute(0, 1, 3, 4, 2)
    o = torch.zeros(B, F, H, W, Cpad)
    o[..., :Cin] = v
    out[: o.numel()] = o.reshape(-1)


def _split_heads(qkv, heads):
    hid = heads * HEADS_DIM
    q, k, v = qkv[..., :hid], qkv[..., hid:2 * hid], qkv[..., 2 * hid:]
    f = lambda t: t.reshape(*t.shape[:-1], heads, HEADS_DIM).transpose(-2, -3)  # ... h n d
    return f(q), f(k), f(v)


def temporal_attention(qkv, rope_cos, rope_sin, pos_bias, out, B, F, HW, heads, use_rope=True, precise=False, relative_bias=False):
    hid = heads * HEADS_DIM
    t = qkv[: B * F * HW * 3 * hid].reshape(B, F, HW, 3 * hid).permute(0, 2, 1, 3)  # b hw f c
    q, k, v = _split_heads(t, heads)
    q = q * (HEADS_DIM ** -0.5)
    if use_rope:
        def rot(u):
            x = u.reshape(*u.shape[:-1], HEADS_DIM // 2, 2)
            x1, x2 = x.unbind(-1)
            r = torch.stack((-x2, x1), dim=-1).reshape(u.shape)
            return u * rope_cos + r * rope_sin
        q, k = rot(q), rot(k)
    sim = torch.einsum("...hid,...hjd->...hij", q, k)
    if pos_bias is not None:
        sim = sim + pos_bias
    o = torch.einsum("...hij,...hjd->...hid", sim.softmax(-1), v)
    o = o.transpose(-2, -3).reshape(B, HW, F, hid).permute(0, 2, 1, 3)
    out[: o.numel()] = o.reshape(-1)


def temporal_block_fused(x, w_qkv, w_out, rope_cos, rope_sin, pos_bias, y, B, F, HW, Cn, heads, eps=1e-5):
    if F < 1 or F > 32 or Cn != 64 or heads != 4 or HW % 4:
        return False
    rows = B * F * HW
    v = x[: rows * Cn].reshape(rows, Cn)
    xh = (v - v.mean(1, keepdim=True)) / (v.var(1, unbiased=False, keepdim=True) + eps).sqrt()
    qkv = (xh.double() @ w_qkv.double().t()).float().reshape(-1)
    att = torch.empty(rows * heads * HEADS_DIM)
    temporal_attention(qkv, rope_cos, rope_sin, pos_bias, att, B, F, HW, heads)
    o = (att.reshape(rows, -1).double() @ w_out.double().t()).float() + v
    y[: rows * Cn] = o.reshape(-1)
    return True


def spatial_attention(qkv, out, BF, HW, heads, precise=True):
    hid = heads * HEADS_DIM
    t = qkv[: BF * HW * 3 * hid].reshape(BF, HW, 3 * hid)
    q, k, v = _split_heads(t, heads)
    sim = torch.einsum("bhid,bhjd->bhij", q * (HEADS_DIM ** -0.5), k)
    o = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(BF, HW, hid)
    out[: o.numel()] = o.reshape(-1)


def spatial_linear_attention(qkv, ctx_ws, out, BF, HW, heads):
    hid = heads * HEADS_DIM
    t = qkv[: BF * HW * 3 * hid].reshape(BF, HW, 3 * hid)
    q, k, v = _split_heads(t, heads)  # b h n d
    q = q.softmax(dim=-1) * (HEADS_DIM ** -0.5)
    k = k.softmax(dim=-2)
    ctx = torch.einsum("bhnd,bhne->bhde", k, v)
    o = torch.einsum("bhde,bhnd->bhne", ctx, q).transpose(1, 2).reshape(BF, HW, hid)
    out[: o.numel()] = o.reshape(-1)


def stem_conv(x, w, bias, y, B, F, H, W, Cpad, N, kt, kh, kw):
    """dpc_stem_conv_tcgen05: unpacks the [N][kt*kh*P*32] sliding-window weight layout (packing.pack_stem_conv) and runs the
    convolution in fp64."""
    if Cpad % 4 or Cpad > 16 or N not in (32, 64, 128) or kw != 7 or W < 8:
        return False
    P = Cpad // 4
    wt = w[: N * kt * kh * P * 32].reshape(N, kt, kh, P, 8, 4).permute(0, 3, 5, 1, 2, 4).reshape(N, Cpad, kt, kh, 8)[..., :7]
    xin = x[: B * F * H * W * Cpad].reshape(B, F, H, W, Cpad).permute(0, 4, 1, 2, 3)
    o = torch.nn.functional.conv3d(xin.double(), wt.double(), bias.double(), padding=(kt // 2, kh // 2, 3))
    y[: B * F * H * W * N] = o.permute(0, 2, 3, 4, 1).float().reshape(-1)
    return True


def final_proj(x, w, bias, out, BF, HW, Cn, Cout):
    if Cn != 64 or Cout not in (2, 4, 6):
        return False
    o = x[: BF * HW * Cn].reshape(BF, HW, Cn).double() @ w.double().t()
    if bias is not None:
        o = o + bias.double()
    out.reshape(-1)[: BF * Cout * HW] = o.permute(0, 2, 1).float().reshape(-1)
    return True


def spatial_linear_block_fused(x, w_qkv, w_out, b_out, ctx_ws, mt_ws, y, BF, HW, Cn, heads, eps=1e-5):
    if Cn != 64 or heads != 4 or HW % 128:
        return False
    rows = BF * HW
    v = x[: rows * Cn].reshape(rows, Cn)
    xh = (v - v.mean(1, keepdim=True)) / (v.var(1, unbiased=False, keepdim=True) + eps).sqrt()
    qkv = (xh.double() @ w_qkv.double().t()).float().reshape(-1)
    att = torch.empty(rows * heads * HEADS_DIM)
    spatial_linear_attention(qkv, ctx_ws, att, BF, HW, heads)
    o = (att.reshape(rows, -1).double() @ w_out.double().t()).float() + v
    if b_out is not None:
        o = o + b_out[None, :]
    y[: rows * Cn] = o.reshape(-1)
    return True


def time_embed(t, freqs, w1, b1, w2, b2, hidden_ws, t_emb, B, dim):
    arg = t.float()[:, None] * freqs[None, :]
    emb = torch.cat((arg.sin(), arg.cos()), dim=-1)
    h = torch.nn.functional.gelu(emb @ w1.t() + b1)
    t_emb[: B * dim * 4] = (h @ w2.t() + b2).reshape(-1)


def time_proj(t_emb, W, bias, out, B, tdim, total):
    e = torch.nn.functional.silu(t_emb[: B * tdim].reshape(B, tdim))
    out[: B * total] = (e @ W.t() + bias).reshape(-1)


def renoise(x, z, a, b, out):
    v = torch.tensor(a, dtype=torch.float32) * x
    out.copy_(v + (torch.tensor(b, dtype=torch.float32) * z if z is not None else 0.0))


def predict_x_start(x, eps, sr, srm1, clip, out):
    v = np.float32(sr) * x - np.float32(srm1) * eps
    out.copy_(v.clamp(-1, 1) if clip else v)


def guided_step(ddim, x, eps_joint, eps_w, noise, init, g, c, x_out, x_start_out, B, F, H, W):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    sr, srm1 = f32(c.sqrt_recip_alphas_cumprod), f32(c.sqrt_recipm1_alphas_cumprod)
    clip = (lambda v: v.clamp(-1, 1)) if ddim else (lambda v: v)
    ew = torch.zeros_like(eps_joint)
    ew[:, :, 3:5] = eps_w
    xs0 = clip(sr * x - srm1 * eps_joint)
    if g is None:
        R = torch.tensor(list(c.rescaler), dtype=torch.float32).reshape(1, 1, 6, 1, 1)
        g = torch.zeros_like(x)
        g[:, -1, 5] = -(1.0 / (H * W))
        g[:, :, 3:5] = (f32(c.w_energy) / f32(float(F * 2 * H * W))) * (2.0 * (xs0 * R)[:, :, 3:5])
    pn = eps_joint + (f32(c.guidance_coef) * g + f32(c.prior_coef) * ew)
    xs = clip(sr * x - srm1 * pn)
    if ddim:
        pn = (sr * x - xs) / srm1
        if c.last:
            o = xs
        else:
            o = xs * f32(c.sqrt_alpha_next) + f32(c.c) * pn + f32(c.ddim_sigma) * noise
    else:
        xs = xs.clamp(-1, 1)
        o = f32(c.posterior_mean_coef1) * xs + f32(c.posterior_mean_coef2) * x
        if c.add_noise:
            o = o + f32(c.sigma) * noise
    if init is not None and not (ddim and c.last):
        o = o.clone()
        o[:, 0, 0] = init
    x_out.copy_(o)
    if x_start_out is not None:
        x_start_out.copy_(xs)


def burgers_model_output(x, eps1, eps2, out, x_start, mode, coef, beta, sr, srm1, Cn, plane):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    if mode == 2:
        o = f32(beta) * eps1
        o[:, 0] = 0
    else:
        e2 = eps2.clone()
        e2[:, 0] = 0
        o = eps1 - f32(coef) * e2
        if mode == 1:
            o = o / f32(beta)
    out.copy_(o)
    if x_start is not None:
        x_start.copy_(f32(sr) * x - f32(srm1) * o)


def ddpm_posterior_step(x, eps, g, noise, x_out, x_start_out, pred_noise_out, gscale, sr, srm1, clip, c1, c2, sigma):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    pn = eps if g is None else eps + g * f32(gscale)
    xs = f32(sr) * x - f32(srm1) * pn
    if clip:
        xs = xs.clamp(-1, 1)
    o = f32(c1) * xs + f32(c2) * x
    if noise is not None:
        o = o + f32(sigma) * noise
    x_out.copy_(o)
    if x_start_out is not None:
        x_start_out.copy_(xs)
    if pred_noise_out is not None:
        pred_noise_out.copy_(pn)


def jelly_x_start(x, eps, x_start, sr, srm1, clip):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    x4 = torch.cat([x[:, :, :3], x[:, :, 6:]], dim=2)
    xs = f32(sr) * x4 - f32(srm1) * eps
    x_start.copy_(xs.clamp(-1, 1) if clip else xs)


def jelly_step(x, x_start, eps, eps_w, g, noise, state_0, thetas_0, x_next, x_w, dtheta, theta_mean, ga, gb, c1, c2, sigma,
               ddim, cond_steps):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    x4 = torch.cat([x[:, :, :3], x[:, :, 6:]], dim=2)
    if ddim:
        pn = eps
        if g is not None:
            pad = torch.zeros_like(eps)
            pad[:, :, 3:] = eps_w
            pn = eps + (f32(ga) * g - f32(gb) * pad)
        pred = x_start * f32(c1) + f32(c2) * pn
        if noise is not None:
            pred = pred + f32(sigma) * noise
    else:
        pred = f32(c1) * x_start + f32(c2) * x4
        if noise is not None:
            pred = pred + f32(sigma) * noise
        if g is not None:
            pred = pred - (f32(ga) * g - f32(gb) * eps_w)
    cs, th = cond_steps, thetas_0.reshape(-1, 1, 1, 1).expand(-1, 1, *x.shape[-2:])
    theta = pred[:, :, 3].clone()
    dtheta.copy_(theta.mean((-1, -2)) - thetas_0[:, None])
    states = pred[:, :, :3].clone()
    if cs > 0:                                          # `[:, -0:]` would be the whole tensor
        states[:, :cs] = state_0.unsqueeze(1)
        theta[:, :cs] = th
        theta[:, -cs:] = th
    theta_mean.copy_(theta.mean((-1, -2)))
    x_next[:, :, :3] = states
    x_next[:, :, 6] = theta
    x_w[:, :, 6] = theta


def jelly_write_bd(pred_bd, bd_0, x_next, x_w, cond_steps):
    bd = pred_bd.reshape(x_next.shape[0], x_next.shape[1], 3, *x_next.shape[-2:]).clone()
    if cond_steps > 0:
        bd[:, :cond_steps] = bd_0.unsqueeze(1)
        bd[:, -cond_steps:] = bd_0.unsqueeze(1)
    x_next[:, :, 3:6] = bd
    x_w[:, :, 3:6] = bd



# ---- jellyfish surrogate networks (csrc/nets2d.cu): forward variants and BACKWARD kernels, emulated with torch.autograd ----
def spatial_linear_attention_ex(qkv, ctx_ws, kstat, out, BF, HW, heads, v_scale):
    hid = heads * HEADS_DIM
    t = qkv[: BF * HW * 3 * hid].reshape(BF, HW, 3 * hid)
    q, k, v = _split_heads(t, heads)  # b h n d
    qs = q.softmax(dim=-1) * (HEADS_DIM ** -0.5)
    kmax = k.max(dim=-2).values
    ksum = (k - kmax[:, :, None, :]).exp().sum(dim=-2)
    ks = k.softmax(dim=-2)
    ctx = torch.einsum("bhnd,bhne->bhde", ks, v * v_scale)
    o = torch.einsum("bhde,bhnd->bhne", ctx, qs).transpose(1, 2).reshape(BF, HW, hid)
    out[: o.numel()] = o.reshape(-1)
    ctx_ws[: ctx.numel()] = ctx.reshape(-1)
    if kstat is not None:
        kstat[: BF * heads * HEADS_DIM * 2] = torch.stack([kmax, ksum], dim=-1).reshape(-1)


def linattn2d_bwd(qkv, ctx, kstat, dout, dctx_ws, dqkv, BF, HW, heads, v_scale):
    hid = heads * HEADS_DIM
    with torch.enable_grad():
        t = qkv[: BF * HW * 3 * hid].reshape(BF, HW, 3 * hid).clone().requires_grad_()
        q, k, v = _split_heads(t, heads)
        qs = q.softmax(dim=-1) * (HEADS_DIM ** -0.5)
        ks = k.softmax(dim=-2)
        c = torch.einsum("bhnd,bhne->bhde", ks, v * v_scale)
        o = torch.einsum("bhde,bhnd->bhne", c, qs).transpose(1, 2).reshape(BF, HW, hid)
        (g,) = torch.autograd.grad(o, t, dout[: BF * HW * hid].reshape(BF, HW, hid))
    dqkv[: g.numel()] = g.reshape(-1)


def attention2d_bwd(qkv, out, dout, dqkv, BF, HW, heads):
    hid = heads * HEADS_DIM
    with torch.enable_grad():
        t = qkv[: BF * HW * 3 * hid].reshape(BF, HW, 3 * hid).clone().requires_grad_()
        q, k, v = _split_heads(t, heads)
        sim = torch.einsum("bhid,bhjd->bhij", q * (HEADS_DIM ** -0.5), k)
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(BF, HW, hid)
        (g,) = torch.autograd.grad(o, t, dout[: BF * HW * hid].reshape(BF, HW, hid))
    dqkv[: g.numel()] = g.reshape(-1)


def gn_silu_bwd(y, stats, gamma, beta, scale_shift, ss_stride, ss_off, dout, dy, sums_ws, dss, B, rps, Cn, groups, eps=1e-5):
    with torch.enable_grad():
        v = y[: B * rps * Cn].reshape(B, rps, Cn).clone().requires_grad_()
        t = torch.nn.functional.group_norm(v.permute(0, 2, 1), groups, gamma, beta, eps=eps).permute(0, 2, 1)
        ins = [v]
        if scale_shift is not None:
            ss = scale_shift[: B * ss_stride].reshape(B, ss_stride).clone().requires_grad_()
            t = t * (ss[:, None, ss_off:ss_off + Cn] + 1) + ss[:, None, ss_off + Cn:ss_off + 2 * Cn]
            ins.append(ss)
        o = torch.nn.functional.silu(t)
        gs = torch.autograd.grad(o, ins, dout[: B * rps * Cn].reshape(B, rps, Cn))
    dy[: B * rps * Cn] = gs[0].reshape(-1)
    if dss is not None:
        d = dss[: B * ss_stride].reshape(B, ss_stride)
        d[:, ss_off:ss_off + 2 * Cn] = gs[1][:, ss_off:ss_off + 2 * Cn]


def layernorm_channels_bwd(x, gamma, dy, add, dx, rows, Cn, eps=1e-5, use_rsqrt=True):
    with torch.enable_grad():
        v = x[: rows * Cn].reshape(rows, Cn).clone().requires_grad_()
        mean = v.mean(dim=1, keepdim=True)
        var = v.var(dim=1, unbiased=False, keepdim=True)
        o = (v - mean) * (var + eps).rsqrt() * gamma
        (g,) = torch.autograd.grad(o, v, dy[: rows * Cn].reshape(rows, Cn))
    if add is not None:
        g = g + add[: rows * Cn].reshape(rows, Cn)
    dx[: rows * Cn] = g.reshape(-1)


def add(a, b, out, n):
    out[:n] = a[:n] + b[:n]


def sumpool2x2(dy, dx, N, H, W, Cn):
    v = dy[: N * 4 * H * W * Cn].reshape(N, H, 2, W, 2, Cn).sum((2, 4))
    dx[: v.numel()] = v.reshape(-1)


def mean_head(x, W, bias, out, N, HW, Cn, O):
    v = x[: N * HW * Cn].reshape(N, HW, Cn).mean(1)
    out.copy_(v @ W.t() + bias)


def mean_head_bwd(dout, W, dx, N, HW, Cn, O):
    g = ((dout @ W) / HW)[:, None, :].expand(N, HW, Cn)
    dx[: N * HW * Cn] = g.reshape(-1)


def time_embed_f32(t, freqs, w1, b1, w2, b2, hidden_ws, t_emb, B, dim):
    time_embed(t, freqs, w1, b1, w2, b2, hidden_ws, t_emb, B, dim)


def time_mlp_bwd(t, freqs, w1, b1, w2, w_proj, t_emb, dss, dt, B, dim, total):
    with torch.enable_grad():
        tr = t.clone().requires_grad_()
        arg = tr[:, None] * freqs[None, :]
        emb = torch.cat((arg.sin(), arg.cos()), dim=-1)
        te = torch.nn.functional.gelu(emb @ w1.t() + b1) @ w2.t()
        ss = torch.nn.functional.silu(te + (t_emb[: B * dim * 4].reshape(B, dim * 4) - te).detach()) @ w_proj.t()
        (g,) = torch.autograd.grad(ss, tr, dss[: B * total].reshape(B, total))
    dt.copy_(g)


EMULATED = ("temporal_block_fused", "gn_fold", "gn_stats_merge", "renoise", "final_proj", "spatial_linear_block_fused", "stem_conv", "jelly_x_start", "jelly_step", "jelly_write_bd", "burgers_model_output", "ddpm_posterior_step", "conv", "groupnorm_silu", "layernorm_channels", "pack_input", "temporal_attention", "spatial_attention",
            "spatial_linear_attention", "time_embed", "time_proj", "predict_x_start", "guided_step", "upsample_nearest2x",
            "spatial_linear_attention_ex", "linattn2d_bwd", "attention2d_bwd", "gn_silu_bwd", "layernorm_channels_bwd", "add", "sumpool2x2",
            "mean_head", "mean_head_bwd", "time_embed_f32", "time_mlp_bwd")


def install_raw():
    """Same as install() without pytest (used inside spawned worker processes)."""
    from diffphycon_b200 import _lib, unet3d
    for name in EMULATED:
        setattr(_lib, name, globals()[name])
    _lib.stream_ptr = lambda: None
    unet3d._require_cuda = lambda x: None


def install(monkeypatch):
    """Route diffphycon_b200._lib's launch wrappers to the emulators above (tests only)."""
    from diffphycon_b200 import _lib
    for name in EMULATED:
        monkeypatch.setattr(_lib, name, globals()[name])
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)


_ = uo  # the oracle is imported so that a missing oracle fails loudly at collection time
