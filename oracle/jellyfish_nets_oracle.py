"""CPU restatement (plain PyTorch fp32, functional, differentiable) of the jellyfish surrogate networks and of `force_fn`.

TEST INFRASTRUCTURE ONLY — never imported by the product path (diffphycon_b200/).  Allowed importers: tests/,
__graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.

Follows /root/reference/diffusion/diffusion_2d_jellyfish.py (cited as jf.py:line): `Unet` :276-403, `ForceUnet` :406-481 and
their blocks :86-255; and /root/reference/inference/inference_2d_jellyfish.py:35-36, :47-60, :85-114 (`unnormalize_state`,
`reg_theta`, `force_fn`).  Pinned: tests/golden/make_golden_jellyfish_nets.py runs the UNMODIFIED reference classes (and the
reference's own `force_fn` source, lifted from the file with ast because the module unpickles a data file at import) on
seeded inputs; tests/test_jellyfish_nets.py checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import param_gen

HEADS, DIM_HEAD = 4, 32
HIDDEN = HEADS * DIM_HEAD


# ---- parameter inventory (state_dict keys / shapes of the reference constructors) ------------------------------------------
def _resnet_shapes(pre, cin, cout, time_dim, out):
    if time_dim is not None:
        out[f"{pre}.mlp.1.weight"] = (cout * 2, time_dim)
        out[f"{pre}.mlp.1.bias"] = (cout * 2,)
    for blk, ci in (("block1", cin), ("block2", cout)):
        out[f"{pre}.{blk}.proj.weight"] = (cout, ci, 3, 3)
        out[f"{pre}.{blk}.proj.bias"] = (cout,)
        out[f"{pre}.{blk}.norm.weight"] = (cout,)
        out[f"{pre}.{blk}.norm.bias"] = (cout,)
    if cin != cout:
        out[f"{pre}.res_conv.weight"] = (cout, cin, 1, 1)
        out[f"{pre}.res_conv.bias"] = (cout,)


def _linattn_shapes(pre, c, out):
    out[f"{pre}.fn.fn.to_qkv.weight"] = (HIDDEN * 3, c, 1, 1)
    out[f"{pre}.fn.fn.to_out.0.weight"] = (c, HIDDEN, 1, 1)
    out[f"{pre}.fn.fn.to_out.0.bias"] = (c,)
    out[f"{pre}.fn.fn.to_out.1.g"] = (1, c, 1, 1)
    out[f"{pre}.fn.norm.g"] = (1, c, 1, 1)


def param_shapes(kind: str, dim=64, dim_mults=(1, 2, 4, 8), channels=3, out_dim=None) -> "OrderedDict[str, tuple]":
    """kind = "unet" (jf.py:292-363) or "force" (jf.py:418-460)."""
    assert kind in ("unet", "force")
    out: "OrderedDict[str, tuple]" = OrderedDict()
    dims = [dim, *[dim * m for m in dim_mults]]
    in_out = list(zip(dims[:-1], dims[1:]))
    tdim = dim * 4 if kind == "unet" else None
    out["init_conv.weight"] = (dim, channels, 7, 7)
    out["init_conv.bias"] = (dim,)
    if kind == "unet":
        out["time_mlp.1.weight"], out["time_mlp.1.bias"] = (tdim, dim), (tdim,)
        out["time_mlp.3.weight"], out["time_mlp.3.bias"] = (tdim, tdim), (tdim,)
    n = len(in_out)
    for i, (di, do) in enumerate(in_out):
        _resnet_shapes(f"downs.{i}.0", di, di, tdim, out)
        _resnet_shapes(f"downs.{i}.1", di, di, tdim, out)
        _linattn_shapes(f"downs.{i}.2", di, out)
        if i < n - 1:
            out[f"downs.{i}.3.1.weight"], out[f"downs.{i}.3.1.bias"] = (do, di * 4, 1, 1), (do,)
        else:
            out[f"downs.{i}.3.weight"], out[f"downs.{i}.3.bias"] = (do, di, 3, 3), (do,)
    mid = dims[-1]
    _resnet_shapes("mid_block1", mid, mid, tdim, out)
    out["mid_attn.fn.fn.to_qkv.weight"] = (HIDDEN * 3, mid, 1, 1)
    out["mid_attn.fn.fn.to_out.weight"], out["mid_attn.fn.fn.to_out.bias"] = (mid, HIDDEN, 1, 1), (mid,)
    out["mid_attn.fn.norm.g"] = (1, mid, 1, 1)
    _resnet_shapes("mid_block2", mid, mid, tdim, out)
    if kind == "unet":
        for i, (di, do) in enumerate(reversed(in_out)):
            _resnet_shapes(f"ups.{i}.0", do + di, do, tdim, out)
            _resnet_shapes(f"ups.{i}.1", do + di, do, tdim, out)
            _linattn_shapes(f"ups.{i}.2", do, out)
            if i < n - 1:
                out[f"ups.{i}.3.1.weight"], out[f"ups.{i}.3.1.bias"] = (di, do, 3, 3), (di,)
            else:
                out[f"ups.{i}.3.weight"], out[f"ups.{i}.3.bias"] = (di, do, 3, 3), (di,)
        _resnet_shapes("final_res_block", dim * 2, dim, tdim, out)
        od = channels if out_dim is None else out_dim
        out["final_conv.weight"], out["final_conv.bias"] = (od, dim, 1, 1), (od,)
    else:
        out["final.weight"], out["final.bias"] = (out_dim, 512), (out_dim,)
    return out


def make_params(kind, seed=0, **kw):
    return param_gen.make_params(param_shapes(kind, **kw), seed)


# ---- blocks -----------------------------------------------------------------------------------------------------------------
def ws_conv2d(x, w, b):
    """WeightStandardizedConv2d.forward, fp32 branch (jf.py:113-121)."""
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return F.conv2d(x, (w - mean) * (var + 1e-5).rsqrt(), b, padding=1)


def layer_norm(x, g):
    """LayerNorm.forward (jf.py:128-132)."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def block(p, pre, x, groups, scale_shift=None):
    """Block.forward (jf.py:196-204)."""
    x = ws_conv2d(x, p[f"{pre}.proj.weight"], p[f"{pre}.proj.bias"])
    x = F.group_norm(x, groups, p[f"{pre}.norm.weight"], p[f"{pre}.norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(p, pre, x, t_emb, groups):
    """ResnetBlock.forward (jf.py:217-230)."""
    scale_shift = None
    if t_emb is not None and f"{pre}.mlp.1.weight" in p:
        te = F.linear(F.silu(t_emb), p[f"{pre}.mlp.1.weight"], p[f"{pre}.mlp.1.bias"])[:, :, None, None]
        scale_shift = te.chunk(2, dim=1)
    h = block(p, f"{pre}.block1", x, groups, scale_shift)
    h = block(p, f"{pre}.block2", h, groups)
    res = F.conv2d(x, p[f"{pre}.res_conv.weight"], p[f"{pre}.res_conv.bias"]) if f"{pre}.res_conv.weight" in p else x
    return h + res


def linear_attention_block(p, pre, x):
    """Residual(PreNorm(LinearAttention)) (jf.py:86-92, :134-142, :206-225)."""
    b, c, h, w = x.shape
    xn = layer_norm(x, p[f"{pre}.fn.norm.g"])
    qkv = F.conv2d(xn, p[f"{pre}.fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = [t.reshape(b, HEADS, DIM_HEAD, h * w) for t in qkv]
    q = q.softmax(dim=-2) * DIM_HEAD ** -0.5
    k = k.softmax(dim=-1)
    v = v / (h * w)
    context = torch.einsum('bhdn,bhen->bhde', k, v)
    out = torch.einsum('bhde,bhdn->bhen', context, q).reshape(b, HIDDEN, h, w)
    out = F.conv2d(out, p[f"{pre}.fn.fn.to_out.0.weight"], p[f"{pre}.fn.fn.to_out.0.bias"])
    return layer_norm(out, p[f"{pre}.fn.fn.to_out.1.g"]) + x


def attention_block(p, pre, x):
    """Residual(PreNorm(Attention)) (jf.py:227-255)."""
    b, c, h, w = x.shape
    xn = layer_norm(x, p[f"{pre}.fn.norm.g"])
    qkv = F.conv2d(xn, p[f"{pre}.fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = [t.reshape(b, HEADS, DIM_HEAD, h * w) for t in qkv]
    q = q * DIM_HEAD ** -0.5
    sim = torch.einsum('bhdi,bhdj->bhij', q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bhij,bhdj->bhid', attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, HIDDEN, h, w)
    return F.conv2d(out, p[f"{pre}.fn.fn.to_out.weight"], p[f"{pre}.fn.fn.to_out.bias"]) + x


def sinusoidal_pos_emb(x, dim):
    """SinusoidalPosEmb.forward (jf.py:148-155)."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=x.device) * -emb)
    emb = x[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def pixel_unshuffle_conv(x, w, b):
    """Downsample (jf.py:100-104): 'b c (h p1) (w p2) -> b (c p1 p2) h w' then a 1x1 conv."""
    bsz, c, h, w_ = x.shape
    x = x.reshape(bsz, c, h // 2, 2, w_ // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(bsz, c * 4, h // 2, w_ // 2)
    return F.conv2d(x, w, b)


def _down_path(p, x, t, n_lvl, groups, hs: Optional[list]):
    for i in range(n_lvl):
        x = resnet_block(p, f"downs.{i}.0", x, t, groups)
        if hs is not None:
            hs.append(x)
        x = resnet_block(p, f"downs.{i}.1", x, t, groups)
        x = linear_attention_block(p, f"downs.{i}.2", x)
        if hs is not None:
            hs.append(x)
        if i < n_lvl - 1:
            x = pixel_unshuffle_conv(x, p[f"downs.{i}.3.1.weight"], p[f"downs.{i}.3.1.bias"])
        else:
            x = F.conv2d(x, p[f"downs.{i}.3.weight"], p[f"downs.{i}.3.bias"], padding=1)
    x = resnet_block(p, "mid_block1", x, t, groups)
    x = attention_block(p, "mid_attn", x)
    return resnet_block(p, "mid_block2", x, t, groups)


def unet_forward(p: Dict[str, torch.Tensor], x, time, dim=64, n_lvl=4, groups=8):
    """Unet.forward (jf.py:365-403).  x [N,C,H,W], time [N] float."""
    x = F.conv2d(x, p["init_conv.weight"], p["init_conv.bias"], padding=3)
    r = x.clone()
    t = sinusoidal_pos_emb(time, dim)
    t = F.linear(t, p["time_mlp.1.weight"], p["time_mlp.1.bias"])
    t = F.gelu(t)
    t = F.linear(t, p["time_mlp.3.weight"], p["time_mlp.3.bias"])
    hs = []
    x = _down_path(p, x, t, n_lvl, groups, hs)
    for i in range(n_lvl):
        x = torch.cat((x, hs.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.0", x, t, groups)
        x = torch.cat((x, hs.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.1", x, t, groups)
        x = linear_attention_block(p, f"ups.{i}.2", x)
        if i < n_lvl - 1:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, p[f"ups.{i}.3.1.weight"], p[f"ups.{i}.3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, p[f"ups.{i}.3.weight"], p[f"ups.{i}.3.bias"], padding=1)
    x = torch.cat((x, r), dim=1)
    x = resnet_block(p, "final_res_block", x, t, groups)
    return F.conv2d(x, p["final_conv.weight"], p["final_conv.bias"])


def force_forward(p: Dict[str, torch.Tensor], x, n_lvl=4, groups=8):
    """ForceUnet.forward (jf.py:462-481).  x [N,C,H,W] -> [N,out_dim]."""
    x = F.conv2d(x, p["init_conv.weight"], p["init_conv.bias"], padding=3)
    x = _down_path(p, x, None, n_lvl, groups, None)
    x = x.mean(dim=-1).mean(dim=-1)
    return F.linear(x, p["final.weight"], p["final.bias"])


# ---- guidance (inference/inference_2d_jellyfish.py) ------------------------------------------------------------------------
def reg_theta(theta):
    """:47-60."""
    d = theta[:, 1:] - theta[:, :-1]
    return torch.sum(d * d, dim=1)


def force_fn(x, bd_0, force_params, bd_params, p_min, p_max, reg_ratio, only_vis_pressure=False, dim=64, n_lvl=4
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """:85-114: returns (grad_state, grad_theta_expand) by autograd, exactly as the reference does."""
    with torch.enable_grad():
        if only_vis_pressure:
            state, theta_expand = x[:, :, :1], x[:, :, -1]
        else:
            state, theta_expand = x[:, :, :3], x[:, :, 3]
        state = state.detach().clone().requires_grad_()
        theta_expand = theta_expand.detach().clone().requires_grad_()
        theta = torch.mean(torch.mean(theta_expand, dim=3), dim=2)
        pressure = state[:, :, 0] if only_vis_pressure else state[:, :, 2]
        pressure = (0.5 * pressure + 0.5) * (p_max - p_min) + p_min                                  # :35-36
        B, Fr = bd_0.shape[:2]
        pred_bd = unet_forward(bd_params, bd_0.reshape(B * Fr, *bd_0.shape[2:]), theta.reshape(B * Fr), dim, n_lvl)
        pred_bd = pred_bd.reshape(bd_0.shape)
        inp = torch.cat((pressure.unsqueeze(2), pred_bd), dim=2)
        inp = inp.reshape(B * Fr, *inp.shape[2:])
        force = force_forward(force_params, inp, n_lvl).reshape(B, Fr)
        weight = torch.arange(Fr, 0, -1, dtype=torch.float32).expand(B, Fr)
        average_velocity = torch.mean(force * weight, dim=1)
        guidance = -average_velocity + reg_ratio * reg_theta(theta)
        gs, gt = torch.autograd.grad(guidance, [state, theta_expand], grad_outputs=torch.ones_like(guidance))
    return gs, gt


def design_fn(x, bd_0, **kw):
    """:276-279."""
    gs, gt = force_fn(x, bd_0, **kw)
    return torch.cat([gs, gt.unsqueeze(2)], dim=2)
