"""CPU restatement (plain PyTorch fp32, functional) of the reference `Unet3D_with_Conv3D` forward.

TEST INFRASTRUCTURE ONLY — never imported by the product path (diffphycon_b200/).  Allowed importers: tests/,
__graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.

Follows /root/reference/model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py (cited per function as
`conv3d.py:line`).  Pinned against the reference itself: tests/golden/make_golden.py runs the UNMODIFIED reference
module (imported through oracle/shims) on seeded inputs and commits input/output vectors under tests/golden/;
tests/test_oracle_golden.py checks this restatement against them.  The one piece that cannot be pinned is RoPE
(`rotary-embedding-torch==0.8.4`, not vendored in the reference): restated from its published algorithm in
oracle/shims/rotary_embedding_torch — "parity unpinned" for that function only.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

HEADS = 4
DIM_HEAD = 32
HIDDEN = HEADS * DIM_HEAD  # 128, conv3d.py:237, :287


@dataclass(frozen=True)
class UnetCfg:
    """Constructor arguments of Unet3D_with_Conv3D (conv3d.py:357-372) that the hot path uses."""
    dim: int = 64
    dim_mults: Tuple[int, ...] = (1, 2, 4)
    channels: int = 6
    out_dim: Optional[int] = None
    init_kernel_size: int = 7
    resnet_groups: int = 8

    @property
    def out_channels(self) -> int:
        return self.channels if self.out_dim is None else self.out_dim

    @property
    def dims(self):
        return [self.dim] + [self.dim * m for m in self.dim_mults]

    @property
    def in_out(self):
        d = self.dims
        return list(zip(d[:-1], d[1:]))

    @property
    def time_dim(self) -> int:
        return self.dim * 4


# --------------------------------------------------------------------------------------------------------------
# parameter inventory (state_dict keys / shapes of the reference constructor, conv3d.py:373-471)
# --------------------------------------------------------------------------------------------------------------

def _resnet_shapes(prefix, cin, cout, time_dim, out):
    if time_dim is not None:
        out[f"{prefix}.mlp.1.weight"] = (cout * 2, time_dim)
        out[f"{prefix}.mlp.1.bias"] = (cout * 2,)
    for blk, ci in (("block1", cin), ("block2", cout)):
        out[f"{prefix}.{blk}.proj.weight"] = (cout, ci, 3, 3, 3)
        out[f"{prefix}.{blk}.proj.bias"] = (cout,)
        out[f"{prefix}.{blk}.norm.weight"] = (cout,)
        out[f"{prefix}.{blk}.norm.bias"] = (cout,)
    if cin != cout:
        out[f"{prefix}.res_conv.weight"] = (cout, cin, 1, 1, 1)
        out[f"{prefix}.res_conv.bias"] = (cout,)


def _temporal_shapes(prefix, c, out, rot_dim=16):
    # Residual(PreNorm(LayerNorm, EinopsToAndFrom(Attention))) -> prefix.fn.norm / prefix.fn.fn.fn
    out[f"{prefix}.fn.fn.fn.rotary_emb.freqs"] = (rot_dim,)
    out[f"{prefix}.fn.fn.fn.to_qkv.weight"] = (HIDDEN * 3, c)
    out[f"{prefix}.fn.fn.fn.to_out.weight"] = (c, HIDDEN)
    out[f"{prefix}.fn.norm.gamma"] = (1, c, 1, 1, 1)


def _linattn_shapes(prefix, c, out):
    out[f"{prefix}.fn.fn.to_qkv.weight"] = (HIDDEN * 3, c, 1, 1)
    out[f"{prefix}.fn.fn.to_out.weight"] = (c, HIDDEN, 1, 1)
    out[f"{prefix}.fn.fn.to_out.bias"] = (c,)
    out[f"{prefix}.fn.norm.gamma"] = (1, c, 1, 1, 1)


def param_shapes(cfg: UnetCfg) -> "OrderedDict[str, tuple]":
    """state_dict() keys and shapes, in the reference's registration order (conv3d.py:373-471)."""
    o: "OrderedDict[str, tuple]" = OrderedDict()
    k = cfg.init_kernel_size
    td = cfg.time_dim
    o["time_rel_pos_bias.relative_attention_bias.weight"] = (32, HEADS)
    o["init_conv.weight"] = (cfg.dim, cfg.channels, k, k, k)
    o["init_conv.bias"] = (cfg.dim,)
    _temporal_shapes("init_temporal_attn", cfg.dim, o)
    o["time_mlp.1.weight"] = (td, cfg.dim)
    o["time_mlp.1.bias"] = (td,)
    o["time_mlp.3.weight"] = (td, td)
    o["time_mlp.3.bias"] = (td,)
    n = len(cfg.in_out)
    for i, (ci, co) in enumerate(cfg.in_out):
        _resnet_shapes(f"downs.{i}.0", ci, co, td, o)
        _resnet_shapes(f"downs.{i}.1", co, co, td, o)
        _linattn_shapes(f"downs.{i}.2", co, o)
        _temporal_shapes(f"downs.{i}.3", co, o)
        if i < n - 1:
            o[f"downs.{i}.4.weight"] = (co, co, 1, 4, 4)
            o[f"downs.{i}.4.bias"] = (co,)
    for i, (ci, co) in enumerate(reversed(cfg.in_out)):
        _resnet_shapes(f"ups.{i}.0", co * 2, ci, td, o)
        _resnet_shapes(f"ups.{i}.1", ci, ci, td, o)
        _linattn_shapes(f"ups.{i}.2", ci, o)
        _temporal_shapes(f"ups.{i}.3", ci, o)
        if i < n - 1:
            o[f"ups.{i}.4.weight"] = (ci, ci, 1, 4, 4)  # ConvTranspose3d: [Cin, Cout, 1, 4, 4]
            o[f"ups.{i}.4.bias"] = (ci,)
    # `downs` and `ups` ModuleLists are registered before the mid blocks (conv3d.py:422-423, :447-454)
    mid = cfg.dims[-1]
    _resnet_shapes("mid_block1", mid, mid, td, o)
    # mid spatial attention: Residual(PreNorm(EinopsToAndFrom(Attention without rotary)))
    o["mid_spatial_attn.fn.fn.fn.to_qkv.weight"] = (HIDDEN * 3, mid)
    o["mid_spatial_attn.fn.fn.fn.to_out.weight"] = (mid, HIDDEN)
    o["mid_spatial_attn.fn.norm.gamma"] = (1, mid, 1, 1, 1)
    _temporal_shapes("mid_temporal_attn", mid, o)
    _resnet_shapes("mid_block2", mid, mid, td, o)
    _resnet_shapes("final_conv.0", cfg.dim * 2, cfg.dim, None, o)
    o["final_conv.1.weight"] = (cfg.out_channels, cfg.dim, 1, 1, 1)
    o["final_conv.1.bias"] = (cfg.out_channels,)
    return o


def make_params(cfg: UnetCfg, seed: int = 0, out_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights, independent of module construction order: each tensor is drawn from its own
    generator seeded by crc32(key) ^ seed, so the same dict can be rebuilt on any box without the reference.
    Scales follow PyTorch's default inits (U(-1/sqrt(fan_in), 1/sqrt(fan_in))); norm gains are 1 + 0.1 N(0,1) and
    norm biases 0.1 N(0,1) so that a kernel that ignored them would be caught."""
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if key.endswith("rotary_emb.freqs"):
            d = shape[0] * 2
            t = 1.0 / (10000 ** (torch.arange(0, d, 2)[: d // 2].float() / d))
        elif key.endswith("norm.weight") or key.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("norm.bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("relative_attention_bias.weight"):
            t = torch.randn(shape, generator=g)
        else:
            if key.endswith(".bias"):
                wshape = param_shapes(cfg)[key[:-5] + ".weight"]
            else:
                wshape = shape
            if ".4.weight" in key and key.startswith("ups."):
                fan_in = wshape[1] * wshape[2] * wshape[3] * wshape[4]  # ConvTranspose: torch uses weight.size(1)*k
            else:
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if key.startswith("final_conv.1"):
                t = t * out_gain
        params[key] = t.float().contiguous()
    return params


# --------------------------------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------------------------------

def relative_position_bucket(relative_position, num_buckets=32, max_distance=32):
    """conv3d.py:86-104 (T5 bidirectional bucketing, n = -rel)."""
    ret = 0
    n = -relative_position
    num_buckets //= 2
    ret = ret + (n < 0).long() * num_buckets
    n = torch.abs(n)
    max_exact = num_buckets // 2
    is_small = n < max_exact
    val_if_large = max_exact + (
        torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).long()
    val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, num_buckets - 1))
    ret = ret + torch.where(is_small, n, val_if_large)
    return ret


def time_rel_pos_bias(weight: torch.Tensor, n: int) -> torch.Tensor:
    """conv3d.py:106-112 -> [heads, n, n]."""
    q_pos = torch.arange(n, dtype=torch.long)
    k_pos = torch.arange(n, dtype=torch.long)
    rel_pos = k_pos[None, :] - q_pos[:, None]
    bucket = relative_position_bucket(rel_pos, num_buckets=32, max_distance=32)
    values = F.embedding(bucket, weight)  # [i, j, h]
    return values.permute(2, 0, 1).contiguous()


def sinusoidal_pos_emb(x: torch.Tensor, dim: int) -> torch.Tensor:
    """conv3d.py:139-151."""
    half_dim = dim // 2
    emb = math.log(10000) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim) * -emb)
    emb = x[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def layer_norm_c(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """conv3d.py:165-174: channel-wise, biased variance, divide by sqrt(var+eps), gain only."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma


def rope(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """rotary-embedding-torch 0.8.4 rotate_queries_or_keys, positions along dim -2 (see oracle/shims)."""
    n = t.shape[-2]
    pos = torch.arange(n, dtype=freqs.dtype)
    ang = (pos[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
    x = t.reshape(*t.shape[:-1], t.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    rot = torch.stack((-x2, x1), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def attention(x, wqkv, wout, freqs=None, pos_bias=None):
    """conv3d.py:293-352 with focus_present_mask all-False. x: [..., n, c]."""
    qkv = F.linear(x, wqkv).chunk(3, dim=-1)
    q, k, v = [t.reshape(*t.shape[:-1], HEADS, DIM_HEAD).transpose(-2, -3) for t in qkv]  # ... h n d
    q = q * (DIM_HEAD ** -0.5)
    if freqs is not None:
        q = rope(q, freqs)
        k = rope(k, freqs)
    sim = torch.einsum("...hid,...hjd->...hij", q, k)
    if pos_bias is not None:
        sim = sim + pos_bias
    sim = sim - sim.amax(dim=-1, keepdim=True)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("...hij,...hjd->...hid", attn, v)
    out = out.transpose(-2, -3).reshape(*x.shape[:-1], HIDDEN)
    return F.linear(out, wout)


def temporal_attn_block(p, prefix, x, pos_bias):
    """Residual(PreNorm(EinopsToAndFrom('b c f h w','b (h w) f c', Attention+RoPE))) — conv3d.py:382, :394, :442."""
    b, c, f, h, w = x.shape
    xn = layer_norm_c(x, p[f"{prefix}.fn.norm.gamma"])
    t = xn.permute(0, 3, 4, 2, 1).reshape(b, h * w, f, c)
    o = attention(t, p[f"{prefix}.fn.fn.fn.to_qkv.weight"], p[f"{prefix}.fn.fn.fn.to_out.weight"],
                  freqs=p[f"{prefix}.fn.fn.fn.rotary_emb.freqs"], pos_bias=pos_bias)
    o = o.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2)
    return o + x


def mid_spatial_attn_block(p, prefix, x):
    """Residual(PreNorm(EinopsToAndFrom('b c f h w','b f (h w) c', Attention))) — conv3d.py:449-451."""
    b, c, f, h, w = x.shape
    xn = layer_norm_c(x, p[f"{prefix}.fn.norm.gamma"])
    t = xn.permute(0, 2, 3, 4, 1).reshape(b, f, h * w, c)
    o = attention(t, p[f"{prefix}.fn.fn.fn.to_qkv.weight"], p[f"{prefix}.fn.fn.fn.to_out.weight"])
    o = o.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3)
    return o + x


def spatial_linear_attn_block(p, prefix, x):
    """Residual(PreNorm(SpatialLinearAttention)) — conv3d.py:232-257 (note: v is NOT divided by h*w)."""
    b, c, f, h, w = x.shape
    xn = layer_norm_c(x, p[f"{prefix}.fn.norm.gamma"])
    t = xn.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    qkv = F.conv2d(t, p[f"{prefix}.fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = [u.reshape(b * f, HEADS, DIM_HEAD, h * w) for u in qkv]
    q = q.softmax(dim=-2)
    k = k.softmax(dim=-1)
    q = q * (DIM_HEAD ** -0.5)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b * f, HIDDEN, h, w)
    out = F.conv2d(out, p[f"{prefix}.fn.fn.to_out.weight"], p[f"{prefix}.fn.fn.to_out.bias"])
    out = out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
    return out + x


def block(p, prefix, x, groups, scale_shift=None):
    """conv3d.py:189-204: Conv3d 3x3x3 -> GroupNorm -> (scale+1, shift) -> SiLU."""
    x = F.conv3d(x, p[f"{prefix}.proj.weight"], p[f"{prefix}.proj.bias"], padding=1)
    x = F.group_norm(x, groups, p[f"{prefix}.norm.weight"], p[f"{prefix}.norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(p, prefix, x, t_emb, groups):
    """conv3d.py:206-230."""
    scale_shift = None
    if f"{prefix}.mlp.1.weight" in p:
        te = F.linear(F.silu(t_emb), p[f"{prefix}.mlp.1.weight"], p[f"{prefix}.mlp.1.bias"])
        te = te[:, :, None, None, None]
        scale_shift = te.chunk(2, dim=1)
    h = block(p, f"{prefix}.block1", x, groups, scale_shift)
    h = block(p, f"{prefix}.block2", h, groups)
    if f"{prefix}.res_conv.weight" in p:
        res = F.conv3d(x, p[f"{prefix}.res_conv.weight"], p[f"{prefix}.res_conv.bias"])
    else:
        res = x
    return h + res


@torch.no_grad()
def forward(p: Dict[str, torch.Tensor], cfg: UnetCfg, x: torch.Tensor, time: torch.Tensor, taps: Optional[dict] = None):
    """conv3d.py:486-552.  x: [B,F,C,H,W] fp32, time: [B] int64 -> [B,F,out,H,W].
    `taps`, if given, is filled with named intermediate activations (NCDHW) for per-stage parity checks."""
    g = cfg.resnet_groups
    x = x.permute(0, 2, 1, 3, 4)
    nfr = x.shape[2]
    pos_bias = time_rel_pos_bias(p["time_rel_pos_bias.relative_attention_bias.weight"], nfr)
    pad = cfg.init_kernel_size // 2
    x = F.conv3d(x, p["init_conv.weight"], p["init_conv.bias"], padding=pad)
    if taps is not None:
        taps["init_conv"] = x.clone()
    x = temporal_attn_block(p, "init_temporal_attn", x, pos_bias)
    if taps is not None:
        taps["init_temporal_attn"] = x.clone()
    r = x.clone()
    t = sinusoidal_pos_emb(time.float() if not time.is_floating_point() else time, cfg.dim)
    t = F.linear(t, p["time_mlp.1.weight"], p["time_mlp.1.bias"])
    t = F.gelu(t)
    t = F.linear(t, p["time_mlp.3.weight"], p["time_mlp.3.bias"])
    if taps is not None:
        taps["time_emb"] = t.clone()
    hs = []
    n = len(cfg.in_out)
    for i in range(n):
        x = resnet_block(p, f"downs.{i}.0", x, t, g)
        if taps is not None:
            taps[f"downs.{i}.0"] = x.clone()
        x = resnet_block(p, f"downs.{i}.1", x, t, g)
        x = spatial_linear_attn_block(p, f"downs.{i}.2", x)
        if taps is not None:
            taps[f"downs.{i}.2"] = x.clone()
        x = temporal_attn_block(p, f"downs.{i}.3", x, pos_bias)
        if taps is not None:
            taps[f"downs.{i}.3"] = x.clone()
        hs.append(x)
        if i < n - 1:
            x = F.conv3d(x, p[f"downs.{i}.4.weight"], p[f"downs.{i}.4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))
            if taps is not None:
                taps[f"downs.{i}.4"] = x.clone()
    x = resnet_block(p, "mid_block1", x, t, g)
    x = mid_spatial_attn_block(p, "mid_spatial_attn", x)
    if taps is not None:
        taps["mid_spatial_attn"] = x.clone()
    x = temporal_attn_block(p, "mid_temporal_attn", x, pos_bias)
    x = resnet_block(p, "mid_block2", x, t, g)
    if taps is not None:
        taps["mid_block2"] = x.clone()
    for i in range(n):
        x = torch.cat((x, hs.pop()), dim=1)
        x = resnet_block(p, f"ups.{i}.0", x, t, g)
        x = resnet_block(p, f"ups.{i}.1", x, t, g)
        x = spatial_linear_attn_block(p, f"ups.{i}.2", x)
        x = temporal_attn_block(p, f"ups.{i}.3", x, pos_bias)
        if taps is not None:
            taps[f"ups.{i}.3"] = x.clone()
        if i < n - 1:
            x = F.conv_transpose3d(x, p[f"ups.{i}.4.weight"], p[f"ups.{i}.4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))
            if taps is not None:
                taps[f"ups.{i}.4"] = x.clone()
    x = torch.cat((x, r), dim=1)
    x = resnet_block(p, "final_conv.0", x, None, g)
    x = F.conv3d(x, p["final_conv.1.weight"], p["final_conv.1.bias"])
    return x.permute(0, 2, 1, 3, 4).contiguous()
