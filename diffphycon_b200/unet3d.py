"""B200-native `Unet3D_with_Conv3D` — drop-in for the reference class of the same name
(model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py:356-552, cited below as conv3d.py:line).

Same constructor arguments, same `state_dict()` keys/shapes (so reference checkpoints load with strict=True), same
`forward(x [B,F,C,H,W], time [B]) -> [B,F,out_dim,H,W]`.  The module tree below only HOLDS parameters under the
reference's names; the arithmetic runs in the hand-written sm_100a kernels of libdpc_b200.so through `_lib`:
activations stay channels-last [B,F,H,W,C] from the stem to the final 1x1x1 conv, every Conv3d/Linear is a
tensor-core implicit GEMM (TF32 inputs, fp32 accumulate — the reference's own GPU numerics class),
GroupNorm statistics are produced by the conv epilogue, torch.cat is replaced by two-source operand loads.
There is no PyTorch fallback path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib, packing

HEAD_DIM = 32


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (names mirror the reference so that state_dict keys match; no forward methods)
# ----------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in Unet3D_with_Conv3D.forward")


class RotaryEmbedding(_Holder):
    """Parameter layout of rotary-embedding-torch 0.8.4 (`freqs`, non-trainable)."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)


class RelativePositionBias(_Holder):
    def __init__(self, heads=8, num_buckets=32, max_distance=128):
        super().__init__()
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)


class LayerNorm(_Holder):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.gamma = nn.Parameter(torch.ones(1, dim, 1, 1, 1))


class Attention(_Holder):
    def __init__(self, dim, heads=4, dim_head=32, rotary_emb=None):
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.rotary_emb = rotary_emb
        self.to_qkv = nn.Linear(dim, hidden * 3, bias=False)
        self.to_out = nn.Linear(hidden, dim, bias=False)


class SpatialLinearAttention(_Holder):
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class EinopsToAndFrom(_Holder):
    def __init__(self, from_einops, to_einops, fn):
        super().__init__()
        self.from_einops, self.to_einops = from_einops, to_einops
        self.fn = fn


class PreNorm(_Holder):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)


class Residual(_Holder):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class Block(_Holder):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = nn.Conv3d(dim, dim_out, (3, 3, 3), padding=(1, 1, 1))
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class ResnetBlock(_Holder):
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if time_emb_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv3d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out, self.groups = dim, dim_out, groups


class SinusoidalPosEmb(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


def _t5_buckets(n: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bucket index [n, n] of the T5 bidirectional relative position scheme with the reference's sign convention
    (conv3d.py:86-112).  Float32 torch ops in the reference's order, so ties round identically."""
    q = torch.arange(n, dtype=torch.long)
    rel = q[None, :] - q[:, None]
    neg = -rel
    half = num_buckets // 2
    ret = (neg < 0).long() * half
    a = neg.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(a.float() / max_exact) / math.log(max_distance / max_exact)
                         * (half - max_exact)).long()
    large = torch.min(large, torch.full_like(large, half - 1))
    return ret + torch.where(a < max_exact, a, large)


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("diffphycon_b200 runs on CUDA (sm_100a) only; there is no CPU path")


class _Pool:
    """Size-keyed free lists of device buffers: stable pointers across forwards (CUDA-graph friendly)."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[int, List[torch.Tensor]] = {}
        self.bytes = 0

    def get(self, numel: int, dtype=torch.float32) -> torch.Tensor:
        key = (numel, dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        self.bytes += numel * torch.empty((), dtype=dtype).element_size()
        return torch.empty(numel, dtype=dtype, device=self.device)

    def put(self, t: torch.Tensor):
        self.free.setdefault((t.numel(), t.dtype), []).append(t)


def pool_for(device) -> _Pool:
    """The scratch-buffer pool of (device, current stream).  Buffers are recycled in launch order on ONE stream, so a pool is
    never shared between streams: two models sampled concurrently on different streams get disjoint scratch memory."""
    device = torch.device(device)
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
    key = (device, stream)
    p = Unet3D_with_Conv3D._pools.get(key)
    if p is None:
        p = Unet3D_with_Conv3D._pools[key] = _Pool(device)
    return p


def release_pool(device=None):
    """Drop the cached scratch buffers (of one device, or all): the pool otherwise keeps the peak activation memory of the
    largest batch it has served for the life of the process."""
    for key in list(Unet3D_with_Conv3D._pools):
        if device is None or key[0] == torch.device(device):
            del Unet3D_with_Conv3D._pools[key]


class Unet3D_with_Conv3D(nn.Module):
    """See module docstring.  Constructor signature: conv3d.py:357-372.

    Scope: the SAMPLING surface of the reference class (forward under no_grad, the engine's kernels have no backward for this
    network).  Calling forward with autograd enabled on inputs / parameters that require grad raises — a reference training
    script must keep using the reference module."""

    def __init__(self, dim, cond_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=6, attn_heads=4,
                 attn_dim_head=32, use_bert_text_cond=False, init_dim=None, init_kernel_size=7,
                 use_sparse_linear_attn=True, block_type='resnet', resnet_groups=8):
        super().__init__()
        if cond_dim is not None or use_bert_text_cond:
            raise NotImplementedError("text conditioning is dead code on the DiffPhyCon path (use_bert_text_cond=False)")
        if attn_dim_head != HEAD_DIM:
            raise NotImplementedError("the attention kernels are specialised for dim_head = 32 (the reference default)")
        if not use_sparse_linear_attn or block_type != 'resnet':
            raise NotImplementedError("only the reference's default block configuration is implemented")
        assert init_kernel_size % 2 == 1
        self.channels = channels
        self.self_condition = False
        self.dim = dim
        self.heads = attn_heads
        self.groups = resnet_groups
        self.init_kernel_size = init_kernel_size
        self.has_cond = False
        self.null_cond_emb = None

        rotary_emb = RotaryEmbedding(min(32, attn_dim_head))

        def temporal_attn(d):
            return EinopsToAndFrom('b c f h w', 'b (h w) f c',
                                   Attention(d, heads=attn_heads, dim_head=attn_dim_head, rotary_emb=rotary_emb))

        self.time_rel_pos_bias = RelativePositionBias(heads=attn_heads, max_distance=32)
        init_dim = dim if init_dim is None else init_dim
        pad = init_kernel_size // 2
        self.init_conv = nn.Conv3d(channels, init_dim, (init_kernel_size,) * 3, padding=(pad,) * 3)
        self.init_temporal_attn = Residual(PreNorm(init_dim, temporal_attn(init_dim)))
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.in_out = in_out
        time_dim = dim * 4
        self.time_dim = time_dim
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n_res = len(in_out)
        for ind, (d_in, d_out) in enumerate(in_out):
            is_last = ind >= n_res - 1
            self.downs.append(nn.ModuleList([
                ResnetBlock(d_in, d_out, time_emb_dim=time_dim, groups=resnet_groups),
                ResnetBlock(d_out, d_out, time_emb_dim=time_dim, groups=resnet_groups),
                Residual(PreNorm(d_out, SpatialLinearAttention(d_out, heads=attn_heads))),
                Residual(PreNorm(d_out, temporal_attn(d_out))),
                nn.Conv3d(d_out, d_out, (1, 4, 4), (1, 2, 2), (0, 1, 1)) if not is_last else nn.Identity(),
            ]))
        mid = dims[-1]
        self.mid_block1 = ResnetBlock(mid, mid, time_emb_dim=time_dim, groups=resnet_groups)
        self.mid_spatial_attn = Residual(PreNorm(mid, EinopsToAndFrom('b c f h w', 'b f (h w) c',
                                                                      Attention(mid, heads=attn_heads))))
        self.mid_temporal_attn = Residual(PreNorm(mid, temporal_attn(mid)))
        self.mid_block2 = ResnetBlock(mid, mid, time_emb_dim=time_dim, groups=resnet_groups)
        for ind, (d_in, d_out) in enumerate(reversed(in_out)):
            is_last = ind >= n_res - 1
            self.ups.append(nn.ModuleList([
                ResnetBlock(d_out * 2, d_in, time_emb_dim=time_dim, groups=resnet_groups),
                ResnetBlock(d_in, d_in, time_emb_dim=time_dim, groups=resnet_groups),
                Residual(PreNorm(d_in, SpatialLinearAttention(d_in, heads=attn_heads))),
                Residual(PreNorm(d_in, temporal_attn(d_in))),
                nn.ConvTranspose3d(d_in, d_in, (1, 4, 4), (1, 2, 2), (0, 1, 1)) if not is_last else nn.Identity(),
            ]))
        self.out_dim = channels if out_dim is None else out_dim
        self.final_conv = nn.Sequential(ResnetBlock(dim * 2, dim, groups=resnet_groups), nn.Conv3d(dim, self.out_dim, 1))

        # engine state (not part of state_dict)
        self.use_tcgen05 = True       # TMA/tcgen05 kernel for the 3x3x3 convs when the shape allows it
        # "tf32": plain TF32 tensor-core products (the reference's own GPU numerics class: cudnn.allow_tf32=True);
        # "3xtf32": error-compensated split products on the generic kernel, near-fp32 (parity / debugging mode)
        self.precision = "tf32"
        self.micro_batch: Optional[int] = None  # samples per pass through the network (None = whole batch)
        self.paranoid_weight_check = False      # fingerprint the weights every forward (see invalidate_packed)
        self._packed = None
        self._packed_key = None
        self._geo_cache: Dict[tuple, dict] = {}
        self.taps: Optional[dict] = None  # set to {} to capture named intermediates (NCDHW copies) for parity tests

    # ------------------------------------------------------------------------------------------------------------
    # packing
    # ------------------------------------------------------------------------------------------------------------
    def _resnet_blocks(self):
        out = []
        for lvl in self.downs:
            out += [lvl[0], lvl[1]]
        out += [self.mid_block1, self.mid_block2]
        for lvl in self.ups:
            out += [lvl[0], lvl[1]]
        return out

    def _param_key(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self.paranoid_weight_check:
            # content fingerprint (one multi-tensor launch + one device->host read per forward): catches in-place edits through
            # `.data`, which neither move the storage nor bump the version counter
            norms = torch._foreach_norm([p.detach() for p in self.parameters()])
            key += (tuple(torch.stack(norms).double().cpu().tolist()),)
        return key

    def invalidate_packed(self):
        """Drop the packed weight copies (TF32-rounded K-major matrices, folded LayerNorm gains, position-bias tables).
        They are rebuilt automatically when a parameter is replaced or modified through autograd-visible in-place ops, by
        load_state_dict() and by .to()/.cuda(); call this after editing weights through `.data` (e.g. `p.data.copy_(...)`,
        ema_pytorch-style updates), which PyTorch does not version — or set `paranoid_weight_check = True`."""
        self._packed = None
        self._packed_key = None
        self._geo_cache.clear()

    def _load_from_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def _ensure_packed(self, device):
        assert self.precision in ("tf32", "3xtf32")
        key = (str(device), self.precision, self._param_key())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        P = {}
        dev = device
        rnd = self.precision == "tf32"
        cpad = packing.round_up(self.channels, 4)
        P["cpad"] = cpad
        P["init_conv.w"], _, _ = packing.pack_conv3d(self.init_conv.weight.to(dev), cin_pad=cpad, tf32=rnd)
        P["init_conv.b"] = self.init_conv.bias.detach().float().to(dev).contiguous()
        if rnd and self.init_kernel_size == 7 and self.init_conv.weight.shape[0] in (32, 64, 128):
            P["init_conv.ws"] = packing.pack_stem_conv(self.init_conv.weight.to(dev), cpad)   # dpc_stem_conv_tcgen05

        def pack_resnet(name, blk: ResnetBlock):
            for bn in ("block1", "block2"):
                b = getattr(blk, bn)
                P[f"{name}.{bn}.w"], _, _ = packing.pack_conv3d(b.proj.weight.to(dev), tf32=rnd)
                P[f"{name}.{bn}.b"] = b.proj.bias.detach().float().to(dev).contiguous()
                P[f"{name}.{bn}.gamma"] = b.norm.weight.detach().float().to(dev).contiguous()
                P[f"{name}.{bn}.beta"] = b.norm.bias.detach().float().to(dev).contiguous()
            if isinstance(blk.res_conv, nn.Conv3d):
                P[f"{name}.res.w"] = packing.pack_linear(blk.res_conv.weight.to(dev), tf32=rnd)
                P[f"{name}.res.b"] = blk.res_conv.bias.detach().float().to(dev).contiguous()

        def pack_attn(name, res: Residual, kind):
            pre = res.fn
            P[f"{name}.gamma"] = pre.norm.gamma.detach().float().to(dev).reshape(-1).contiguous()
            att = pre.fn.fn if kind != "linear" else pre.fn
            P[f"{name}.qkv.w"] = packing.pack_linear(att.to_qkv.weight.to(dev), tf32=rnd)
            P[f"{name}.out.w"] = packing.pack_linear(att.to_out.weight.to(dev), tf32=rnd)
            if att.to_out.bias is not None:
                P[f"{name}.out.b"] = att.to_out.bias.detach().float().to(dev).contiguous()
            if kind in ("temporal", "linear") and rnd:
                # operands of the fused block kernels (dpc_temporal_block_fused, dpc_spatial_linear_block_fused): LayerNorm
                # gain folded into to_qkv
                c_in = att.to_qkv.weight.shape[1]
                wq = att.to_qkv.weight.detach().float().to(dev).reshape(-1, c_in) * P[f"{name}.gamma"][None, :]
                P[f"{name}.qkv.wf"] = packing.tf32_round(wq).contiguous()
                wo = att.to_out.weight.detach().float().to(dev).reshape(c_in, -1)
                P[f"{name}.out.wf"] = (packing.tf32_round(wo) if kind == "temporal" else wo).contiguous()

        pack_attn("init_temporal_attn", self.init_temporal_attn, "temporal")
        for i, lvl in enumerate(self.downs):
            pack_resnet(f"downs.{i}.0", lvl[0])
            pack_resnet(f"downs.{i}.1", lvl[1])
            pack_attn(f"downs.{i}.2", lvl[2], "linear")
            pack_attn(f"downs.{i}.3", lvl[3], "temporal")
            if isinstance(lvl[4], nn.Conv3d):
                P[f"downs.{i}.4.w"], _, _ = packing.pack_conv3d(lvl[4].weight.to(dev), tf32=rnd)
                P[f"downs.{i}.4.b"] = lvl[4].bias.detach().float().to(dev).contiguous()
        pack_resnet("mid_block1", self.mid_block1)
        pack_attn("mid_spatial_attn", self.mid_spatial_attn, "spatial")
        pack_attn("mid_temporal_attn", self.mid_temporal_attn, "temporal")
        pack_resnet("mid_block2", self.mid_block2)
        for i, lvl in enumerate(self.ups):
            pack_resnet(f"ups.{i}.0", lvl[0])
            pack_resnet(f"ups.{i}.1", lvl[1])
            pack_attn(f"ups.{i}.2", lvl[2], "linear")
            pack_attn(f"ups.{i}.3", lvl[3], "temporal")
            if isinstance(lvl[4], nn.ConvTranspose3d):
                for cls, w in packing.pack_conv_transpose_1x4x4(lvl[4].weight.to(dev), tf32=rnd).items():
                    P[f"ups.{i}.4.w{cls[0]}{cls[1]}"] = w
                P[f"ups.{i}.4.b"] = lvl[4].bias.detach().float().to(dev).contiguous()
        pack_resnet("final_conv.0", self.final_conv[0])
        P["final_conv.1.w"] = packing.pack_linear(self.final_conv[1].weight.to(dev), tf32=rnd)
        P["final_conv.1.b"] = self.final_conv[1].bias.detach().float().to(dev).contiguous()
        P["final_conv.1.wraw"] = self.final_conv[1].weight.detach().float().to(dev).reshape(self.out_dim, -1).contiguous()

        # time conditioning: one row-concatenated matrix for all ResnetBlock mlps (conv3d.py:211-214)
        ws, bs, offs, off = [], [], {}, 0
        names = ([f"downs.{i}.{j}" for i in range(len(self.downs)) for j in (0, 1)] + ["mid_block1", "mid_block2"]
                 + [f"ups.{i}.{j}" for i in range(len(self.ups)) for j in (0, 1)])
        for name, blk in zip(names, self._resnet_blocks()):
            lin = blk.mlp[1]
            ws.append(lin.weight.detach().float().to(dev))
            bs.append(lin.bias.detach().float().to(dev))
            offs[name] = off
            off += lin.weight.shape[0]
        P["time_proj.w"] = torch.cat(ws, 0).contiguous()
        P["time_proj.b"] = torch.cat(bs, 0).contiguous()
        P["time_proj.offs"] = offs
        P["time_proj.total"] = off
        half = self.dim // 2
        emb = math.log(10000) / (half - 1)
        P["time.freqs"] = torch.exp(torch.arange(half) * -emb).float().to(dev).contiguous()  # conv3d.py:146-148
        P["time.w1"] = self.time_mlp[1].weight.detach().float().to(dev).contiguous()
        P["time.b1"] = self.time_mlp[1].bias.detach().float().to(dev).contiguous()
        P["time.w2"] = self.time_mlp[3].weight.detach().float().to(dev).contiguous()
        P["time.b2"] = self.time_mlp[3].bias.detach().float().to(dev).contiguous()
        P["rel_bias.w"] = self.time_rel_pos_bias.relative_attention_bias.weight.detach().float().to(dev)
        P["rope.freqs"] = self.init_temporal_attn.fn.fn.fn.rotary_emb.freqs.detach().float().cpu()
        self._packed, self._packed_key = P, key
        self._geo_cache.clear()
        return P

    def _geometry(self, device, F, H, W):
        """Per-(F,H,W) tables: tap tables per level, RoPE angles, relative position bias."""
        key = (str(device), F, H, W, self._packed_key)
        g = self._geo_cache.get(key)
        if g is not None:
            return g
        P = self._packed
        g = {}
        k = self.init_kernel_size
        g["taps.init"] = packing.tap_table(k, k, k, H, W, device)
        h, w = H, W
        n = len(self.in_out)
        for lvl in range(n):
            g[f"taps.333.{lvl}"] = packing.tap_table(3, 3, 3, h, w, device)
            g[f"taps.111.{lvl}"] = packing.tap_table(1, 1, 1, h, w, device)
            g[f"taps.144.{lvl}"] = packing.tap_table(1, 4, 4, h, w, device)
            g[f"taps.122.{lvl}"] = packing.tap_table(1, 2, 2, h, w, device)
            g[f"hw.{lvl}"] = (h, w)
            if lvl < n - 1:
                h, w = h // 2, w // 2
        # RoPE angle table, rotary-embedding-torch 0.8.4 semantics (interleaved pairs share a frequency)
        freqs = P["rope.freqs"]
        pos = torch.arange(F, dtype=freqs.dtype)
        ang = (pos[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
        g["rope.cos"] = ang.cos().to(device).contiguous()
        g["rope.sin"] = ang.sin().to(device).contiguous()
        rb = self.time_rel_pos_bias
        bucket = _t5_buckets(F, rb.num_buckets, rb.max_distance).to(device)
        g["pos_bias"] = P["rel_bias.w"][bucket].permute(2, 0, 1).contiguous()  # [heads, F, F]
        self._geo_cache[key] = g
        return g

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------------------
    def forward(self, x, time, cond=None, null_cond_prob=0., focus_present_mask=None, prob_focus_present=0.):
        """conv3d.py:486-552.  x: [B,F,C,H,W] fp32 CUDA, time: [B] -> [B,F,out_dim,H,W].  Sampling only (no autograd)."""
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("diffphycon_b200.Unet3D_with_Conv3D is a sampling engine: its kernels have no backward pass, "
                               "so an input that requires grad cannot be differentiated through it")
        return self._forward_nograd(x, time, cond, null_cond_prob, focus_present_mask, prob_focus_present)

    @torch.no_grad()
    @_lib.device_guarded
    def _forward_nograd(self, x, time, cond=None, null_cond_prob=0., focus_present_mask=None, prob_focus_present=0.):
        _require_cuda(x)
        if cond is not None:
            raise NotImplementedError("cond is unused on the DiffPhyCon path")
        if focus_present_mask is not None and bool(focus_present_mask.any()):
            raise NotImplementedError("focus_present_mask is all-False on the DiffPhyCon path (prob_focus_present=0)")
        B, F, C, H, W = x.shape
        assert C == self.channels, f"expected {self.channels} channels, got {C}"
        n_lvl = len(self.in_out)
        assert H % (2 ** (n_lvl - 1)) == 0 and W % (2 ** (n_lvl - 1)) == 0
        x = x.contiguous().float()
        time = time.to(device=x.device, dtype=torch.long).contiguous()
        out = torch.empty(B, F, self.out_dim, H, W, dtype=torch.float32, device=x.device)
        self._ensure_packed(x.device)
        mb = B if not self.micro_batch else min(self.micro_batch, B)
        for b0 in range(0, B, mb):
            b1 = min(B, b0 + mb)
            self._forward_chunk(x[b0:b1], time[b0:b1], out[b0:b1], 0, C)
        return out

    @torch.no_grad()
    @_lib.device_guarded
    def forward_slice(self, x_full, c0, time, out):
        """Same as forward but reads channels [c0, c0+self.channels) of a wider reference-layout tensor without
        materialising the slice (replaces x[:, :, 3:5] at smoke.py:612) and writes into a caller-provided `out`."""
        _require_cuda(x_full)
        B = x_full.shape[0]
        x_full = x_full.contiguous()
        time = time.to(device=x_full.device, dtype=torch.long).contiguous()
        self._ensure_packed(x_full.device)
        mb = B if not self.micro_batch else min(self.micro_batch, B)
        for b0 in range(0, B, mb):
            b1 = min(B, b0 + mb)
            self._forward_chunk(x_full[b0:b1], time[b0:b1], out[b0:b1], c0, x_full.shape[2])
        return out

    _pools: Dict[tuple, _Pool] = {}   # one buffer pool per (device, stream), shared by every network instance

    def _pool(self, device) -> _Pool:
        return pool_for(device)

    def _forward_chunk(self, x, time, out, c0, ctot):
        P = self._packed
        dev = x.device
        B, F, _, H, W = x.shape
        G = self._geometry(dev, F, H, W)
        pool = self._pool(dev)
        heads, hid = self.heads, self.heads * HEAD_DIM
        groups = self.groups
        taps_dbg = self.taps
        precise = self.precision == "3xtf32"
        n_lvl = len(self.in_out)
        sp = _lib.stream_ptr  # noqa: F841

        def rows(lvl):
            h, w = G[f"hw.{lvl}"]
            return B * F * h * w

        def tap(name, t, lvl, c):
            if taps_dbg is not None:
                h, w = G[f"hw.{lvl}"]
                taps_dbg[name] = t[: B * F * h * w * c].reshape(B, F, h, w, c).permute(0, 4, 1, 2, 3).clone()

        # GroupNorm statistics arena: one [B, groups, 2] double slot per Block, zeroed once per pass
        n_gn = 2 * (len(self._resnet_blocks()) + 1)
        stats = pool.get(n_gn * B * groups * 2, torch.float64)
        stats.zero_()
        gn_slot = [0]

        def next_stats():
            s = stats[gn_slot[0] * B * groups * 2:(gn_slot[0] + 1) * B * groups * 2]
            gn_slot[0] += 1
            return s

        def conv(xa, ca, w, bias, y, cout, lvl_in, kind, xb=None, cb=0, residual=None, gn=None, out_layout=0,
                 transposed_cls=None, tc=False, res_affine=None, tc_only=False):
            hi, wi = G[f"hw.{lvl_in}"]
            p = _lib.ConvParams()
            p.x1, p.x2 = xa.data_ptr(), (xb.data_ptr() if xb is not None else None)
            p.C1, p.C2 = ca, cb
            p.w, p.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
            p.residual = residual.data_ptr() if residual is not None else None
            p.res_scale, p.res_shift = (res_affine[0].data_ptr(), res_affine[1].data_ptr()) if res_affine is not None else (None, None)
            p.y = y.data_ptr()
            p.gn_stats = gn.data_ptr() if gn is not None else None
            p.gn_groups = groups if gn is not None else 0
            p.B, p.Fi, p.Hi, p.Wi = B, F, hi, wi
            p.Fo, p.Ho, p.Wo = F, hi, wi
            p.st = p.sh = p.sw = 1
            p.pt = p.ph = p.pw = 0
            p.oh_mul = p.ow_mul = 1
            p.oh_off = p.ow_off = 0
            if kind == "333":
                p.taps, p.ntaps = G[f"taps.333.{lvl_in}"].data_ptr(), 27
                p.pt = p.ph = p.pw = 1
            elif kind == "111":
                p.taps, p.ntaps = G[f"taps.111.{lvl_in}"].data_ptr(), 1
            elif kind == "init":
                k = self.init_kernel_size
                p.taps, p.ntaps = G["taps.init"].data_ptr(), k * k * k
                p.pt = p.ph = p.pw = k // 2
            elif kind == "down":
                p.taps, p.ntaps = G[f"taps.144.{lvl_in}"].data_ptr(), 16
                p.sh = p.sw = 2
                p.ph = p.pw = 1
                p.Ho, p.Wo = hi // 2, wi // 2
            elif kind == "up":
                ph_, pw_ = transposed_cls
                p.taps, p.ntaps = G[f"taps.122.{lvl_in}"].data_ptr(), 4
                p.ph, p.pw = 1 - ph_, 1 - pw_
                p.oh_mul = p.ow_mul = 2
                p.oh_off, p.ow_off = ph_, pw_
            else:
                raise AssertionError(kind)
            p.Hfull, p.Wfull = p.Ho * p.oh_mul, p.Wo * p.ow_mul
            p.Cout, p.Npad, p.Kpad = cout, w.shape[0], w.shape[1]
            p.out_layout = out_layout
            p.precise = 1 if precise else 0
            return _lib.conv(p, tcgen05=tc and not precise, tc_only=tc_only)

        ss = None  # [B, total] time scale/shift for every ResnetBlock

        def resnet(name, xa, ca, lvl, cout, xb=None, cb=0, has_time=True):
            m = rows(lvl)
            rps = m // B
            y1 = pool.get(m * cout)
            s1 = next_stats()
            conv(xa, ca, P[f"{name}.block1.w"], P[f"{name}.block1.b"], y1, cout, lvl, "333", xb=xb, cb=cb, gn=s1,
                 tc=self.use_tcgen05)
            _lib.groupnorm_silu(y1, s1, P[f"{name}.block1.gamma"], P[f"{name}.block1.beta"],
                                ss if has_time else None, P["time_proj.total"],
                                P["time_proj.offs"].get(name, 0), None, y1, B, rps, cout, groups)
            y2 = pool.get(m * cout)
            s2 = next_stats()
            conv(y1, cout, P[f"{name}.block2.w"], P[f"{name}.block2.b"], y2, cout, lvl, "333", gn=s2,
                 tc=self.use_tcgen05)
            pool.put(y1)
            if f"{name}.res.w" in P and self.use_tcgen05 and not precise:
                # block2's GroupNorm-apply + SiLU folded into the residual operand of the res_conv GEMM (one pass instead of
                # res_conv -> res, then norm(y2) + res): out = silu(GN(y2)) + res_conv(x)
                ga, gd = pool.get(B * cout), pool.get(B * cout)
                _lib.gn_fold(s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], ga, gd, B, rps, cout, groups)
                outb = pool.get(m * cout)
                ok = conv(xa, ca, P[f"{name}.res.w"], P[f"{name}.res.b"], outb, cout, lvl, "111", xb=xb, cb=cb, residual=y2,
                          res_affine=(ga, gd), tc=True, tc_only=True)
                pool.put(ga)
                pool.put(gd)
                if ok:
                    pool.put(y2)
                    return outb
                pool.put(outb)
            if f"{name}.res.w" in P:
                res = pool.get(m * cout)
                conv(xa, ca, P[f"{name}.res.w"], P[f"{name}.res.b"], res, cout, lvl, "111", xb=xb, cb=cb, tc=self.use_tcgen05)
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, res, y2,
                                    B, rps, cout, groups)
                pool.put(res)
            else:
                assert xb is None
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, xa, y2,
                                    B, rps, cout, groups)
            return y2

        def attention(name, xa, c, lvl, kind):
            m = rows(lvl)
            h, w = G[f"hw.{lvl}"]
            if kind == "temporal" and self.use_tcgen05 and not precise and f"{name}.qkv.wf" in P:
                y = pool.get(m * c)
                if _lib.temporal_block_fused(xa, P[f"{name}.qkv.wf"], P[f"{name}.out.wf"], G["rope.cos"], G["rope.sin"],
                                             G["pos_bias"], y, B, F, h * w, c, heads):
                    return y
                pool.put(y)
            if kind == "linear" and self.use_tcgen05 and not precise and f"{name}.qkv.wf" in P:
                y = pool.get(m * c)
                ctx = pool.get(B * F * heads * HEAD_DIM * HEAD_DIM)
                mt = pool.get(B * F * c * hid)
                ok = _lib.spatial_linear_block_fused(xa, P[f"{name}.qkv.wf"], P[f"{name}.out.wf"], P.get(f"{name}.out.b"),
                                                     ctx, mt, y, B * F, h * w, c, heads)
                pool.put(ctx)
                pool.put(mt)
                if ok:
                    return y
                pool.put(y)
            xn = pool.get(m * c)
            _lib.layernorm_channels(xa, P[f"{name}.gamma"], xn, m, c)
            qkv = pool.get(m * 3 * hid)
            conv(xn, c, P[f"{name}.qkv.w"], None, qkv, 3 * hid, lvl, "111", tc=self.use_tcgen05)
            pool.put(xn)
            att = pool.get(m * hid)
            if kind == "temporal":
                _lib.temporal_attention(qkv, G["rope.cos"], G["rope.sin"], G["pos_bias"], att, B, F, h * w, heads, True, precise,
                                        relative_bias=True)
            elif kind == "spatial":
                _lib.spatial_attention(qkv, att, B * F, h * w, heads, precise=precise)
            else:
                ctx = pool.get(B * F * heads * HEAD_DIM * HEAD_DIM)
                _lib.spatial_linear_attention(qkv, ctx, att, B * F, h * w, heads)
                pool.put(ctx)
            pool.put(qkv)
            y = pool.get(m * c)
            conv(att, hid, P[f"{name}.out.w"], P.get(f"{name}.out.b"), y, c, lvl, "111", residual=xa, tc=self.use_tcgen05)
            pool.put(att)
            return y

        # ---- stem (conv3d.py:495-505) ----
        cpad = P["cpad"]
        m0 = rows(0)
        xin = pool.get(m0 * cpad)
        _lib.pack_input(x, xin, B, F, ctot, c0, self.channels, H, W, cpad)
        d0 = self.in_out[0][0]
        h0 = pool.get(m0 * d0)
        k = self.init_kernel_size
        if not (self.use_tcgen05 and not precise and "init_conv.ws" in P
                and _lib.stem_conv(xin, P["init_conv.ws"], P["init_conv.b"], h0, B, F, H, W, cpad, d0, k, k, k)):
            conv(xin, cpad, P["init_conv.w"], P["init_conv.b"], h0, d0, 0, "init")
        pool.put(xin)
        tap("init_conv", h0, 0, d0)
        h1 = attention("init_temporal_attn", h0, d0, 0, "temporal")
        pool.put(h0)
        tap("init_temporal_attn", h1, 0, d0)
        r = h1
        # ---- time conditioning (conv3d.py:507-509) ----
        tdim = self.time_dim
        hidden = pool.get(B * tdim)
        t_emb = pool.get(B * tdim)
        _lib.time_embed(time, P["time.freqs"], P["time.w1"], P["time.b1"], P["time.w2"], P["time.b2"], hidden, t_emb, B,
                        self.dim)
        ss = pool.get(B * P["time_proj.total"])
        _lib.time_proj(t_emb, P["time_proj.w"], P["time_proj.b"], ss, B, tdim, P["time_proj.total"])
        if taps_dbg is not None:
            taps_dbg["time_emb"] = t_emb[: B * tdim].reshape(B, tdim).clone()
        pool.put(hidden)
        pool.put(t_emb)

        # ---- down path (conv3d.py:522-530) ----
        cur, cur_c = r, d0
        skips = []
        for i, (d_in, d_out) in enumerate(self.in_out):
            a = resnet(f"downs.{i}.0", cur, cur_c, i, d_out)
            if cur is not r:
                pool.put(cur)
            tap(f"downs.{i}.0", a, i, d_out)
            b = resnet(f"downs.{i}.1", a, d_out, i, d_out)
            pool.put(a)
            c = attention(f"downs.{i}.2", b, d_out, i, "linear")
            pool.put(b)
            tap(f"downs.{i}.2", c, i, d_out)
            d = attention(f"downs.{i}.3", c, d_out, i, "temporal")
            pool.put(c)
            tap(f"downs.{i}.3", d, i, d_out)
            skips.append((d, d_out))
            if i < n_lvl - 1:
                e = pool.get(rows(i + 1) * d_out)
                conv(d, d_out, P[f"downs.{i}.4.w"], P[f"downs.{i}.4.b"], e, d_out, i, "down", tc=self.use_tcgen05)
                tap(f"downs.{i}.4", e, i + 1, d_out)
                cur, cur_c = e, d_out
            else:
                cur, cur_c = d, d_out
        # ---- middle (conv3d.py:532-535) ----
        top = n_lvl - 1
        a = resnet("mid_block1", cur, cur_c, top, cur_c)   # cur is also the last skip: keep it alive
        b = attention("mid_spatial_attn", a, cur_c, top, "spatial")
        pool.put(a)
        tap("mid_spatial_attn", b, top, cur_c)
        c = attention("mid_temporal_attn", b, cur_c, top, "temporal")
        pool.put(b)
        cur = resnet("mid_block2", c, cur_c, top, cur_c)
        pool.put(c)
        tap("mid_block2", cur, top, cur_c)
        # ---- up path (conv3d.py:537-543) ----
        for i, (d_in, d_out) in enumerate(reversed(self.in_out)):
            lvl = top - i
            skip, skip_c = skips.pop()
            a = resnet(f"ups.{i}.0", cur, cur_c, lvl, d_in, xb=skip, cb=skip_c)
            pool.put(cur)
            pool.put(skip)
            b = resnet(f"ups.{i}.1", a, d_in, lvl, d_in)
            pool.put(a)
            c = attention(f"ups.{i}.2", b, d_in, lvl, "linear")
            pool.put(b)
            d = attention(f"ups.{i}.3", c, d_in, lvl, "temporal")
            pool.put(c)
            tap(f"ups.{i}.3", d, lvl, d_in)
            if i < n_lvl - 1:
                e = pool.get(rows(lvl - 1) * d_in)
                for cls in ((0, 0), (0, 1), (1, 0), (1, 1)):
                    conv(d, d_in, P[f"ups.{i}.4.w{cls[0]}{cls[1]}"], P[f"ups.{i}.4.b"], e, d_in, lvl, "up",
                         transposed_cls=cls, tc=self.use_tcgen05)
                pool.put(d)
                tap(f"ups.{i}.4", e, lvl - 1, d_in)
                cur, cur_c = e, d_in
            else:
                cur, cur_c = d, d_in
        # ---- head (conv3d.py:545-549) ----
        f0 = resnet("final_conv.0", cur, cur_c, 0, self.dim, xb=r, cb=d0, has_time=False)
        pool.put(cur)
        pool.put(r)
        if not (out.is_contiguous() and _lib.final_proj(f0, P["final_conv.1.wraw"], P["final_conv.1.b"], out, B * F, H * W,
                                                       self.dim, self.out_dim)):
            conv(f0, self.dim, P["final_conv.1.w"], P["final_conv.1.b"], out, self.out_dim, 0, "111", out_layout=1)
        pool.put(f0)
        pool.put(ss)
        pool.put(stats)
