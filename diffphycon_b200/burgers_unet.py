"""B200-native `Unet2D` — drop-in for model/burgers_1d/unet.py:268-431 (cited as unet.py:line), the (time, space) U-Net of
the Burgers task.  Same constructor arguments, same `state_dict()` keys/shapes, same
`forward(x [B,C,H,W], time [B]) -> [B,out_dim,H,W]`.  The module tree only holds parameters; the arithmetic runs in the
kernels of libdpc_b200.so: activations are channels-last [B,H,W,C], every Conv2d (7x7 stem, 3x3, 1x1, the pixel-unshuffle
down-sampling expressed as a 2x2 stride-2 conv) is the tensor-core implicit GEMM with fused bias / residual / GroupNorm
statistics, torch.cat is a two-source operand load.  No PyTorch fallback."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib, packing
from .unet3d import _Pool, _require_cuda

HEAD_DIM = 32


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in Unet2D.forward")


class LayerNorm(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


class Block(_Holder):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class ResnetBlock(_Holder):
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if time_emb_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim_out, self.groups = dim_out, groups


class LinearAttention(_Holder):
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), LayerNorm(dim))


class Attention(_Holder):
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads = heads
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class PreNorm(_Holder):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)


class Residual(_Holder):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class SinusoidalPosEmb(_Holder):
    def __init__(self, dim, theta=10000):
        super().__init__()
        self.dim, self.theta = dim, theta


class Unet2D(nn.Module):
    """Constructor: unet.py:273-290."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=2, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000,
                 attn_dim_head=32, attn_heads=4, condition_on_residual=None):
        super().__init__()
        if self_condition or learned_sinusoidal_cond or random_fourier_features or condition_on_residual:
            raise NotImplementedError("only the configuration used by the DiffPhyCon Burgers runs is implemented")
        if attn_dim_head != HEAD_DIM:
            raise NotImplementedError("the attention kernels are specialised for dim_head = 32")
        assert dim % 2 == 0
        self.condition_on_residual = None
        self.channels = channels
        self.self_condition = False
        self.random_or_learned_sinusoidal_cond = False
        self.dim = dim
        self.heads = attn_heads
        self.groups = resnet_block_groups
        time_dim = dim * 4
        self.time_dim = time_dim
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim, theta=sinusoidal_pos_emb_theta), nn.Linear(dim, time_dim),
                                      nn.GELU(), nn.Linear(time_dim, time_dim))
        init_dim = dim if init_dim is None else init_dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.in_out = in_out
        rb = lambda a, b: ResnetBlock(a, b, time_emb_dim=time_dim, groups=resnet_block_groups)
        self.downs = nn.ModuleList([])
        n = len(in_out)
        for ind, (d_in, d_out) in enumerate(in_out):
            is_last = ind >= n - 1
            down = (nn.Sequential(nn.Identity(), nn.Conv2d(d_in * 4, d_out, 1)) if not is_last
                    else nn.Conv2d(d_in, d_out, 3, padding=1))
            self.downs.append(nn.ModuleList([rb(d_in, d_in), rb(d_in, d_in), Residual(PreNorm(d_in, LinearAttention(d_in))),
                                             down]))
        mid = dims[-1]
        self.mid_block1 = rb(mid, mid)
        self.mid_attn = Residual(PreNorm(mid, Attention(mid, dim_head=attn_dim_head, heads=attn_heads)))
        self.mid_block2 = rb(mid, mid)
        self.ups = nn.ModuleList([])
        for ind, (d_in, d_out) in enumerate(reversed(in_out)):
            is_last = ind == n - 1
            up = (nn.Sequential(nn.Upsample(scale_factor=2, mode='nearest'), nn.Conv2d(d_out, d_in, 3, padding=1))
                  if not is_last else nn.Conv2d(d_out, d_in, 3, padding=1))
            self.ups.append(nn.ModuleList([rb(d_out + d_in, d_out), rb(d_out + d_in, d_out),
                                           Residual(PreNorm(d_out, LinearAttention(d_out))), up]))
        self.out_dim = (channels * (1 if not learned_variance else 2)) if out_dim is None else out_dim
        self.final_res_block = rb(dim * 2, dim)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        # engine state
        self.precision = "tf32"     # or "3xtf32" (fp32-class), see Unet3D_with_Conv3D
        self.use_tcgen05 = True     # TF32 mode: 3x3 / 1x1 layers on the TMA + tcgen05 kernel where it tiles the shape
        self._packed = None
        self._packed_key = None
        self._taps: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------------------
    def _resnets(self):
        out = []
        for i, lvl in enumerate(self.downs):
            out += [(f"downs.{i}.0", lvl[0]), (f"downs.{i}.1", lvl[1])]
        out += [("mid_block1", self.mid_block1), ("mid_block2", self.mid_block2)]
        for i, lvl in enumerate(self.ups):
            out += [(f"ups.{i}.0", lvl[0]), (f"ups.{i}.1", lvl[1])]
        out += [("final_res_block", self.final_res_block)]
        return out

    def invalidate_packed(self):
        """Drop the packed weight copies; call after editing weights through `.data` (see Unet3D_with_Conv3D.invalidate_packed)."""
        self._packed = None
        self._packed_key = None

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_packed(self, dev):
        key = (str(dev), self.precision, self._param_key())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        rnd = self.precision == "tf32"
        P = {}
        f32 = lambda t: t.detach().float().to(dev).contiguous()

        def conv_w(w, cin_pad=None):
            w5 = w.to(dev).unsqueeze(2)  # [Cout, Cin, 1, kh, kw]
            return packing.pack_conv3d(w5, cin_pad=cin_pad, tf32=rnd)[0]

        cpad = packing.round_up(self.channels, 4)
        P["cpad"] = cpad
        P["init.w"], P["init.b"] = conv_w(self.init_conv.weight, cpad), f32(self.init_conv.bias)
        for name, blk in self._resnets():
            for bn in ("block1", "block2"):
                b = getattr(blk, bn)
                P[f"{name}.{bn}.w"], P[f"{name}.{bn}.b"] = conv_w(b.proj.weight), f32(b.proj.bias)
                P[f"{name}.{bn}.gamma"], P[f"{name}.{bn}.beta"] = f32(b.norm.weight), f32(b.norm.bias)
            if isinstance(blk.res_conv, nn.Conv2d):
                P[f"{name}.res.w"] = packing.pack_linear(blk.res_conv.weight.to(dev), tf32=rnd)
                P[f"{name}.res.b"] = f32(blk.res_conv.bias)

        def pack_attn(name, res: Residual):
            P[f"{name}.g"] = f32(res.fn.norm.g).reshape(-1)
            att = res.fn.fn
            P[f"{name}.qkv.w"] = packing.pack_linear(att.to_qkv.weight.to(dev), tf32=rnd)
            out = att.to_out[0] if isinstance(att, LinearAttention) else att.to_out
            P[f"{name}.out.w"], P[f"{name}.out.b"] = packing.pack_linear(out.weight.to(dev), tf32=rnd), f32(out.bias)
            if isinstance(att, LinearAttention):
                P[f"{name}.out.g"] = f32(att.to_out[1].g).reshape(-1)

        for i, lvl in enumerate(self.downs):
            pack_attn(f"downs.{i}.2", lvl[2])
            if isinstance(lvl[3], nn.Sequential):
                w = lvl[3][1].weight.to(dev)  # [Cout, 4*C, 1, 1], input channel = c*4 + p1*2 + p2 (unet.py:46-50)
                co, c4 = w.shape[0], w.shape[1]
                wk = w.reshape(co, c4 // 4, 4).permute(0, 2, 1).reshape(co, c4)   # k = (p1*2 + p2)*C + c
                P[f"downs.{i}.3.w"] = packing.pack_linear(wk, tf32=rnd)
                P[f"downs.{i}.3.b"] = f32(lvl[3][1].bias)
            else:
                P[f"downs.{i}.3.w"], P[f"downs.{i}.3.b"] = conv_w(lvl[3].weight), f32(lvl[3].bias)
        pack_attn("mid_attn", self.mid_attn)
        for i, lvl in enumerate(self.ups):
            pack_attn(f"ups.{i}.2", lvl[2])
            c = lvl[3][1] if isinstance(lvl[3], nn.Sequential) else lvl[3]
            P[f"ups.{i}.3.w"], P[f"ups.{i}.3.b"] = conv_w(c.weight), f32(c.bias)
        P["final.w"] = packing.pack_linear(self.final_conv.weight.to(dev), tf32=rnd)
        P["final.b"] = f32(self.final_conv.bias)
        ws, bs, offs, off = [], [], {}, 0
        for name, blk in self._resnets():
            lin = blk.mlp[1]
            ws.append(f32(lin.weight))
            bs.append(f32(lin.bias))
            offs[name] = off
            off += lin.weight.shape[0]
        P["tp.w"], P["tp.b"], P["tp.offs"], P["tp.total"] = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous(), offs, off
        half = self.dim // 2
        theta = self.time_mlp[0].theta
        P["t.freqs"] = torch.exp(torch.arange(half) * -(math.log(theta) / (half - 1))).float().to(dev).contiguous()
        P["t.w1"], P["t.b1"] = f32(self.time_mlp[1].weight), f32(self.time_mlp[1].bias)
        P["t.w2"], P["t.b2"] = f32(self.time_mlp[3].weight), f32(self.time_mlp[3].bias)
        self._packed, self._packed_key = P, key
        return P

    def _tap(self, kh, kw, h, w, dev):
        key = (kh, kw, h, w, str(dev))
        t = self._taps.get(key)
        if t is None:
            t = self._taps[key] = packing.tap_table(1, kh, kw, h, w, dev)
        return t

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    @_lib.device_guarded
    def forward(self, x, time, x_self_cond=None, residual=None):
        """unet.py:387-431.  x: [B,C,H,W] fp32 CUDA, time: [B] -> [B,out_dim,H,W]."""
        _require_cuda(x)
        if x_self_cond is not None or residual is not None:
            raise NotImplementedError("self-conditioning / residual conditioning are unused by the DiffPhyCon Burgers runs")
        B, C, H, W = x.shape
        assert C == self.channels
        n = len(self.in_out)
        assert H % (2 ** (n - 1)) == 0 and W % (2 ** (n - 1)) == 0
        dev = x.device
        x = x.contiguous().float()
        time = time.to(device=dev, dtype=torch.long).contiguous()
        P = self._ensure_packed(dev)
        pool = Unet3D_pool(dev)
        precise = self.precision == "3xtf32"
        heads, hid, groups = self.heads, self.heads * HEAD_DIM, self.groups
        out = torch.empty(B, self.out_dim, H, W, dtype=torch.float32, device=dev)

        def conv(xa, ca, w, bias, y, cout, h, wd, kh=1, kw=1, stride=1, pad=0, xb=None, cb=0, residual=None, gn=None,
                 out_layout=0):
            p = _lib.ConvParams()
            p.x1, p.x2 = xa.data_ptr(), (xb.data_ptr() if xb is not None else None)
            p.C1, p.C2 = ca, cb
            p.w, p.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
            p.residual = residual.data_ptr() if residual is not None else None
            p.y = y.data_ptr()
            p.gn_stats = gn.data_ptr() if gn is not None else None
            p.gn_groups = groups if gn is not None else 0
            p.B, p.Fi, p.Hi, p.Wi = B, 1, h, wd
            p.Fo, p.Ho, p.Wo = 1, (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
            p.st, p.sh, p.sw = 1, stride, stride
            p.pt, p.ph, p.pw = 0, pad, pad
            p.oh_mul = p.ow_mul = 1
            p.oh_off = p.ow_off = 0
            p.Hfull, p.Wfull = p.Ho, p.Wo
            p.taps, p.ntaps = self._tap(kh, kw, h, wd, dev).data_ptr(), kh * kw
            p.Cout, p.Npad, p.Kpad = cout, w.shape[0], w.shape[1]
            p.out_layout, p.precise = out_layout, (1 if precise else 0)
            if not tc:
                _lib.conv(p, tcgen05=False)
                return
            # TF32 mode: the TMA / tcgen05 kernel serves the 3x3 and 1x1 layers whose channel counts it tiles (multiples of 32 in, 64 /
            # 128 / 256 / 512 out); its epilogue produces GroupNorm(8) statistics, merged here into the coarser grouping of
            # resnet_block_groups = 1 / 2 / 4 nets.  Everything it declines (-2) runs on the mma.sync implicit GEMM as before.
            if gn is not None and groups != 8 and 8 % groups == 0 and cout % 8 == 0:
                s8 = stats8[slot8[0] * B * 16:(slot8[0] + 1) * B * 16]
                slot8[0] += 1
                p.gn_stats, p.gn_groups = s8.data_ptr(), 8
                if _lib.conv(p, tcgen05=True, tc_only=True):
                    _lib.gn_stats_merge(s8, gn, B, 8, groups)
                    return
                p.gn_stats, p.gn_groups = gn.data_ptr(), groups
                _lib.conv(p, tcgen05=False)
            else:
                _lib.conv(p, tcgen05=True)

        n_gn = 2 * len(self._resnets())
        stats = pool.get(n_gn * B * groups * 2, torch.float64)
        stats.zero_()
        slot = [0]
        tc = bool(self.use_tcgen05) and not precise
        stats8, slot8 = None, [0]
        if tc and groups != 8:
            stats8 = pool.get(n_gn * B * 16, torch.float64)
            stats8.zero_()

        def next_stats():
            s = stats[slot[0] * B * groups * 2:(slot[0] + 1) * B * groups * 2]
            slot[0] += 1
            return s

        # time conditioning (unet.py:401, :155-181)
        tdim = self.time_dim
        hidden, t_emb = pool.get(B * tdim), pool.get(B * tdim)
        _lib.time_embed(time, P["t.freqs"], P["t.w1"], P["t.b1"], P["t.w2"], P["t.b2"], hidden, t_emb, B, self.dim)
        ss = pool.get(B * P["tp.total"])
        _lib.time_proj(t_emb, P["tp.w"], P["tp.b"], ss, B, tdim, P["tp.total"])
        pool.put(hidden)
        pool.put(t_emb)

        def resnet(name, xa, ca, h, wd, cout, xb=None, cb=0):
            m = B * h * wd
            y1, s1 = pool.get(m * cout), next_stats()
            conv(xa, ca, P[f"{name}.block1.w"], P[f"{name}.block1.b"], y1, cout, h, wd, 3, 3, 1, 1, xb=xb, cb=cb, gn=s1)
            _lib.groupnorm_silu(y1, s1, P[f"{name}.block1.gamma"], P[f"{name}.block1.beta"], ss, P["tp.total"],
                                P["tp.offs"][name], None, y1, B, h * wd, cout, groups)
            y2, s2 = pool.get(m * cout), next_stats()
            conv(y1, cout, P[f"{name}.block2.w"], P[f"{name}.block2.b"], y2, cout, h, wd, 3, 3, 1, 1, gn=s2)
            pool.put(y1)
            if f"{name}.res.w" in P:
                res = pool.get(m * cout)
                conv(xa, ca, P[f"{name}.res.w"], P[f"{name}.res.b"], res, cout, h, wd, xb=xb, cb=cb)
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, res, y2, B,
                                    h * wd, cout, groups)
                pool.put(res)
            else:
                assert xb is None
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, xa, y2, B,
                                    h * wd, cout, groups)
            return y2

        def attention(name, xa, c, h, wd, linear):
            m = B * h * wd
            xn = pool.get(m * c)
            _lib.layernorm_channels(xa, P[f"{name}.g"], xn, m, c, use_rsqrt=True)              # PreNorm (unet.py:72-83)
            qkv = pool.get(m * 3 * hid)
            conv(xn, c, P[f"{name}.qkv.w"], None, qkv, 3 * hid, h, wd)
            pool.put(xn)
            att = pool.get(m * hid)
            if linear:
                ctx = pool.get(B * heads * HEAD_DIM * HEAD_DIM)
                _lib.spatial_linear_attention(qkv, ctx, att, B, h * wd, heads)                   # unet.py:209-222
                pool.put(ctx)
            else:
                _lib.spatial_attention(qkv, att, B, h * wd, heads)                               # unet.py:246-262
            pool.put(qkv)
            y = pool.get(m * c)
            if linear:
                conv(att, hid, P[f"{name}.out.w"], P[f"{name}.out.b"], y, c, h, wd)
                _lib.layernorm_channels(y, P[f"{name}.out.g"], y, m, c, residual=xa, use_rsqrt=True)   # to_out[1], Residual
            else:
                conv(att, hid, P[f"{name}.out.w"], P[f"{name}.out.b"], y, c, h, wd, residual=xa)
            pool.put(att)
            return y

        # stem (unet.py:398-399)
        cpad = P["cpad"]
        xin = pool.get(B * H * W * cpad)
        _lib.pack_input(x, xin, B, 1, C, 0, C, H, W, cpad)
        d0 = self.in_out[0][0]
        r = pool.get(B * H * W * d0)
        conv(xin, cpad, P["init.w"], P["init.b"], r, d0, H, W, 7, 7, 1, 3)
        pool.put(xin)
        cur, cur_c, h, w = r, d0, H, W
        skips = []
        for i, (d_in, d_out) in enumerate(self.in_out):
            a = resnet(f"downs.{i}.0", cur, cur_c, h, w, d_in)
            if cur is not r:
                pool.put(cur)
            skips.append((a, d_in))
            b = resnet(f"downs.{i}.1", a, d_in, h, w, d_in)
            c = attention(f"downs.{i}.2", b, d_in, h, w, True)
            pool.put(b)
            skips.append((c, d_in))
            if i < n - 1:
                e = pool.get(B * (h // 2) * (w // 2) * d_out)
                conv(c, d_in, P[f"downs.{i}.3.w"], P[f"downs.{i}.3.b"], e, d_out, h, w, 2, 2, 2, 0)   # pixel-unshuffle + 1x1
                h, w = h // 2, w // 2
            else:
                e = pool.get(B * h * w * d_out)
                conv(c, d_in, P[f"downs.{i}.3.w"], P[f"downs.{i}.3.b"], e, d_out, h, w, 3, 3, 1, 1)
            cur, cur_c = e, d_out
        a = resnet("mid_block1", cur, cur_c, h, w, cur_c)
        pool.put(cur)
        b = attention("mid_attn", a, cur_c, h, w, False)
        pool.put(a)
        cur = resnet("mid_block2", b, cur_c, h, w, cur_c)
        pool.put(b)
        for i, (d_in, d_out) in enumerate(reversed(self.in_out)):
            s1, s1c = skips.pop()
            a = resnet(f"ups.{i}.0", cur, cur_c, h, w, d_out, xb=s1, cb=s1c)
            pool.put(cur)
            pool.put(s1)
            s2, s2c = skips.pop()
            b = resnet(f"ups.{i}.1", a, d_out, h, w, d_out, xb=s2, cb=s2c)
            pool.put(a)
            pool.put(s2)
            c = attention(f"ups.{i}.2", b, d_out, h, w, True)
            pool.put(b)
            if i < n - 1:
                up = pool.get(B * 4 * h * w * d_out)
                _lib.upsample_nearest2x(c, up, B, h, w, d_out)
                pool.put(c)
                h, w = 2 * h, 2 * w
                e = pool.get(B * h * w * d_in)
                conv(up, d_out, P[f"ups.{i}.3.w"], P[f"ups.{i}.3.b"], e, d_in, h, w, 3, 3, 1, 1)
                pool.put(up)
            else:
                e = pool.get(B * h * w * d_in)
                conv(c, d_out, P[f"ups.{i}.3.w"], P[f"ups.{i}.3.b"], e, d_in, h, w, 3, 3, 1, 1)
                pool.put(c)
            cur, cur_c = e, d_in
        f0 = resnet("final_res_block", cur, cur_c, H, W, self.dim, xb=r, cb=d0)
        pool.put(cur)
        pool.put(r)
        conv(f0, self.dim, P["final.w"], P["final.b"], out, self.out_dim, H, W, out_layout=1)
        pool.put(f0)
        pool.put(ss)
        pool.put(stats)
        if stats8 is not None:
            pool.put(stats8)
        return out


def Unet3D_pool(device) -> _Pool:
    """The per-(device, stream) buffer pool shared with Unet3D_with_Conv3D."""
    from .unet3d import pool_for
    return pool_for(device)
