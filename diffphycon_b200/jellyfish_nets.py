"""B200-native jellyfish surrogate networks — SURVEY.md 8(a) row A12 / 8(f) rank 1:

  * `Unet`      drop-in for diffusion/diffusion_2d_jellyfish.py:276-403 (cited as jf.py:line) — the boundary updater
                `bd_updater(bd_0 [N,3,H,W], theta [N]) -> [N,3,H,W]`, conditioned on a FLOAT "time" (the flapping angle);
  * `ForceUnet` drop-in for jf.py:406-481 — `force_model([pressure, boundary] [N,4,H,W]) -> [N,1]`;
  * `JellyfishGuidance` the `design_fn(x, bd_0)` closure of inference/inference_2d_jellyfish.py:276-279 around `force_fn`
                (:85-114): forward through both networks AND the gradient of the guidance objective w.r.t. the state and the
                angle field, which the reference obtains with torch.autograd.grad through the two networks.

Same constructor arguments and `state_dict()` keys / shapes as the reference classes (their bare state_dict checkpoints load
with strict=True).  The module tree only HOLDS parameters; forward and backward run on the kernels of libdpc_b200.so:
activations channels-last [N,H,W,C]; weight-standardised 3x3 convs (jf.py:108-121) are standardised once at pack time and run
as tensor-core implicit GEMMs with fused bias / GroupNorm statistics; every backward contraction (dgrad) is the same conv kernel
with transposed, spatially flipped weights; GroupNorm+SiLU, LayerNorm, linear attention (v / (h*w) variant, jf.py:219), softmax
attention, nearest up-sampling, pixel-unshuffle down-sampling, the mean + Linear(512, out) head and the float time MLP have
hand-written forward and backward kernels (csrc/nets2d.cu).  Only gradients w.r.t. activations and the time conditioning
exist (what the sampler needs) — no weight gradients, no training.  There is no PyTorch / autograd fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib, packing
from .burgers_unet import (Attention, Block, LayerNorm, LinearAttention, PreNorm, Residual, ResnetBlock, SinusoidalPosEmb,
                           Unet3D_pool)
from .unet3d import _require_cuda

HEAD_DIM = 32
WS_EPS = 1e-5   # WeightStandardizedConv2d, fp32 branch (jf.py:114)


def _standardize(w: torch.Tensor) -> torch.Tensor:
    """jf.py:116-119: per-output-channel (w - mean) * rsqrt(var + eps), biased variance."""
    w = w.detach().float()
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + WS_EPS).rsqrt()


class _Tape:
    """Reverse-mode record of one forward pass: closures run in reverse; gradients live in owned device buffers keyed by the
    forward tensor they belong to."""

    def __init__(self):
        self.ops = []
        self.grads: Dict[int, torch.Tensor] = {}
        self.nograd = set()
        self.buffers: List[torch.Tensor] = []
        self.pool = None
        self.out = None
        self.result = {}

    def release(self):
        if self.pool is not None:
            for b in self.buffers:
                self.pool.put(b)
        self.buffers = []
        self.ops = []
        self.grads = {}


class _Net2D(nn.Module):
    """Shared engine of `Unet` and `ForceUnet`: parameter tree, packing, forward (optionally taped) and backward."""

    def _build(self, dim, init_dim, out_dim, dim_mults, channels, groups, with_time, with_ups):
        assert dim % 2 == 0
        self.channels = channels
        self.self_condition = False
        self.dim = dim
        self.heads = 4
        self.groups = groups
        self.with_time, self.with_ups = with_time, with_ups
        init_dim = dim if init_dim is None else init_dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.in_out = in_out
        time_dim = dim * 4
        self.time_dim = time_dim
        tdim = time_dim if with_time else None
        if with_time:
            self.random_or_learned_sinusoidal_cond = False
            self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_dim), nn.GELU(),
                                          nn.Linear(time_dim, time_dim))
        rb = lambda a, b: ResnetBlock(a, b, time_emb_dim=tdim, groups=groups)
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
        self.mid_attn = Residual(PreNorm(mid, Attention(mid)))
        self.mid_block2 = rb(mid, mid)
        if with_ups:
            self.ups = nn.ModuleList([])
            for ind, (d_in, d_out) in enumerate(reversed(in_out)):
                is_last = ind == n - 1
                up = (nn.Sequential(nn.Upsample(scale_factor=2, mode='nearest'), nn.Conv2d(d_out, d_in, 3, padding=1))
                      if not is_last else nn.Conv2d(d_out, d_in, 3, padding=1))
                self.ups.append(nn.ModuleList([rb(d_out + d_in, d_out), rb(d_out + d_in, d_out),
                                               Residual(PreNorm(d_out, LinearAttention(d_out))), up]))
            self.out_dim = channels if out_dim is None else out_dim
            self.final_res_block = rb(dim * 2, dim)
            self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        # engine state
        self.precision = "tf32"      # or "3xtf32" (fp32-class), see Unet3D_with_Conv3D
        self.use_tcgen05 = True
        self._packed = None
        self._packed_key = None
        self._taps: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------------------
    def _resnets(self):
        out = []
        for i, lvl in enumerate(self.downs):
            out += [(f"downs.{i}.0", lvl[0]), (f"downs.{i}.1", lvl[1])]
        out += [("mid_block1", self.mid_block1), ("mid_block2", self.mid_block2)]
        if self.with_ups:
            for i, lvl in enumerate(self.ups):
                out += [(f"ups.{i}.0", lvl[0]), (f"ups.{i}.1", lvl[1])]
            out += [("final_res_block", self.final_res_block)]
        return out

    def invalidate_packed(self):
        """Drop the packed (standardised, TF32-rounded, transposed) weight copies; call after mutating parameters through
        `.data` (in-place edits that do not bump the version counter)."""
        self._packed = None
        self._packed_key = None

    def load_state_dict(self, *a, **k):
        self.invalidate_packed()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def _ensure_packed(self, dev):
        key = (str(dev), self.precision, tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        rnd = self.precision == "tf32"
        P = {}
        f32 = lambda t: t.detach().float().to(dev).contiguous()

        def conv_w(w, cin_pad=None):
            return packing.pack_conv3d(w.to(dev).unsqueeze(2), cin_pad=cin_pad, tf32=rnd)[0]

        def dgrad_w(w, c0=None, c1=None, cout_pad=None):
            """[Cout,Cin,kh,kw] forward weight -> packed weight of the input-gradient conv: [Cin(slice), Cout, kh, kw] flipped."""
            wd = w.to(dev).flip(2, 3).transpose(0, 1)
            if c0 is not None:
                wd = wd[c0:c1]
            if cout_pad is not None and cout_pad != wd.shape[0]:
                wd = torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, 0, 0, cout_pad - wd.shape[0]))
            return packing.pack_conv3d(wd.contiguous().unsqueeze(2), tf32=rnd)[0]

        def lin_t(w2d, c0=None, c1=None):
            """[N,K] forward matrix -> packed [K(slice), N] of the input-gradient GEMM."""
            wt = w2d.to(dev).float().t()
            if c0 is not None:
                wt = wt[c0:c1]
            return packing.pack_linear(wt.contiguous(), tf32=rnd)

        cpad = packing.round_up(self.channels, 4)
        P["cpad"] = cpad
        P["init.w"], P["init.b"] = conv_w(self.init_conv.weight, cpad), f32(self.init_conv.bias)
        P["init.wd"] = dgrad_w(self.init_conv.weight.detach().float(), cout_pad=cpad)     # [cpad, init_dim, 7, 7]
        for name, blk in self._resnets():
            cin = blk.block1.proj.weight.shape[1]
            cout = blk.dim_out
            # ups / final blocks read torch.cat((x, skip), dim=1) (jf.py:390-401): x carries `cout` channels, the skip the rest
            split = cout if cin > cout and name.startswith(("ups", "final")) else None
            for bn in ("block1", "block2"):
                b = getattr(blk, bn)
                ws = _standardize(b.proj.weight.to(dev))
                P[f"{name}.{bn}.w"], P[f"{name}.{bn}.b"] = conv_w(ws), f32(b.proj.bias)
                P[f"{name}.{bn}.gamma"], P[f"{name}.{bn}.beta"] = f32(b.norm.weight), f32(b.norm.bias)
                if bn == "block1" and split is not None:
                    P[f"{name}.{bn}.wd.a"] = dgrad_w(ws, 0, split)
                    P[f"{name}.{bn}.wd.b"] = dgrad_w(ws, split, cin)
                else:
                    P[f"{name}.{bn}.wd"] = dgrad_w(ws)
            if isinstance(blk.res_conv, nn.Conv2d):
                w2 = blk.res_conv.weight.detach().float().reshape(cout, cin)
                P[f"{name}.res.w"] = packing.pack_linear(w2.to(dev), tf32=rnd)
                P[f"{name}.res.b"] = f32(blk.res_conv.bias)
                if split is not None:
                    P[f"{name}.res.wd.a"], P[f"{name}.res.wd.b"] = lin_t(w2, 0, split), lin_t(w2, split, cin)
                else:
                    P[f"{name}.res.wd"] = lin_t(w2)
            P[f"{name}.split"] = split

        def pack_attn(name, res: Residual):
            P[f"{name}.g"] = f32(res.fn.norm.g).reshape(-1)
            att = res.fn.fn
            wq = att.to_qkv.weight.detach().float().reshape(att.to_qkv.weight.shape[0], -1)
            P[f"{name}.qkv.w"], P[f"{name}.qkv.wd"] = packing.pack_linear(wq.to(dev), tf32=rnd), lin_t(wq)
            out = att.to_out[0] if isinstance(att, LinearAttention) else att.to_out
            wo = out.weight.detach().float().reshape(out.weight.shape[0], -1)
            P[f"{name}.out.w"], P[f"{name}.out.b"] = packing.pack_linear(wo.to(dev), tf32=rnd), f32(out.bias)
            P[f"{name}.out.wd"] = lin_t(wo)
            if isinstance(att, LinearAttention):
                P[f"{name}.out.g"] = f32(att.to_out[1].g).reshape(-1)

        for i, lvl in enumerate(self.downs):
            pack_attn(f"downs.{i}.2", lvl[2])
            if isinstance(lvl[3], nn.Sequential):
                w = lvl[3][1].weight.detach().float().to(dev)     # [Cout, 4*C, 1, 1], input channel = c*4 + p1*2 + p2 (jf.py:100-104)
                co, c4 = w.shape[0], w.shape[1]
                c = c4 // 4
                wk = w.reshape(co, c, 4).permute(0, 2, 1).reshape(co, c4)   # k = (p1*2 + p2)*C + c
                P[f"downs.{i}.3.w"] = packing.pack_linear(wk, tf32=rnd)
                P[f"downs.{i}.3.b"] = f32(lvl[3][1].bias)
                for par in range(4):
                    P[f"downs.{i}.3.wd{par}"] = lin_t(wk[:, par * c:(par + 1) * c])
            else:
                P[f"downs.{i}.3.w"], P[f"downs.{i}.3.b"] = conv_w(lvl[3].weight), f32(lvl[3].bias)
                P[f"downs.{i}.3.wd"] = dgrad_w(lvl[3].weight.detach().float())
        pack_attn("mid_attn", self.mid_attn)
        if self.with_ups:
            for i, lvl in enumerate(self.ups):
                pack_attn(f"ups.{i}.2", lvl[2])
                c = lvl[3][1] if isinstance(lvl[3], nn.Sequential) else lvl[3]
                P[f"ups.{i}.3.w"], P[f"ups.{i}.3.b"] = conv_w(c.weight), f32(c.bias)
                P[f"ups.{i}.3.wd"] = dgrad_w(c.weight.detach().float())
            wf = self.final_conv.weight.detach().float().reshape(self.out_dim, self.dim)
            P["final.w"], P["final.b"] = packing.pack_linear(wf.to(dev), tf32=rnd), f32(self.final_conv.bias)
            opad = packing.round_up(self.out_dim, 4)
            wft = torch.zeros(self.dim, opad, device=dev)
            wft[:, :self.out_dim] = wf.to(dev).t()
            P["final.wd"], P["final.opad"] = packing.pack_linear(wft, tf32=rnd), opad
        else:
            P["head.w"], P["head.b"] = f32(self.final.weight), f32(self.final.bias)
        if self.with_time:
            ws, bs, offs, off = [], [], {}, 0
            for name, blk in self._resnets():
                lin = blk.mlp[1]
                ws.append(f32(lin.weight))
                bs.append(f32(lin.bias))
                offs[name] = off
                off += lin.weight.shape[0]
            P["tp.w"], P["tp.b"], P["tp.offs"], P["tp.total"] = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous(), offs, off
            half = self.dim // 2
            P["t.freqs"] = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).float().to(dev).contiguous()
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
    def _run(self, x, time, tape: Optional[_Tape]):
        """Forward pass (jf.py:365-403 / :462-481).  With a tape, every op records its backward closure and no buffer is
        recycled before `_backward` has run."""
        _require_cuda(x)
        N, C, H, W = x.shape
        assert C == self.channels, f"expected {self.channels} channels, got {C}"
        n_lvl = len(self.in_out)
        assert H % (2 ** (n_lvl - 1)) == 0 and W % (2 ** (n_lvl - 1)) == 0
        dev = x.device
        x = x.detach().contiguous().float()
        P = self._ensure_packed(dev)
        pool = Unet3D_pool(dev)
        precise = self.precision == "3xtf32"
        tc = self.use_tcgen05 and not precise
        heads, hid, groups = self.heads, self.heads * HEAD_DIM, self.groups
        local = tape if tape is not None else _Tape()
        local.pool = pool
        rec = tape is not None

        def buf(numel, dtype=torch.float32):
            t = pool.get(numel, dtype)
            local.buffers.append(t)
            return t

        def conv(xa, ca, w, bias, y, cout, h, wd, kh=1, kw=1, stride=1, pad=0, xb=None, cb=0, residual=None, gn=None,
                 out_layout=0, scatter=None):
            p = _lib.ConvParams()
            p.x1, p.x2 = xa.data_ptr(), (xb.data_ptr() if xb is not None else None)
            p.C1, p.C2 = ca, cb
            p.w, p.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
            p.residual = residual.data_ptr() if residual is not None else None
            p.y = y.data_ptr()
            p.gn_stats = gn.data_ptr() if gn is not None else None
            p.gn_groups = groups if gn is not None else 0
            p.B, p.Fi, p.Hi, p.Wi = N, 1, h, wd
            p.Fo, p.Ho, p.Wo = 1, (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
            p.st, p.sh, p.sw = 1, stride, stride
            p.pt, p.ph, p.pw = 0, pad, pad
            p.oh_mul = p.ow_mul = 1
            p.oh_off = p.ow_off = 0
            if scatter is not None:     # output rows land at (2*ho + ph, 2*wo + pw) of a [2*Ho, 2*Wo] frame
                p.oh_mul = p.ow_mul = 2
                p.oh_off, p.ow_off = scatter
            p.Hfull, p.Wfull = p.Ho * p.oh_mul, p.Wo * p.ow_mul
            p.taps, p.ntaps = self._tap(kh, kw, h, wd, dev).data_ptr(), kh * kw
            p.Cout, p.Npad, p.Kpad = cout, w.shape[0], w.shape[1]
            p.out_layout, p.precise = out_layout, (1 if precise else 0)
            _lib.conv(p, tcgen05=tc and scatter is None and out_layout == 0 and stride == 1)

        def grad_add(t, g):
            if id(t) in local.nograd:
                return
            cur = local.grads.get(id(t))
            if cur is None:
                local.grads[id(t)] = g
            else:
                _lib.add(cur, g, cur, g.numel())

        def needs(*ts):
            return any(t is not None and id(t) not in local.nograd for t in ts)

        # ---- time conditioning (jf.py:371, :313-318) ----
        ss, dss, t_emb, tvec = None, None, None, None
        if self.with_time:
            tdim = self.time_dim
            tvec = time.detach().to(device=dev, dtype=torch.float32).contiguous()
            hidden, t_emb = buf(N * tdim), buf(N * tdim)
            _lib.time_embed_f32(tvec, P["t.freqs"], P["t.w1"], P["t.b1"], P["t.w2"], P["t.b2"], hidden, t_emb, N, self.dim)
            ss = buf(N * P["tp.total"])
            _lib.time_proj(t_emb, P["tp.w"], P["tp.b"], ss, N, tdim, P["tp.total"])
            if rec:
                dss = buf(N * P["tp.total"])
                dss.zero_()
        n_gn = 2 * len(self._resnets())
        stats = buf(n_gn * N * groups * 2, torch.float64)
        stats.zero_()
        slot = [0]

        def next_stats():
            s = stats[slot[0] * N * groups * 2:(slot[0] + 1) * N * groups * 2]
            slot[0] += 1
            return s

        def resnet(name, xa, ca, h, wd, cout, xb=None, cb=0):
            m = N * h * wd
            hw = h * wd
            off = P["tp.offs"][name] if self.with_time else 0
            tot = P["tp.total"] if self.with_time else 0
            y1, s1 = buf(m * cout), next_stats()
            conv(xa, ca, P[f"{name}.block1.w"], P[f"{name}.block1.b"], y1, cout, h, wd, 3, 3, 1, 1, xb=xb, cb=cb, gn=s1)
            a1 = buf(m * cout)
            _lib.groupnorm_silu(y1, s1, P[f"{name}.block1.gamma"], P[f"{name}.block1.beta"], ss, tot, off, None, a1, N, hw, cout,
                                groups)
            y2, s2 = buf(m * cout), next_stats()
            conv(a1, cout, P[f"{name}.block2.w"], P[f"{name}.block2.b"], y2, cout, h, wd, 3, 3, 1, 1, gn=s2)
            out = buf(m * cout)
            has_res = f"{name}.res.w" in P
            if has_res:
                res = buf(m * cout)
                conv(xa, ca, P[f"{name}.res.w"], P[f"{name}.res.b"], res, cout, h, wd, xb=xb, cb=cb)
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, res, out, N, hw, cout,
                                    groups)
            else:
                assert xb is None and ca == cout
                _lib.groupnorm_silu(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, xa, out, N, hw, cout,
                                    groups)
            if rec:
                def bwd():
                    dout = local.grads.pop(id(out), None)
                    if dout is None:
                        return
                    sums = buf(N * cout * 2, torch.float64)
                    dy2 = buf(m * cout)
                    _lib.gn_silu_bwd(y2, s2, P[f"{name}.block2.gamma"], P[f"{name}.block2.beta"], None, 0, 0, dout, dy2, sums, None,
                                     N, hw, cout, groups)
                    da1 = buf(m * cout)
                    conv(dy2, cout, P[f"{name}.block2.wd"], None, da1, cout, h, wd, 3, 3, 1, 1)
                    dy1 = buf(m * cout)
                    _lib.gn_silu_bwd(y1, s1, P[f"{name}.block1.gamma"], P[f"{name}.block1.beta"], ss, tot, off, da1, dy1, sums,
                                     dss if self.with_time else None, N, hw, cout, groups)
                    if not needs(xa, xb):
                        return
                    if xb is None:
                        dxa = buf(m * ca)
                        conv(dy1, cout, P[f"{name}.block1.wd"], None, dxa, ca, h, wd, 3, 3, 1, 1)
                        if has_res:
                            dxr = buf(m * ca)
                            conv(dout, cout, P[f"{name}.res.wd"], None, dxr, ca, h, wd, residual=dxa)
                            dxa = dxr
                        else:
                            _lib.add(dxa, dout, dxa, m * ca)
                        grad_add(xa, dxa)
                    else:
                        for src, cs, tag in ((xa, ca, "a"), (xb, cb, "b")):
                            if not needs(src):
                                continue
                            d0 = buf(m * cs)
                            conv(dy1, cout, P[f"{name}.block1.wd.{tag}"], None, d0, cs, h, wd, 3, 3, 1, 1)
                            d1 = buf(m * cs)
                            conv(dout, cout, P[f"{name}.res.wd.{tag}"], None, d1, cs, h, wd, residual=d0)
                            grad_add(src, d1)
                local.ops.append(bwd)
            return out

        def attention(name, xa, c, h, wd, linear):
            m, hw = N * h * wd, h * wd
            xn = buf(m * c)
            _lib.layernorm_channels(xa, P[f"{name}.g"], xn, m, c, use_rsqrt=True)                # PreNorm (jf.py:134-142)
            qkv = buf(m * 3 * hid)
            conv(xn, c, P[f"{name}.qkv.w"], None, qkv, 3 * hid, h, wd)
            att = buf(m * hid)
            y = buf(m * c)
            if linear:
                ctx, kstat = buf(N * heads * HEAD_DIM * HEAD_DIM), buf(N * heads * HEAD_DIM * 2)
                vscale = 1.0 / float(hw)                                                          # jf.py:219
                _lib.spatial_linear_attention_ex(qkv, ctx, kstat, att, N, hw, heads, vscale)
                ypre = buf(m * c)
                conv(att, hid, P[f"{name}.out.w"], P[f"{name}.out.b"], ypre, c, h, wd)
                _lib.layernorm_channels(ypre, P[f"{name}.out.g"], y, m, c, residual=xa, use_rsqrt=True)   # to_out[1] + Residual
            else:
                _lib.spatial_attention(qkv, att, N, hw, heads)                                    # jf.py:241-255
                conv(att, hid, P[f"{name}.out.w"], P[f"{name}.out.b"], y, c, h, wd, residual=xa)
            if rec:
                def bwd():
                    dy = local.grads.pop(id(y), None)
                    if dy is None or not needs(xa):
                        return
                    datt = buf(m * hid)
                    dqkv = buf(m * 3 * hid)
                    if linear:
                        dypre = buf(m * c)
                        _lib.layernorm_channels_bwd(ypre, P[f"{name}.out.g"], dy, None, dypre, m, c)
                        conv(dypre, c, P[f"{name}.out.wd"], None, datt, hid, h, wd)
                        dctx = buf(N * heads * HEAD_DIM * HEAD_DIM)
                        _lib.linattn2d_bwd(qkv, ctx, kstat, datt, dctx, dqkv, N, hw, heads, vscale)
                    else:
                        conv(dy, c, P[f"{name}.out.wd"], None, datt, hid, h, wd)
                        _lib.attention2d_bwd(qkv, att, datt, dqkv, N, hw, heads)
                    dxn = buf(m * c)
                    conv(dqkv, 3 * hid, P[f"{name}.qkv.wd"], None, dxn, c, h, wd)
                    dxa = buf(m * c)
                    _lib.layernorm_channels_bwd(xa, P[f"{name}.g"], dxn, dy, dxa, m, c)
                    grad_add(xa, dxa)
                local.ops.append(bwd)
            return y

        def plain_conv3(name, xa, ca, h, wd, cout):
            y = buf(N * h * wd * cout)
            conv(xa, ca, P[f"{name}.w"], P[f"{name}.b"], y, cout, h, wd, 3, 3, 1, 1)
            if rec:
                def bwd():
                    dy = local.grads.pop(id(y), None)
                    if dy is None or not needs(xa):
                        return
                    dx = buf(N * h * wd * ca)
                    conv(dy, cout, P[f"{name}.wd"], None, dx, ca, h, wd, 3, 3, 1, 1)
                    grad_add(xa, dx)
                local.ops.append(bwd)
            return y

        def downsample(name, xa, ca, h, wd, cout):
            """pixel-unshuffle + 1x1 conv (jf.py:100-104) as one 2x2 stride-2 conv."""
            y = buf(N * (h // 2) * (wd // 2) * cout)
            conv(xa, ca, P[f"{name}.w"], P[f"{name}.b"], y, cout, h, wd, 2, 2, 2, 0)
            if rec:
                def bwd():
                    dy = local.grads.pop(id(y), None)
                    if dy is None or not needs(xa):
                        return
                    dx = buf(N * h * wd * ca)
                    for par in range(4):        # input pixel (2i + p1, 2j + p2) only feeds output (i, j) through tap (p1, p2)
                        conv(dy, cout, P[f"{name}.wd{par}"], None, dx, ca, h // 2, wd // 2, scatter=(par >> 1, par & 1))
                    grad_add(xa, dx)
                local.ops.append(bwd)
            return y

        def upsample(name, xa, ca, h, wd, cout):
            """nearest 2x + 3x3 conv (jf.py:94-98)."""
            up = buf(N * 4 * h * wd * ca)
            _lib.upsample_nearest2x(xa, up, N, h, wd, ca)
            y = buf(N * 4 * h * wd * cout)
            conv(up, ca, P[f"{name}.w"], P[f"{name}.b"], y, cout, 2 * h, 2 * wd, 3, 3, 1, 1)
            if rec:
                def bwd():
                    dy = local.grads.pop(id(y), None)
                    if dy is None or not needs(xa):
                        return
                    dup = buf(N * 4 * h * wd * ca)
                    conv(dy, cout, P[f"{name}.wd"], None, dup, ca, 2 * h, 2 * wd, 3, 3, 1, 1)
                    dx = buf(N * h * wd * ca)
                    _lib.sumpool2x2(dup, dx, N, h, wd, ca)
                    grad_add(xa, dx)
                local.ops.append(bwd)
            return y

        # ---- stem (jf.py:368-369) ----
        cpad = P["cpad"]
        xin = buf(N * H * W * cpad)
        _lib.pack_input(x, xin, N, 1, C, 0, C, H, W, cpad)
        d0 = self.in_out[0][0]
        r = buf(N * H * W * d0)
        conv(xin, cpad, P["init.w"], P["init.b"], r, d0, H, W, 7, 7, 1, 3)
        want_dx = rec and local.result.get("want_input_grad", False)
        if rec and not want_dx:
            local.nograd.add(id(r))
        cur, cur_c, h, w = r, d0, H, W
        skips = []
        for i, (d_in, d_out) in enumerate(self.in_out):
            a = resnet(f"downs.{i}.0", cur, cur_c, h, w, d_in)
            b = resnet(f"downs.{i}.1", a, d_in, h, w, d_in)
            c = attention(f"downs.{i}.2", b, d_in, h, w, True)
            if self.with_ups:
                skips.append((a, d_in))
                skips.append((c, d_in))
            if i < n_lvl - 1:
                e = downsample(f"downs.{i}.3", c, d_in, h, w, d_out)
                h, w = h // 2, w // 2
            else:
                e = plain_conv3(f"downs.{i}.3", c, d_in, h, w, d_out)
            cur, cur_c = e, d_out
        a = resnet("mid_block1", cur, cur_c, h, w, cur_c)
        b = attention("mid_attn", a, cur_c, h, w, False)
        cur = resnet("mid_block2", b, cur_c, h, w, cur_c)
        if not self.with_ups:
            # ForceUnet head (jf.py:478-479)
            O = self.final.weight.shape[0]
            out = torch.empty(N, O, dtype=torch.float32, device=dev)
            _lib.mean_head(cur, P["head.w"], P["head.b"], out, N, h * w, cur_c, O)
            if rec:
                feat, fc, fhw = cur, cur_c, h * w

                def bwd_head():
                    dout = local.result["dout"]
                    dx = buf(N * fhw * fc)
                    _lib.mean_head_bwd(dout, P["head.w"], dx, N, fhw, fc, O)
                    grad_add(feat, dx)
                local.ops.append(bwd_head)
        else:
            for i, (d_in, d_out) in enumerate(reversed(self.in_out)):
                s1, s1c = skips.pop()
                a = resnet(f"ups.{i}.0", cur, cur_c, h, w, d_out, xb=s1, cb=s1c)
                s2, s2c = skips.pop()
                b = resnet(f"ups.{i}.1", a, d_out, h, w, d_out, xb=s2, cb=s2c)
                c = attention(f"ups.{i}.2", b, d_out, h, w, True)
                if i < n_lvl - 1:
                    e = upsample(f"ups.{i}.3", c, d_out, h, w, d_in)
                    h, w = 2 * h, 2 * w
                else:
                    e = plain_conv3(f"ups.{i}.3", c, d_out, h, w, d_in)
                cur, cur_c = e, d_in
            f0 = resnet("final_res_block", cur, cur_c, H, W, self.dim, xb=r, cb=d0)
            out = torch.empty(N, self.out_dim, H, W, dtype=torch.float32, device=dev)
            conv(f0, self.dim, P["final.w"], P["final.b"], out, self.out_dim, H, W, out_layout=1)       # jf.py:403
            if rec:
                def bwd_final():
                    dout = local.result["dout"]                      # [N, out_dim, H, W]
                    opad = P["final.opad"]
                    dcl = buf(N * H * W * opad)
                    _lib.pack_input(dout, dcl, N, 1, self.out_dim, 0, self.out_dim, H, W, opad)
                    df0 = buf(N * H * W * self.dim)
                    conv(dcl, opad, P["final.wd"], None, df0, self.dim, H, W)
                    grad_add(f0, df0)
                local.ops.append(bwd_final)
        if rec:
            local.result.update(dict(N=N, H=H, W=W, r=r, xin_c=cpad, dss=dss, t_emb=t_emb, tvec=tvec, conv=conv, buf=buf, P=P,
                                     d0=d0))
        else:
            local.release()
        return out

    def _backward(self, tape: _Tape, dout: torch.Tensor):
        """Runs the tape in reverse.  Returns (d input [N,C,H,W] or None, d time [N] or None)."""
        with _lib.on_device(dout):
            tape.result["dout"] = dout.detach().contiguous().float()
            for fn in reversed(tape.ops):
                fn()
            R = tape.result
            N, H, W, P = R["N"], R["H"], R["W"], R["P"]
            dx = None
            if R.get("want_input_grad", False):
                dr = tape.grads.pop(id(R["r"]), None)
                assert dr is not None
                dx = torch.empty(N, R["xin_c"], H, W, dtype=torch.float32, device=dout.device)
                R["conv"](dr, R["d0"], P["init.wd"], None, dx, R["xin_c"], H, W, 7, 7, 1, 3, out_layout=1)
                dx = dx[:, :self.channels]
            dt = None
            if self.with_time:
                dt = torch.empty(N, dtype=torch.float32, device=dout.device)
                _lib.time_mlp_bwd(R["tvec"], P["t.freqs"], P["t.w1"], P["t.b1"], P["t.w2"], P["tp.w"], R["t_emb"], R["dss"], dt, N,
                                  self.dim, P["tp.total"])
            tape.release()
            return dx, dt


class Unet(_Net2D):
    """Constructor: jf.py:277-291.  forward(x [N,C,H,W], time [N] float) -> [N,out_dim,H,W] (jf.py:365-403)."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False, random_fourier_features=False,
                 learned_sinusoidal_dim=16):
        super().__init__()
        if self_condition or learned_variance or learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("only the configuration the DiffPhyCon jellyfish runs use is implemented")
        self._build(dim, init_dim, out_dim, dim_mults, channels, resnet_block_groups, with_time=True, with_ups=True)

    @torch.no_grad()
    @_lib.device_guarded
    def forward(self, x, time, x_self_cond=None):
        if x_self_cond is not None:
            raise NotImplementedError("self-conditioning is unused by the DiffPhyCon jellyfish runs")
        return self._run(x, time, None)

    @torch.no_grad()
    @_lib.device_guarded
    def forward_taped(self, x, time, want_input_grad=False):
        """Forward that keeps what `backward` needs.  Returns (out, tape); call `self.backward(tape, d_out)` exactly once."""
        tape = _Tape()
        tape.result["want_input_grad"] = want_input_grad
        return self._run(x, time, tape), tape

    @torch.no_grad()
    def backward(self, tape, dout):
        """(d loss / d x or None, d loss / d time [N]) given d loss / d out [N,out_dim,H,W]."""
        return self._backward(tape, dout)


class ForceUnet(_Net2D):
    """Constructor: jf.py:407-417.  forward(x [N,C,H,W]) -> [N,out_dim] (jf.py:462-481).  `final = nn.Linear(512, out_dim)` is
    hard-coded in the reference (quirk 6): dim * dim_mults[-1] must be 512."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, self_condition=False,
                 resnet_block_groups=8, learned_variance=False):
        super().__init__()
        if self_condition or learned_variance:
            raise NotImplementedError("only the configuration the DiffPhyCon jellyfish runs use is implemented")
        self._build(dim, init_dim, out_dim, dim_mults, channels, resnet_block_groups, with_time=False, with_ups=False)
        self.final = nn.Linear(512, out_dim)
        if dim * dim_mults[-1] != 512:
            raise ValueError("ForceUnet.final is nn.Linear(512, out_dim) (jf.py:454): dim * dim_mults[-1] must be 512")

    @torch.no_grad()
    @_lib.device_guarded
    def forward(self, x, x_self_cond=None):
        if x_self_cond is not None:
            raise NotImplementedError("self-conditioning is unused by the DiffPhyCon jellyfish runs")
        return self._run(x, None, None)

    @torch.no_grad()
    @_lib.device_guarded
    def forward_taped(self, x, want_input_grad=True):
        tape = _Tape()
        tape.result["want_input_grad"] = want_input_grad
        return self._run(x, None, tape), tape

    @torch.no_grad()
    def backward(self, tape, dout):
        """d loss / d x [N,C,H,W] given d loss / d out [N,out_dim]."""
        return self._backward(tape, dout)[0]


class JellyfishGuidance:
    """`design_fn(x, bd_0)` of inference/inference_2d_jellyfish.py:276-279 = force_fn (:85-114) on the engine.

        theta = mean_hw(theta_expand) ; pressure = unnormalize(state[:, :, 2]) ; pred_bd = bd_updater(bd_0, theta)
        force = force_model([pressure, pred_bd]) ; J_b = -mean_t(force * w_t) + reg_ratio * sum_t (theta_{t+1} - theta_t)^2,
        w_t = T, T-1, .., 1 ;  returns cat([dJ/d state, dJ/d theta_expand]) of shape x.

    x: [B,F,4,H,W] (3 state channels + the angle field; with only_vis_pressure=True [B,F,2,H,W]: pressure + angle),
    bd_0: [B,F,3,H,W].  p_min / p_max: the pressure range of normalization_max_min.pkl (:29-32)."""

    def __init__(self, force_model: ForceUnet, bd_updater: Unet, p_min: float, p_max: float, reg_ratio: float,
                 only_vis_pressure: bool = False):
        self.force_model, self.bd_updater = force_model, bd_updater
        self.p_min, self.p_max, self.reg_ratio = float(p_min), float(p_max), float(reg_ratio)
        self.only_vis_pressure = only_vis_pressure

    @torch.no_grad()
    def __call__(self, x, bd_0):
        B, F, C, H, W = x.shape
        ns = 1 if self.only_vis_pressure else 3
        pch = 0 if self.only_vis_pressure else 2
        with _lib.on_device(x):
            x = x.detach().float()
            theta = x[:, :, ns].mean((-1, -2))                                                   # [B,F]  (:93)
            pressure = (0.5 * x[:, :, pch] + 0.5) * (self.p_max - self.p_min) + self.p_min      # unnormalize_state (:35-36)
            bd_flat = bd_0.reshape(B * F, *bd_0.shape[2:]).contiguous().float()
            pred_bd, tape_bd = self.bd_updater.forward_taped(bd_flat, theta.reshape(B * F))      # (:99-102)
            inp = torch.cat((pressure.reshape(B * F, 1, H, W), pred_bd), dim=1).contiguous()     # (:104-105)
            force, tape_f = self.force_model.forward_taped(inp, want_input_grad=True)           # [B*F, 1]  (:106)
            wts = torch.arange(F, 0, -1, dtype=torch.float32, device=x.device)                  # (:108)
            dforce = (-(wts / F)).reshape(1, F).expand(B, F).reshape(B * F, 1).contiguous()     # d(-mean_t(force w))/d force
            dinp = self.force_model.backward(tape_f, dforce)                                    # [B*F,4,H,W]
            _, dtheta = self.bd_updater.backward(tape_bd, dinp[:, 1:4].contiguous())
            dtheta = dtheta.reshape(B, F)
            diff = theta[:, 1:] - theta[:, :-1]                                                 # reg_theta (:47-60)
            dreg = torch.zeros_like(theta)
            dreg[:, 1:] += 2.0 * diff
            dreg[:, :-1] -= 2.0 * diff
            dtheta = dtheta + self.reg_ratio * dreg
            g = torch.zeros_like(x)
            g[:, :, pch] = dinp[:, 0].reshape(B, F, H, W) * (0.5 * (self.p_max - self.p_min))
            g[:, :, ns] = (dtheta / float(H * W)).reshape(B, F, 1, 1).expand(B, F, H, W)
            return g
