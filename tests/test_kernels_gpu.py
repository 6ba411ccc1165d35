"""GPU parity tests, one per C-ABI entry point: the CUDA kernels (called through ctypes, include/dpc_b200.h) against
plain fp32/fp64 PyTorch references of the same op on seeded inputs.  Tolerances:
  - elementwise / normalisation / attention (fp32 SIMT): 2e-5 relative to the output scale
  - TF32 tensor-core contractions: 3e-3 relative to the output scale (10-bit mantissa operands, fp32 accumulate)
  - 3xTF32 contractions: 5e-5 (operand error ~2^-22; the rest is the tensor core's non-IEEE fp32 accumulation over K<=6912)
  - fused sampler step: bit-exact against the reference's recorded step (tests/golden/sampler_step_*.npz)
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffphycon_b200 import _lib, packing
from tests import cpu_emulator as emu

pytestmark = pytest.mark.gpu

DEV = "cuda"
TOL_F32 = 2e-5
TOL_TF32 = 3e-3
TOL_3X = 5e-5


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


def g(seed):
    return torch.Generator().manual_seed(seed)


def run_conv(x1, w_packed, ntaps_kind, *, x2=None, bias=None, residual=None, cout, stride=(1, 1, 1), pad=(0, 0, 0),
             kernel=(1, 1, 1), out_hw=None, up_cls=None, gn_groups=0, out_layout=0, precise=False, tcgen05=False,
             y=None):
    B, Fi, Hi, Wi, C1 = x1.shape
    p = _lib.ConvParams()
    p.x1, p.C1 = x1.data_ptr(), C1
    p.x2, p.C2 = (x2.data_ptr(), x2.shape[-1]) if x2 is not None else (None, 0)
    p.w, p.bias = w_packed.data_ptr(), (bias.data_ptr() if bias is not None else None)
    p.residual = residual.data_ptr() if residual is not None else None
    taps = packing.tap_table(*kernel, Hi, Wi, x1.device)
    p.taps, p.ntaps = taps.data_ptr(), taps.shape[0]
    p.B, p.Fi, p.Hi, p.Wi = B, Fi, Hi, Wi
    p.st, p.sh, p.sw = stride
    p.pt, p.ph, p.pw = pad
    Fo = (Fi + 2 * pad[0] - kernel[0]) // stride[0] + 1
    Ho = (Hi + 2 * pad[1] - kernel[1]) // stride[1] + 1
    Wo = (Wi + 2 * pad[2] - kernel[2]) // stride[2] + 1
    p.oh_mul = p.ow_mul = 1
    p.oh_off = p.ow_off = 0
    if up_cls is not None:
        Fo, Ho, Wo = Fi, Hi, Wi
        p.oh_mul = p.ow_mul = 2
        p.oh_off, p.ow_off = up_cls
    p.Fo, p.Ho, p.Wo = Fo, Ho, Wo
    p.Hfull, p.Wfull = Ho * p.oh_mul, Wo * p.ow_mul
    p.Cout, p.Npad, p.Kpad = cout, w_packed.shape[0], w_packed.shape[1]
    p.out_layout, p.precise = out_layout, int(precise)
    stats = None
    if gn_groups:
        stats = torch.zeros(B, gn_groups, 2, dtype=torch.float64, device=x1.device)
        p.gn_stats, p.gn_groups = stats.data_ptr(), gn_groups
    if y is None:
        shape = (B, Fo, p.Hfull, p.Wfull, cout) if out_layout == 0 else (B, Fo, cout, p.Hfull, p.Wfull)
        y = torch.full(shape, float("nan"), device=x1.device)
    p.y = y.data_ptr()
    ran_tc = _lib.conv(p, tcgen05=tcgen05)
    torch.cuda.synchronize()
    return y, stats, ran_tc


def ncdhw(x_cl):
    return x_cl.permute(0, 4, 1, 2, 3).contiguous()


CONV_CASES = [
    # name, B, F, H, W, C1, C2, Cout, groups
    ("c32", 2, 4, 16, 16, 32, 0, 32, 8),
    ("c64_to_128", 1, 3, 8, 16, 64, 0, 128, 8),
    ("concat_256_to_64", 1, 2, 8, 8, 128, 128, 64, 8),
    ("ragged_rows", 3, 5, 6, 10, 32, 0, 64, 8),       # M = 900: tail tile + tiles straddling samples
    ("c256", 1, 2, 4, 8, 256, 0, 256, 8),
]


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3x3x3_bias_stats(case, precise):
    _, B, Fr, H, W, C1, C2, Cout, groups = case
    gen = g(1)
    x1 = torch.randn(B, Fr, H, W, C1, generator=gen)
    x2 = torch.randn(B, Fr, H, W, C2, generator=gen) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, generator=gen) / (27 * (C1 + C2)) ** 0.5
    bias = torch.randn(Cout, generator=gen)
    xin = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.conv3d(ncdhw(xin).double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1)
    wp, _, _ = packing.pack_conv3d(w.to(DEV), tf32=not precise)
    y, stats, _ = run_conv(x1.to(DEV), wp, 27, x2=x2.to(DEV) if C2 else None, bias=bias.to(DEV), cout=Cout, pad=(1, 1, 1),
                           kernel=(3, 3, 3), gn_groups=groups, precise=precise)
    assert rel_err(y, ref) <= (TOL_3X if precise else TOL_TF32)
    # statistics are those of the values actually written
    v = y.double().reshape(B, -1, groups, Cout // groups)
    assert torch.allclose(stats[:, :, 0], v.sum(dim=(1, 3)), rtol=1e-6, atol=1e-6)
    assert torch.allclose(stats[:, :, 1], (v * v).sum(dim=(1, 3)), rtol=1e-6, atol=1e-6)


def test_conv_stem_7x7x7_padded_channels():
    gen = g(2)
    B, Fr, H, W, C = 2, 4, 16, 16, 6
    x = torch.randn(B, Fr, C, H, W, generator=gen)
    w = torch.randn(64, C, 7, 7, 7, generator=gen) / (343 * C) ** 0.5
    bias = torch.randn(64, generator=gen)
    ref = F.conv3d(x.permute(0, 2, 1, 3, 4).double(), w.double(), bias.double(), padding=3).permute(0, 2, 3, 4, 1)
    xin = torch.empty(B, Fr, H, W, 8, device=DEV)
    _lib.pack_input(x.to(DEV).contiguous(), xin, B, Fr, C, 0, C, H, W, 8)
    wp, _, _ = packing.pack_conv3d(w.to(DEV), cin_pad=8)
    y, _, _ = run_conv(xin, wp, 343, bias=bias.to(DEV), cout=64, pad=(3, 3, 3), kernel=(7, 7, 7))
    assert rel_err(y, ref) <= TOL_TF32


@pytest.mark.parametrize("B,Fr,H,W,C,Cout", [(2, 4, 16, 16, 6, 64), (1, 3, 12, 20, 2, 32), (1, 9, 64, 64, 6, 64), (2, 2, 8, 40, 7, 128)])
def test_stem_conv_tcgen05_sliding_window(B, Fr, H, W, C, Cout):
    """dpc_stem_conv_tcgen05 (7x7x7 init_conv, conv3d.py:392, on no-swizzle overlapping-core-matrix descriptors) against
    F.conv3d in fp64; TF32 operand class.  Cases: 2 planes / 1 plane / the metric frame (9 full tiles per frame, frames past
    both temporal borders) / 7 channels with a ragged padded width."""
    gen = g(21)
    cpad = packing.round_up(C, 4)
    x = torch.randn(B, Fr, C, H, W, generator=gen)
    w = torch.randn(Cout, C, 7, 7, 7, generator=gen) / (343 * C) ** 0.5
    bias = torch.randn(Cout, generator=gen)
    ref = F.conv3d(x.permute(0, 2, 1, 3, 4).double(), w.double(), bias.double(), padding=3).permute(0, 2, 3, 4, 1)
    xin = torch.empty(B, Fr, H, W, cpad, device=DEV)
    _lib.pack_input(x.to(DEV).contiguous(), xin, B, Fr, C, 0, C, H, W, cpad)
    ws = packing.pack_stem_conv(w.to(DEV), cpad)
    y = torch.full((B, Fr, H, W, Cout), float("nan"), device=DEV)
    assert _lib.stem_conv(xin, ws, bias.to(DEV), y, B, Fr, H, W, cpad, Cout, 7, 7, 7)
    torch.cuda.synchronize()
    assert rel_err(y, ref) <= TOL_TF32


@pytest.mark.parametrize("BF,HW,Cout,bias", [(3, 64, 6, True), (2, 4096, 2, True), (5, 100, 4, False), (1, 33, 6, True)])
def test_final_proj_reference_layout(BF, HW, Cout, bias):
    """dpc_final_proj (final_conv[1], conv3d.py:427) against fp64: fp32 FMAs, output in the reference layout [BF, Cout, HW]."""
    gen = g(51)
    x = torch.randn(BF * HW, 64, generator=gen)
    w = torch.randn(Cout, 64, generator=gen) / 8
    b = torch.randn(Cout, generator=gen) if bias else None
    ref = x.double() @ w.double().t() + (b.double() if bias else 0)
    ref = ref.reshape(BF, HW, Cout).permute(0, 2, 1)
    out = torch.full((BF, Cout, HW), float("nan"), device=DEV)
    assert _lib.final_proj(x.to(DEV), w.to(DEV), b.to(DEV) if bias else None, out, BF, HW, 64, Cout)
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL_F32


def test_pack_input_slice():
    gen = g(3)
    x = torch.randn(2, 3, 6, 8, 12, generator=gen)
    out = torch.empty(2, 3, 8, 12, 4, device=DEV)
    _lib.pack_input(x.to(DEV), out, 2, 3, 6, 3, 2, 8, 12, 4)
    ref = torch.zeros(2, 3, 8, 12, 4)
    ref[..., :2] = x[:, :, 3:5].permute(0, 1, 3, 4, 2)
    assert torch.equal(out.cpu(), ref)


def test_conv_down_1x4x4_stride2():
    gen = g(4)
    x = torch.randn(2, 3, 16, 16, 64, generator=gen)
    w = torch.randn(64, 64, 1, 4, 4, generator=gen) / (16 * 64) ** 0.5
    bias = torch.randn(64, generator=gen)
    ref = F.conv3d(ncdhw(x).double(), w.double(), bias.double(), stride=(1, 2, 2), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    wp, _, _ = packing.pack_conv3d(w.to(DEV))
    y, _, _ = run_conv(x.to(DEV), wp, 16, bias=bias.to(DEV), cout=64, stride=(1, 2, 2), pad=(0, 1, 1), kernel=(1, 4, 4))
    assert y.shape == (2, 3, 8, 8, 64)
    assert rel_err(y, ref) <= TOL_TF32


def test_conv_transpose_1x4x4_four_parity_classes():
    gen = g(5)
    x = torch.randn(2, 3, 8, 8, 64, generator=gen)
    w = torch.randn(64, 64, 1, 4, 4, generator=gen) / (4 * 64) ** 0.5
    bias = torch.randn(64, generator=gen)
    ref = F.conv_transpose3d(ncdhw(x).double(), w.double(), bias.double(), stride=(1, 2, 2),
                             padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    y = torch.full((2, 3, 16, 16, 64), float("nan"), device=DEV)
    xd = x.to(DEV)
    for cls, wp in packing.pack_conv_transpose_1x4x4(w.to(DEV)).items():
        run_conv(xd, wp, 4, bias=bias.to(DEV), cout=64, pad=(0, 1 - cls[0], 1 - cls[1]), kernel=(1, 2, 2), up_cls=cls, y=y)
    assert rel_err(y, ref) <= TOL_TF32


@pytest.mark.parametrize("B,Fr,H,W,C", [(2, 3, 16, 16, 64), (1, 2, 64, 64, 64), (1, 3, 32, 32, 128), (2, 1, 12, 40, 64)])
def test_conv_down_tcgen05_parity_boxes(B, Fr, H, W, C):
    """1x4x4 stride-2 down-conv (conv3d.py:163) on the tcgen05 kernel: four parity boxes fetched with TMA traversal stride 2,
    2x2 taps each, against F.conv3d in fp64 (TF32 operand class)."""
    gen = g(41)
    x = torch.randn(B, Fr, H, W, C, generator=gen)
    w = torch.randn(C, C, 1, 4, 4, generator=gen) / (16 * C) ** 0.5
    bias = torch.randn(C, generator=gen)
    ref = F.conv3d(ncdhw(x).double(), w.double(), bias.double(), stride=(1, 2, 2), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    wp, _, _ = packing.pack_conv3d(w.to(DEV))
    y, _, ran_tc = run_conv(x.to(DEV), wp, 16, bias=bias.to(DEV), cout=C, stride=(1, 2, 2), pad=(0, 1, 1), kernel=(1, 4, 4),
                            tcgen05=True)
    assert ran_tc, "shape should be served by the tcgen05 kernel"
    assert y.shape == (B, Fr, H // 2, W // 2, C)
    assert rel_err(y, ref) <= TOL_TF32


@pytest.mark.parametrize("B,Fr,H,W,C", [(2, 3, 8, 8, 64), (1, 2, 32, 32, 64), (1, 3, 16, 16, 128), (2, 1, 6, 20, 64)])
def test_conv_transpose_tcgen05_parity_classes(B, Fr, H, W, C):
    """ConvTranspose3d 1x4x4 stride 2 (conv3d.py:160) as four 2x2-tap parity classes on the tcgen05 kernel (strided output)."""
    gen = g(42)
    x = torch.randn(B, Fr, H, W, C, generator=gen)
    w = torch.randn(C, C, 1, 4, 4, generator=gen) / (4 * C) ** 0.5
    bias = torch.randn(C, generator=gen)
    ref = F.conv_transpose3d(ncdhw(x).double(), w.double(), bias.double(), stride=(1, 2, 2),
                             padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    y = torch.full((B, Fr, 2 * H, 2 * W, C), float("nan"), device=DEV)
    xd = x.to(DEV)
    for cls, wp in packing.pack_conv_transpose_1x4x4(w.to(DEV)).items():
        _, _, ran_tc = run_conv(xd, wp, 4, bias=bias.to(DEV), cout=C, pad=(0, 1 - cls[0], 1 - cls[1]), kernel=(1, 2, 2),
                                up_cls=cls, y=y, tcgen05=True)
        assert ran_tc, "shape should be served by the tcgen05 kernel"
    assert rel_err(y, ref) <= TOL_TF32


def test_linear_residual_and_reference_layout_output():
    gen = g(6)
    x = torch.randn(2, 3, 8, 8, 128, generator=gen)
    w = torch.randn(64, 128, generator=gen) / 128 ** 0.5
    bias = torch.randn(64, generator=gen)
    res = torch.randn(2, 3, 8, 8, 64, generator=gen)
    ref = x.double() @ w.double().t() + bias.double() + res.double()
    y, _, _ = run_conv(x.to(DEV), packing.pack_linear(w.to(DEV)), 1, bias=bias.to(DEV), residual=res.to(DEV), cout=64)
    assert rel_err(y, ref) <= TOL_TF32
    # Cout = 6 written straight into the reference layout [B,F,C,H,W]
    w6 = torch.randn(6, 128, generator=gen) / 128 ** 0.5
    b6 = torch.randn(6, generator=gen)
    ref6 = (x.double() @ w6.double().t() + b6.double()).permute(0, 1, 4, 2, 3)
    y6, _, _ = run_conv(x.to(DEV), packing.pack_linear(w6.to(DEV)), 1, bias=b6.to(DEV), cout=6, out_layout=1)
    assert y6.shape == (2, 3, 6, 8, 8)
    assert rel_err(y6, ref6) <= TOL_TF32
    # qkv-shaped projection (N = 384 -> three 128-wide tiles), near-fp32 mode
    w3 = torch.randn(384, 128, generator=gen) / 128 ** 0.5
    ref3 = x.double() @ w3.double().t()
    y3, _, _ = run_conv(x.to(DEV), packing.pack_linear(w3.to(DEV), tf32=False), 1, cout=384, precise=True)
    assert rel_err(y3, ref3) <= TOL_3X


@pytest.mark.parametrize("with_ss,with_res", [(True, False), (False, True), (False, False)])
def test_groupnorm_silu(with_ss, with_res):
    gen = g(7)
    B, rps, C, G = 3, 200, 64, 8
    y = torch.randn(B, rps, C, generator=gen) * 2 + 0.5
    gamma, beta = torch.randn(C, generator=gen), torch.randn(C, generator=gen)
    ss = torch.randn(B, 4 * C, generator=gen) if with_ss else None
    res = torch.randn(B, rps, C, generator=gen) if with_res else None
    ref = F.group_norm(y.permute(0, 2, 1).double(), G, gamma.double(), beta.double(), eps=1e-5).permute(0, 2, 1)
    if with_ss:
        ref = ref * (ss[:, None, C:2 * C].double() + 1) + ss[:, None, 2 * C:3 * C].double()
    ref = F.silu(ref)
    if with_res:
        ref = ref + res.double()
    v = y.double().reshape(B, rps, G, C // G)
    stats = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous().to(DEV)
    out = torch.empty(B, rps, C, device=DEV)
    _lib.groupnorm_silu(y.to(DEV), stats, gamma.to(DEV), beta.to(DEV), ss.to(DEV) if with_ss else None, 4 * C, C,
                        res.to(DEV) if with_res else None, out, B, rps, C, G)
    assert rel_err(out, ref) <= TOL_F32


@pytest.mark.parametrize("C", [32, 64, 128, 256])
def test_layernorm_channels(C):
    gen = g(8)
    x = torch.randn(777, C, generator=gen) * 3 + 1
    gamma = torch.randn(C, generator=gen)
    out = torch.empty(777, C, device=DEV)
    _lib.layernorm_channels(x.to(DEV), gamma.to(DEV), out, 777, C)
    ref = torch.empty(777 * C)
    emu.layernorm_channels(x.reshape(-1).double(), gamma.double(), ref, 777, C)
    assert rel_err(out.reshape(-1), ref) <= TOL_F32


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("Fr", [4, 20, 32, 40, 64])
def test_temporal_attention(Fr, precise):
    """Tensor-core kernel for every F <= 64 (frames padded to 32 / 64, padded keys masked): 3xTF32 = fp32 class, or TF32 operands."""
    gen = g(9)
    B, HW, heads = 2, 24, 4
    qkv = torch.randn(B, Fr, HW, 384, generator=gen)
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    ang = (torch.arange(Fr).float()[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
    bias = torch.randn(heads, Fr, Fr, generator=gen)
    tol = TOL_F32 if precise else 2e-3
    out = torch.empty(B, Fr, HW, 128, device=DEV)
    _lib.temporal_attention(qkv.to(DEV), ang.cos().to(DEV), ang.sin().to(DEV), bias.to(DEV), out, B, Fr, HW, heads, True,
                            precise)
    ref = torch.empty(out.numel(), dtype=torch.float64)
    emu.temporal_attention(qkv.reshape(-1).double(), ang.cos().double(), ang.sin().double(), bias.double(), ref, B, Fr, HW,
                           heads, True)
    assert rel_err(out.reshape(-1), ref) <= tol
    out2 = torch.empty_like(out)
    _lib.temporal_attention(qkv.to(DEV), None, None, None, out2, B, Fr, HW, heads, False, precise)
    emu.temporal_attention(qkv.reshape(-1).double(), None, None, None, ref, B, Fr, HW, heads, False)
    assert rel_err(out2.reshape(-1), ref) <= tol
    # a relative bias (function of j - i, like RelativePositionBias) through the shared-memory table: same result as the general form
    tab = torch.randn(heads, 2 * Fr - 1, generator=gen)
    idx = torch.arange(Fr)[None, :] - torch.arange(Fr)[:, None] + Fr - 1
    rbias = tab[:, idx].contiguous()
    out3, out4 = torch.empty_like(out), torch.empty_like(out)
    _lib.temporal_attention(qkv.to(DEV), ang.cos().to(DEV), ang.sin().to(DEV), rbias.to(DEV), out3, B, Fr, HW, heads, True, precise,
                            relative_bias=True)
    _lib.temporal_attention(qkv.to(DEV), ang.cos().to(DEV), ang.sin().to(DEV), rbias.to(DEV), out4, B, Fr, HW, heads, True, precise)
    assert torch.equal(out3, out4)


@pytest.mark.parametrize("HW", [16, 100, 256])
def test_spatial_attention(HW):
    gen = g(10)
    BF, heads = 5, 4
    qkv = torch.randn(BF, HW, 384, generator=gen) * 1.5
    out = torch.empty(BF, HW, 128, device=DEV)
    _lib.spatial_attention(qkv.to(DEV), out, BF, HW, heads)
    ref = torch.empty(out.numel(), dtype=torch.float64)
    emu.spatial_attention(qkv.reshape(-1).double(), ref, BF, HW, heads)
    assert rel_err(out.reshape(-1), ref) <= TOL_F32


@pytest.mark.parametrize("HW,scale", [(64, 1.5), (256, 1.5), (256, 4.0), (512, 1.0)])
def test_spatial_attention_tensor_cores(HW, scale):
    """dpc_spatial_attention_mma (TF32 mma.sync q k^T and P v, online softmax over 64-key blocks) against the fp64 emulator;
    TF32 operand class.  scale 4 makes the softmax peaked (running-max rescaling matters); HW 512 = two query blocks."""
    gen = g(11)
    BF, heads = 3, 4
    qkv = torch.randn(BF, HW, 384, generator=gen) * scale
    out = torch.full((BF, HW, 128), float("nan"), device=DEV)
    _lib.spatial_attention(qkv.to(DEV), out, BF, HW, heads, precise=False)
    ref = torch.empty(out.numel(), dtype=torch.float64)
    emu.spatial_attention(qkv.reshape(-1).double(), ref, BF, HW, heads)
    assert rel_err(out.reshape(-1), ref) <= TOL_TF32 * (2 if scale > 2 else 1)


@pytest.mark.parametrize("HW", [16, 100, 1024])
def test_spatial_linear_attention(HW):
    gen = g(11)
    BF, heads = 3, 4
    qkv = torch.randn(BF, HW, 384, generator=gen) * 1.5
    out = torch.empty(BF, HW, 128, device=DEV)
    ctx = torch.empty(BF * heads * 32 * 32, device=DEV)
    _lib.spatial_linear_attention(qkv.to(DEV), ctx, out, BF, HW, heads)
    ref = torch.empty(out.numel(), dtype=torch.float64)
    emu.spatial_linear_attention(qkv.reshape(-1).double(), None, ref, BF, HW, heads)
    assert rel_err(out.reshape(-1), ref) <= TOL_F32


def test_time_embedding():
    gen = g(12)
    B, dim = 5, 64
    t = torch.tensor([0, 1, 499, 998, 999])
    half = dim // 2
    freqs = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1))).float()
    w1, b1 = torch.randn(256, dim, generator=gen) / 8, torch.randn(256, generator=gen)
    w2, b2 = torch.randn(256, 256, generator=gen) / 16, torch.randn(256, generator=gen)
    W, bb = torch.randn(1000, 256, generator=gen) / 16, torch.randn(1000, generator=gen)
    hidden, t_emb, out = torch.empty(B * 256, device=DEV), torch.empty(B * 256, device=DEV), torch.empty(B * 1000, device=DEV)
    _lib.time_embed(t.to(DEV), freqs.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), hidden, t_emb, B, dim)
    _lib.time_proj(t_emb, W.to(DEV), bb.to(DEV), out, B, 256, 1000)
    r_emb, r_out = torch.empty(B * 256), torch.empty(B * 1000)
    emu.time_embed(t, freqs, w1, b1, w2, b2, None, r_emb, B, dim)
    emu.time_proj(r_emb, W, bb, r_out, B, 256, 1000)
    assert rel_err(t_emb, r_emb) <= 5e-5  # sinf/cosf of arguments up to ~1e3 rad
    assert rel_err(out, r_out) <= 5e-5


@pytest.mark.parametrize("tag,guidance", [("std", "standard"), ("alpha", "standard-alpha")])
def test_fused_ddpm_step_bit_exact_vs_reference(tag, guidance, golden_dir):
    """x_{t-1} and x_start of the reference's own p_sample (recorded eps / noise), through the C-ABI."""
    import diffphycon_b200 as dpc
    z = np.load(os.path.join(golden_dir, f"sampler_step_{tag}.npz"))
    mj = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=2)
    d = dpc.GaussianDiffusion([mj, mw], image_size=16, frames=4, timesteps=1000, eval_2ddpm=True,
                              standard_fixed_ratio=float(z["standard_fixed_ratio"]), coeff_ratio=float(z["coeff_ratio"]),
                              w_prob_exp=float(z["w_prob_exp"]))
    fn = dpc.StockSmokeGuidance(w_energy=float(z["w_energy"]))
    init = torch.from_numpy(z["init"]).to(DEV)
    for t in (999, 500, 1, 0):
        gt = lambda k: torch.from_numpy(z[f"t{t}/{k}"])
        d._eps = lambda x, tt, gt=gt: (gt("eps_joint").to(DEV), gt("eps_w").to(DEV))
        d.sample_noise = lambda shape, device, gt=gt: gt("z").to(DEV)
        pred, x_start = d.p_sample(gt("x").shape, gt("x").to(DEV), t, design_fn=fn, design_guidance=guidance, init=init,
                                   _impose_init=True)
        assert torch.equal(x_start.cpu(), gt("x_start")), t
        assert torch.equal(pred.cpu(), gt("pred")), t
        # user-callable design_fn path (autograd on the device) agrees with the fused closed form
        pred2, _ = d.p_sample(gt("x").shape, gt("x").to(DEV), t, design_fn=lambda x, low=None, init=None, init_u=None: fn(x),
                              design_guidance=guidance, init=init, _impose_init=True)
        assert (pred2 - pred).abs().max().item() <= 1e-6


def test_fused_ddim_step_matches_oracle():
    from oracle import smoke_sampler_oracle as so
    gen = g(13)
    B, Fr, S = 2, 4, 16
    sched = so.make_schedule(1000, "sigmoid")
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    init = torch.rand(B, S, S, generator=gen)
    import diffphycon_b200 as dpc
    mj = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=2)
    d = dpc.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=1000, sampling_timesteps=3,
                              ddim_sampling_eta=1.0, standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True,
                              w_prob_exp=0.97)
    s = d._sched()
    for time, time_next in so.ddim_times(1000, 3):
        x = torch.randn(B, Fr, 6, S, S, generator=gen)
        ej = torch.randn(B, Fr, 6, S, S, generator=gen)
        ew = torch.randn(B, Fr, 2, S, S, generator=gen)
        nz = torch.randn(B, Fr, 6, S, S, generator=gen)
        ref, _ = so.ddim_step(sched, x, time, time_next, ej, ew, nz, init, lambda v: so.guidance_fn(v, R, 0.25), eta=1.0,
                              design_guidance="standard", standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
        c = d._coefs(time, dpc.StockSmokeGuidance(w_energy=0.25), "standard")
        if time_next < 0:
            c.last = 1
        else:
            a, an = s['alphas_cumprod'][time], s['alphas_cumprod'][time_next]
            sigma = 1.0 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
            c.sqrt_alpha_next, c.c, c.ddim_sigma = float(an.sqrt()), float((1 - an - sigma ** 2).sqrt()), float(sigma)
        out = torch.empty(B, Fr, 6, S, S, device=DEV)
        _lib.guided_step(True, x.to(DEV), ej.to(DEV), ew.to(DEV), nz.to(DEV), init.to(DEV), None, c, out, None, B, Fr, S, S)
        assert torch.equal(out.cpu(), ref), (time, (out.cpu() - ref).abs().max())


TC_CASES = [
    # name, B, F, H, W, C1, C2, Cout
    ("w16_64", 2, 4, 16, 16, 64, 0, 64),
    ("w64_64", 1, 3, 64, 64, 64, 0, 64),
    ("w32_64_to_128", 2, 3, 32, 32, 64, 0, 128),
    ("w32_concat_128", 1, 2, 32, 32, 128, 128, 128),
    ("w16_256", 1, 3, 16, 16, 256, 0, 256),
    ("w16_concat_512_to_128", 1, 2, 16, 16, 256, 256, 128),
    ("partial_tile_h40", 1, 2, 40, 16, 64, 0, 64),
    ("w64_128_to_64_concat", 1, 2, 64, 64, 64, 64, 64),
    ("w8_128", 2, 4, 8, 8, 128, 0, 128),
    ("w4_256", 1, 6, 4, 4, 256, 0, 256),
    ("nonsquare_h12_w20", 2, 3, 12, 20, 64, 0, 64),
    # quad mode (Cout = 64, W >= 32, whole frame quads, even number of quads): stacked temporal taps, N = 64/128/192 MMAs
    ("quad_w32_64", 1, 8, 32, 32, 64, 0, 64),
    ("quad_w64_concat_128_to_64", 2, 4, 64, 64, 64, 64, 64),
    ("quad_h40_w36_ragged", 1, 8, 40, 36, 64, 0, 64),
    ("quad_f16_w32_256_to_64", 1, 16, 8, 32, 256, 0, 64),
]


@pytest.mark.parametrize("pair", [True, False, "quad", "duo"], ids=["cta_pair", "single_cta", "cta_pair_quad", "cta_pair_duo"])
@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv3d_tcgen05(case, pair, monkeypatch):
    """The TMA/tcgen05 kernel against fp64 conv3d and against the generic tensor-core kernel (same numerics class).
    cta_pair: cta_group::2 clusters (the default when B*F is even); single_cta: DPC_TC_PAIR=0 forces one CTA per tile;
    cta_pair_quad: DPC_TC_QUAD=1, four output frames per tile with the temporal taps stacked into N = 64/128/192 MMAs."""
    _, B, Fr, H, W, C1, C2, Cout = case
    monkeypatch.setenv("DPC_TC_PAIR", "1" if pair else "0")
    monkeypatch.setenv("DPC_TC_QUAD", {"quad": "4", "duo": "2"}.get(pair, "0"))   # stacked-temporal-tap modes (Cout = 64 shapes)
    gen = g(21)
    x1 = torch.randn(B, Fr, H, W, C1, generator=gen)
    x2 = torch.randn(B, Fr, H, W, C2, generator=gen) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, generator=gen) / (27 * (C1 + C2)) ** 0.5
    bias = torch.randn(Cout, generator=gen)
    xin = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.conv3d(ncdhw(xin).double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1)
    wp, _, _ = packing.pack_conv3d(w.to(DEV))
    kw = dict(x2=x2.to(DEV) if C2 else None, bias=bias.to(DEV), cout=Cout, pad=(1, 1, 1), kernel=(3, 3, 3), gn_groups=8)
    y, stats, ran_tc = run_conv(x1.to(DEV), wp, 27, tcgen05=True, **kw)
    assert ran_tc, "shape should be served by the tcgen05 kernel"
    assert rel_err(y, ref) <= TOL_TF32
    y2, stats2, _ = run_conv(x1.to(DEV), wp, 27, tcgen05=False, **kw)
    assert rel_err(y, y2) <= 2e-4        # both truncate the same TF32 operands; only the accumulation order differs
    # per-thread partial sums are fp32 (<= 64 values), everything above is double: 2e-6 of the L1 mass
    v = y.double().reshape(B, -1, 8, Cout // 8)
    assert ((stats[:, :, 0] - v.sum(dim=(1, 3))).abs() <= 2e-6 * v.abs().sum(dim=(1, 3)) + 1e-9).all()
    assert ((stats[:, :, 1] - (v * v).sum(dim=(1, 3))).abs() <= 2e-6 * (v * v).sum(dim=(1, 3)) + 1e-9).all()


TC_GEMM_CASES = [
    # name, B, F, H, W, C1, C2, Cout, bias, residual
    ("qkv_64_to_384", 2, 4, 16, 16, 64, 0, 384, False, False),
    ("qkv_256_to_384_w8", 1, 6, 8, 8, 256, 0, 384, False, False),
    ("out_128_to_64_res", 2, 3, 32, 32, 128, 0, 64, True, True),
    ("out_128_to_256_res", 1, 4, 16, 16, 128, 0, 256, False, True),
    ("res_concat_512_to_128", 1, 2, 16, 16, 256, 256, 128, True, False),
    ("w64_64_to_384", 1, 2, 64, 64, 64, 0, 384, False, False),
    ("w4_128_to_128", 1, 5, 4, 4, 128, 0, 128, True, True),
]


@pytest.mark.parametrize("case", TC_GEMM_CASES, ids=[c[0] for c in TC_GEMM_CASES])
def test_linear_tcgen05(case):
    """1x1x1 conv / Linear layers through the persistent TMA/tcgen05 kernel (column tiles of 64/128/256)."""
    _, B, Fr, H, W, C1, C2, Cout, with_bias, with_res = case
    gen = g(31)
    x1 = torch.randn(B, Fr, H, W, C1, generator=gen)
    x2 = torch.randn(B, Fr, H, W, C2, generator=gen) if C2 else None
    w = torch.randn(Cout, C1 + C2, generator=gen) / (C1 + C2) ** 0.5
    bias = torch.randn(Cout, generator=gen) if with_bias else None
    res = torch.randn(B, Fr, H, W, Cout, generator=gen) if with_res else None
    xin = torch.cat([x1, x2], -1) if C2 else x1
    ref = xin.double() @ w.double().t()
    if with_bias:
        ref = ref + bias.double()
    if with_res:
        ref = ref + res.double()
    y, _, ran_tc = run_conv(x1.to(DEV), packing.pack_linear(w.to(DEV)), 1, x2=x2.to(DEV) if C2 else None,
                            bias=bias.to(DEV) if with_bias else None, residual=res.to(DEV) if with_res else None,
                            cout=Cout, tcgen05=True)
    assert ran_tc, "shape should be served by the tcgen05 kernel"
    assert rel_err(y, ref) <= TOL_TF32


@pytest.mark.parametrize("B,Fr,S,C1,C2,Cout", [(2, 4, 16, 128, 0, 64), (3, 2, 32, 64, 64, 64), (1, 4, 16, 256, 256, 128), (2, 3, 8, 64, 0, 128)])
def test_linear_tcgen05_folded_groupnorm_residual(B, Fr, S, C1, C2, Cout):
    """ResnetBlock tail in one pass (conv3d.py:229-230): out = silu(GroupNorm(y2)) + res_conv(x).  dpc_gn_fold turns the conv
    epilogue's statistics into per-(sample, channel) coefficients; the 1x1x1 GEMM applies them to its residual operand."""
    gen = g(61)
    x1 = torch.randn(B, Fr, S, S, C1, generator=gen)
    x2 = torch.randn(B, Fr, S, S, C2, generator=gen) if C2 else None
    y2 = torch.randn(B, Fr, S, S, Cout, generator=gen) * 1.7 + 0.4
    w = torch.randn(Cout, C1 + C2, generator=gen) / (C1 + C2) ** 0.5
    bias, gamma, beta = (torch.randn(Cout, generator=gen) for _ in range(3))
    xin = torch.cat([x1, x2], -1) if C2 else x1
    gn = F.group_norm(y2.reshape(B, -1, Cout).permute(0, 2, 1).double(), 8, gamma.double(), beta.double(), eps=1e-5)
    ref = F.silu(gn).permute(0, 2, 1).reshape(B, Fr, S, S, Cout) + xin.double() @ w.double().t() + bias.double()
    v = y2.double().reshape(B, -1, 8, Cout // 8)
    stats = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous().to(DEV)
    a, d = torch.empty(B, Cout, device=DEV), torch.empty(B, Cout, device=DEV)
    _lib.gn_fold(stats, gamma.to(DEV), beta.to(DEV), a, d, B, Fr * S * S, Cout, 8)
    p = _lib.ConvParams()
    x1d, y2d, wp, bd = x1.to(DEV), y2.to(DEV), packing.pack_linear(w.to(DEV)), bias.to(DEV)
    x2d = x2.to(DEV) if C2 else None
    taps = packing.tap_table(1, 1, 1, S, S, DEV)
    out = torch.full((B, Fr, S, S, Cout), float("nan"), device=DEV)
    p.x1, p.C1, p.x2, p.C2 = x1d.data_ptr(), C1, (x2d.data_ptr() if C2 else None), C2
    p.w, p.bias, p.residual, p.y = wp.data_ptr(), bd.data_ptr(), y2d.data_ptr(), out.data_ptr()
    p.res_scale, p.res_shift = a.data_ptr(), d.data_ptr()
    p.taps, p.ntaps = taps.data_ptr(), 1
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = B, Fr, S, S, Fr, S, S
    p.st = p.sh = p.sw = 1
    p.oh_mul = p.ow_mul = 1
    p.Hfull, p.Wfull = S, S
    p.Cout, p.Npad, p.Kpad = Cout, wp.shape[0], wp.shape[1]
    assert _lib.conv(p, tcgen05=True, tc_only=True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL_TF32


@pytest.mark.parametrize("ddim", [False, True])
@pytest.mark.parametrize("guided", [False, True])
def test_jellyfish_step_kernels_match_torch_restatement(ddim, guided):
    """dpc_jelly_x_start / dpc_jelly_step / dpc_jelly_write_bd against the torch restatement of jellyfish.py:744-806, :858-876
    (tests/cpu_emulator.py, itself pinned to the reference traces in test_jellyfish_sampler.py): elementwise results
    bit-exact, the theta means to 1e-6 (reduction order)."""
    torch.manual_seed(11)
    B, Fr, H, W, cs = 3, 5, 12, 20, 1
    x = torch.randn(B, Fr, 7, H, W)
    eps, eps_w = torch.randn(B, Fr, 4, H, W), torch.randn(B, Fr, 1, H, W)
    g = torch.randn(B, Fr, 4, H, W) if guided else None
    noise = torch.randn(B, Fr, 4, H, W)
    state_0, bd_0, th0 = torch.randn(B, 3, H, W), torch.randn(B, 3, H, W), torch.rand(B)
    pred_bd = torch.randn(B * Fr, 3, H, W)
    sc = dict(ga=0.37, gb=0.21, c1=0.83, c2=0.41, sigma=0.6)
    outs = {}
    for where in ("cpu", "cuda"):
        mod = emu if where == "cpu" else _lib
        mv = lambda t: None if t is None else t.to(where)
        xs = torch.empty(B, Fr, 4, H, W, device=where)
        mod.jelly_x_start(mv(x), mv(eps), xs, 1.7, 0.9, True)
        x_next, x_w = torch.zeros(B, Fr, 7, H, W, device=where), torch.zeros(B, Fr, 7, H, W, device=where)
        dth, thm = torch.empty(B, Fr, device=where), torch.empty(B, Fr, device=where)
        mod.jelly_step(mv(x), xs, mv(eps), mv(eps_w), mv(g), mv(noise), mv(state_0), mv(th0), x_next, x_w, dth, thm,
                       sc["ga"], sc["gb"], sc["c1"], sc["c2"], sc["sigma"], ddim, cs)
        mod.jelly_write_bd(mv(pred_bd), mv(bd_0), x_next, x_w, cs)
        outs[where] = [t.cpu() for t in (xs, x_next, x_w, dth, thm)]
    for a, b in zip(outs["cpu"][:3], outs["cuda"][:3]):
        assert torch.equal(a, b)
    for a, b in zip(outs["cpu"][3:], outs["cuda"][3:]):
        assert (a - b).abs().max().item() <= 1e-6


@pytest.mark.parametrize("Fr", [32, 20, 5])
@pytest.mark.parametrize("B,HW", [(1, 4), (2, 64), (3, 148 * 4 + 8)])
def test_temporal_block_fused_tcgen05(B, HW, Fr):
    """dpc_temporal_block_fused (LayerNorm + to_qkv + RoPE/bias attention + to_out + residual, one launch) against the fp32
    restatement of conv3d.py:165-174, :293-352 (tests/cpu_emulator.py).  The two projections are TF32 contractions:
    tolerance 3e-3 of the output scale; the attention itself is fp32.  Fr < 32 (e.g. the jellyfish configuration's 20 frames): the
    tile still spans 32 frame rows per pixel, the missing ones are TMA zero fill, masked as keys and clipped by the store."""
    torch.manual_seed(5)
    Cn, heads = 64, 4
    x = torch.randn(B * Fr * HW * Cn) * 1.5 + 0.3
    gamma = 1 + 0.1 * torch.randn(Cn)
    wq = packing.tf32_round((torch.randn(384, Cn) / 8) * gamma[None, :]).contiguous()
    wo = packing.tf32_round(torch.randn(Cn, 128) / 11).contiguous()
    ang = torch.arange(Fr, dtype=torch.float32)[:, None] * (10000.0 ** (-torch.arange(0, 32, 2, dtype=torch.float32) / 32))[None, :]
    ang = ang.repeat_interleave(2, dim=1)
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    rel = torch.randn(heads, 63)
    idx = torch.arange(Fr)[None, :] - torch.arange(Fr)[:, None] + 31
    bias = rel[:, idx].contiguous()                                      # [heads, i, j], a function of j - i
    ref = torch.empty_like(x)
    assert emu.temporal_block_fused(x, wq, wo, cos, sin, bias, ref, B, Fr, HW, Cn, heads)
    y = torch.empty_like(x, device="cuda")
    ran = _lib.temporal_block_fused(x.cuda(), wq.cuda(), wo.cuda(), cos.cuda(), sin.cuda(), bias.cuda(), y, B, Fr, HW, Cn, heads)
    assert ran
    err = (y.cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 3e-3, err


@pytest.mark.parametrize("BF,HW,kscale,bias", [(1, 128, 1.0, True), (3, 256, 1.0, False), (2, 4096, 1.0, True),
                                                (150, 256, 1.0, True), (5, 1024, 8.0, True)])
def test_spatial_linear_block_fused_tcgen05(BF, HW, kscale, bias):
    """dpc_spatial_linear_block_fused (LayerNorm + to_qkv + linear attention + to_out + bias + residual) against the fp32
    restatement of conv3d.py:165-174, :232-257 (tests/cpu_emulator.py).  All four contractions are TF32: tolerance 3e-3 of
    the output scale.  kscale = 8 spreads the keys over ~e^40 so the online rescaling of the pixel softmax is exercised;
    BF = 150 gives some persistent CTAs a second frame."""
    torch.manual_seed(6)
    Cn, heads = 64, 4
    x = torch.randn(BF * HW * Cn) * 1.5 + 0.3
    gamma = 1 + 0.1 * torch.randn(Cn)
    w = torch.randn(384, Cn) / 8
    w[128:256] *= kscale
    wq = packing.tf32_round(w * gamma[None, :]).contiguous()
    wo = (torch.randn(Cn, 128) / 11 * 30).contiguous()       # the context averages v over the frame: keep the branch visible
    bo = torch.randn(Cn) if bias else None
    ref = torch.empty_like(x)
    ctx = torch.empty(BF * heads * 32 * 32)
    assert emu.spatial_linear_block_fused(x, wq, wo, bo, ctx, None, ref, BF, HW, Cn, heads)
    y = torch.empty_like(x, device="cuda")
    ctx_d = torch.empty(BF * heads * 32 * 32, device="cuda")
    mt_d = torch.empty(BF * Cn * 128, device="cuda")
    ran = _lib.spatial_linear_block_fused(x.cuda(), wq.cuda(), wo.cuda(), bo.cuda() if bias else None, ctx_d, mt_d, y, BF, HW,
                                          Cn, heads)
    assert ran
    branch = (ref - x).abs().max().item()
    err = (y.cpu() - ref).abs().max().item()
    assert err <= 3e-3 * max(branch, 1.0), (err, branch)


def test_spatial_linear_block_fused_declines_other_shapes():
    z = torch.zeros(16, device="cuda")
    assert _lib.spatial_linear_block_fused(z, z, z, None, z, z, z, 1, 128, 128, 4) is False
    assert _lib.spatial_linear_block_fused(z, z, z, None, z, z, z, 1, 64, 64, 4) is False


def test_temporal_block_fused_declines_other_shapes():
    z = torch.zeros(16, device="cuda")
    assert _lib.temporal_block_fused(z, z, z, z, z, z, z, 1, 40, 4, 64, 4) is False
    assert _lib.temporal_block_fused(z, z, z, z, z, z, z, 1, 32, 4, 128, 4) is False
    assert _lib.temporal_block_fused(z, z, z, z, z, z, z, 1, 32, 6, 64, 4) is False


TC_2D_CASES = [
    # name, N images, H, W, C1, C2, Cout   (3x3 convs of the jellyfish 2-D networks, diffusion_2d_jellyfish.py:189-204)
    ("img64_c64", 4, 64, 64, 64, 0, 64),
    ("img128_c64", 2, 128, 128, 64, 0, 64),
    ("img32_c128", 6, 32, 32, 128, 0, 128),
    ("img16_c256_to_512", 4, 16, 16, 256, 0, 512),       # Cout = 512: two 256-column launches, GroupNorm groups span both
    ("img8_c512", 6, 8, 8, 512, 0, 512),
    ("img16_concat_1024_to_512", 2, 16, 16, 512, 512, 512),
    ("img64_concat_128_to_64", 2, 64, 64, 64, 64, 64),
    ("odd_batch_c128", 3, 32, 32, 128, 0, 64),           # N odd: single-CTA mode
]


@pytest.mark.parametrize("case", TC_2D_CASES, ids=[c[0] for c in TC_2D_CASES])
def test_conv2d_3x3_tcgen05(case):
    """kt = 1 mode of dpc_conv3d_tcgen05 (ntaps = 9, images as single-frame samples) against fp64, with GroupNorm statistics."""
    _, N, H, W, C1, C2, Cout = case
    gen = g(17)
    x1 = torch.randn(N, 1, H, W, C1, generator=gen)
    x2 = torch.randn(N, 1, H, W, C2, generator=gen) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, generator=gen) / (9 * (C1 + C2)) ** 0.5
    bias = torch.randn(Cout, generator=gen)
    xin = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.conv2d(xin[:, 0].permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    wp, _, _ = packing.pack_conv3d(w.to(DEV).unsqueeze(2))
    y, stats, ran_tc = run_conv(x1.to(DEV), wp, 9, x2=x2.to(DEV) if C2 else None, bias=bias.to(DEV), cout=Cout, pad=(0, 1, 1),
                                kernel=(1, 3, 3), gn_groups=8, tcgen05=True)
    assert ran_tc, "the 2-D 3x3 shape must be served by the tcgen05 kernel"
    assert rel_err(y[:, 0], ref) <= TOL_TF32
    v = y.double().reshape(N, -1, 8, Cout // 8)
    assert torch.allclose(stats[:, :, 0], v.sum(dim=(1, 3)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(stats[:, :, 1], (v * v).sum(dim=(1, 3)), rtol=1e-5, atol=1e-4)
