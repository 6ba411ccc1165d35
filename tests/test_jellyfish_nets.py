"""Jellyfish surrogate networks: the CPU oracle against goldens of the UNMODIFIED reference
(tests/golden/make_golden_jellyfish_nets.py), and — on the GPU — the engine's `Unet` / `ForceUnet` forward and the whole
guidance gradient `JellyfishGuidance` (forward + hand-written backward through both networks) against the same goldens,
plus every backward kernel against torch.autograd on the same op.

Tolerances: fp32 class ("3xtf32" contractions) 2e-4 of the tensor's max magnitude; TF32 contraction class 1e-2 (two chained
networks, forward and backward: ~110 TF32 convolutions between the input and the gradient)."""
import os

import numpy as np
import pytest
import torch

from oracle import jellyfish_nets_oracle as jo

KW_U = dict(dim=64, dim_mults=(1, 2, 4, 8), channels=3, out_dim=3)
KW_F = dict(dim=64, dim_mults=(1, 2, 4, 8), channels=4, out_dim=1)
TOL = {"3xtf32": 2e-4, "tf32": 1e-2}


def rel(a, b):
    return (a - b).abs().max().item() / max(1e-30, b.abs().max().item())


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "jellyfish_nets.npz"))


def test_oracle_matches_reference_golden(gold):
    z = gold
    pu, pf = jo.make_params("unet", int(z["seed_unet"]), **KW_U), jo.make_params("force", int(z["seed_force"]), **KW_F)
    tag = "s32"
    x, bd_0 = torch.from_numpy(z[f"{tag}/x"]), torch.from_numpy(z[f"{tag}/bd_0"])
    B, Fr, _, S, _ = x.shape
    with torch.no_grad():
        theta = x[:, :, 3].mean((-1, -2))
        pred_bd = jo.unet_forward(pu, bd_0.reshape(B * Fr, 3, S, S), theta.reshape(B * Fr))
        assert rel(pred_bd, torch.from_numpy(z[f"{tag}/pred_bd"])) <= 2e-5
        pressure = (0.5 * x[:, :, 2] + 0.5) * (float(z["p_max"]) - float(z["p_min"])) + float(z["p_min"])
        force = jo.force_forward(pf, torch.cat((pressure.reshape(B * Fr, 1, S, S), pred_bd), 1))
        assert rel(force, torch.from_numpy(z[f"{tag}/force"])) <= 2e-5
    for key, reg in (("grad", float(z["reg_ratio"])), ("grad_noreg", 0.0)):
        g = jo.design_fn(x, bd_0, force_params=pf, bd_params=pu, p_min=float(z["p_min"]), p_max=float(z["p_max"]), reg_ratio=reg)
        ref = torch.from_numpy(z[f"{tag}/{key}"])
        assert rel(g[:, :, :3], ref[:, :, :3]) <= 1e-4, key
        assert rel(g[:, :, 3], ref[:, :, 3]) <= 1e-4, key


def test_engine_host_logic_on_cpu_emulator(gold, monkeypatch):
    """The engine's HOST logic — weight standardisation and packing, dgrad weight transposition / flipping / concat splits,
    the backward tape, gradient accumulation over skip connections, the guidance glue — with every kernel replaced by the torch
    emulation of the C-ABI (tests/cpu_emulator.py), against the unmodified reference's gradient."""
    from tests import cpu_emulator
    import diffphycon_b200 as dpc
    from diffphycon_b200 import jellyfish_nets
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(jellyfish_nets, "_require_cuda", lambda x: None)
    z = gold
    bd = dpc.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
    fm = dpc.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
    bd.load_state_dict(jo.make_params("unet", int(z["seed_unet"]), **KW_U), strict=True)
    fm.load_state_dict(jo.make_params("force", int(z["seed_force"]), **KW_F), strict=True)
    bd.precision = fm.precision = "3xtf32"      # no TF32 rounding of the packed weights
    tag = "s32"
    x, bd_0 = torch.from_numpy(z[f"{tag}/x"]), torch.from_numpy(z[f"{tag}/bd_0"])
    B, Fr, _, S, _ = x.shape
    theta = x[:, :, 3].mean((-1, -2))
    pred_bd = bd(bd_0.reshape(B * Fr, 3, S, S).contiguous(), theta.reshape(B * Fr))
    assert rel(pred_bd, torch.from_numpy(z[f"{tag}/pred_bd"])) <= 1e-4
    for key, reg in (("grad", float(z["reg_ratio"])), ("grad_noreg", 0.0)):
        g = dpc.JellyfishGuidance(fm, bd, float(z["p_min"]), float(z["p_max"]), reg)(x, bd_0)
        ref = torch.from_numpy(z[f"{tag}/{key}"])
        assert rel(g[:, :, 2], ref[:, :, 2]) <= 2e-4, key
        assert rel(g[:, :, 3], ref[:, :, 3]) <= 2e-4, key
        assert g[:, :, :2].abs().max().item() == 0.0


# ------------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------------
def engine_nets(z, precision):
    import diffphycon_b200 as dpc
    bd = dpc.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
    fm = dpc.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
    bd.load_state_dict(jo.make_params("unet", int(z["seed_unet"]), **KW_U), strict=True)
    fm.load_state_dict(jo.make_params("force", int(z["seed_force"]), **KW_F), strict=True)
    bd.precision = fm.precision = precision
    return bd.cuda(), fm.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
@pytest.mark.parametrize("tag", ["s32", "s64"])
def test_engine_forward_matches_reference(gold, tag, precision):
    z = gold
    bd, fm = engine_nets(z, precision)
    x, bd_0 = torch.from_numpy(z[f"{tag}/x"]).cuda(), torch.from_numpy(z[f"{tag}/bd_0"]).cuda()
    B, Fr, _, S, _ = x.shape
    theta = x[:, :, 3].mean((-1, -2))
    pred_bd = bd(bd_0.reshape(B * Fr, 3, S, S).contiguous(), theta.reshape(B * Fr))
    ref_bd = torch.from_numpy(z[f"{tag}/pred_bd"])
    assert rel(pred_bd.cpu(), ref_bd) <= TOL[precision]
    pressure = (0.5 * x[:, :, 2] + 0.5) * (float(z["p_max"]) - float(z["p_min"])) + float(z["p_min"])
    inp = torch.cat((pressure.reshape(B * Fr, 1, S, S), ref_bd.cuda()), 1).contiguous()
    force = fm(inp)
    assert rel(force.cpu(), torch.from_numpy(z[f"{tag}/force"])) <= TOL[precision]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
@pytest.mark.parametrize("tag", ["s32", "s64"])
def test_engine_guidance_gradient_matches_reference(gold, tag, precision):
    """design_fn(x, bd_0) = (dJ/dstate, dJ/dtheta_expand) of the unmodified reference's force_fn (autograd through both nets)."""
    import diffphycon_b200 as dpc
    z = gold
    bd, fm = engine_nets(z, precision)
    x, bd_0 = torch.from_numpy(z[f"{tag}/x"]).cuda(), torch.from_numpy(z[f"{tag}/bd_0"]).cuda()
    for key, reg in (("grad", float(z["reg_ratio"])), ("grad_noreg", 0.0)):
        fn = dpc.JellyfishGuidance(fm, bd, float(z["p_min"]), float(z["p_max"]), reg)
        g = fn(x, bd_0).cpu()
        ref = torch.from_numpy(z[f"{tag}/{key}"])
        assert g.shape == ref.shape
        assert rel(g[:, :, 2], ref[:, :, 2]) <= TOL[precision], (key, "pressure gradient")
        assert g[:, :, :2].abs().max().item() == 0.0 and ref[:, :, :2].abs().max().item() == 0.0
        assert rel(g[:, :, 3], ref[:, :, 3]) <= TOL[precision], (key, "theta gradient")


@pytest.mark.gpu
def test_engine_guidance_at_config3_shape_vs_oracle():
    """BASELINE.json config 3 shape (128x128, dim 64) at batch 1 x 2 frames: engine (tf32) against the CPU oracle's autograd."""
    import diffphycon_b200 as dpc
    pu, pf = jo.make_params("unet", 71, **KW_U), jo.make_params("force", 72, **KW_F)
    bd = dpc.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
    fm = dpc.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
    bd.load_state_dict(pu)
    fm.load_state_dict(pf)
    bd, fm = bd.cuda(), fm.cuda()
    g = torch.Generator().manual_seed(73)
    S, B, Fr = 128, 1, 2
    x = torch.rand(B, Fr, 4, S, S, generator=g) * 2 - 1
    x[:, :, 3] = 0.5 + 0.3 * x[:, :, 3]
    bd_0 = torch.cat([(torch.rand(B, Fr, 1, S, S, generator=g) > 0.7).float(), torch.rand(B, Fr, 2, S, S, generator=g) - 0.5], 2)
    ref = jo.design_fn(x, bd_0, force_params=pf, bd_params=pu, p_min=-1.0, p_max=2.0, reg_ratio=0.0)
    got = dpc.JellyfishGuidance(fm, bd, -1.0, 2.0, 0.0)(x.cuda(), bd_0.cuda()).cpu()
    assert rel(got[:, :, 2], ref[:, :, 2]) <= TOL["tf32"]
    assert rel(got[:, :, 3], ref[:, :, 3]) <= TOL["tf32"]


# ---- backward kernels against torch.autograd on the same op (fp32 on the device) -----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("C,hw,with_ss", [(64, 1024, True), (128, 100, False), (512, 16, True)])
def test_gn_silu_bwd_kernel(C, hw, with_ss):
    from diffphycon_b200 import _lib
    torch.manual_seed(0)
    N, G = 3, 8
    y = torch.randn(N, hw, C, device="cuda") * 2 + 0.3
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ss = torch.randn(N, 2 * C + 10, device="cuda") * 0.5 if with_ss else None
    dout = torch.randn(N, hw, C, device="cuda")
    yr = y.clone().requires_grad_()
    ssr = ss.clone().requires_grad_() if with_ss else None
    z = torch.nn.functional.group_norm(yr.permute(0, 2, 1), G, gamma, beta, eps=1e-5).permute(0, 2, 1)
    if with_ss:
        z = z * (ssr[:, None, 5:5 + C] + 1) + ssr[:, None, 5 + C:5 + 2 * C]
    out = torch.nn.functional.silu(z)
    grads = torch.autograd.grad(out, [yr] + ([ssr] if with_ss else []), dout)
    stats = torch.stack([y.double().reshape(N, hw, G, C // G).sum((1, 3)), y.double().square().reshape(N, hw, G, C // G).sum((1, 3))],
                        -1).contiguous()
    dy = torch.empty_like(y)
    sums = torch.empty(N * C * 2, dtype=torch.float64, device="cuda")
    dss = torch.zeros_like(ss) if with_ss else None
    _lib.gn_silu_bwd(y, stats, gamma, beta, ss, (2 * C + 10) if with_ss else 0, 5 if with_ss else 0, dout, dy, sums, dss, N, hw, C, G)
    assert rel(dy, grads[0]) <= 2e-5
    if with_ss:
        assert rel(dss, grads[1]) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("C", [64, 256, 512])
def test_layernorm_bwd_kernel(C):
    from diffphycon_b200 import _lib
    torch.manual_seed(1)
    rows = 777
    x = torch.randn(rows, C, device="cuda") * 1.5 + 0.2
    g = torch.randn(C, device="cuda")
    dy, add = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
    xr = x.clone().requires_grad_()
    var, mean = xr.var(1, unbiased=False, keepdim=True), xr.mean(1, keepdim=True)
    out = (xr - mean) * (var + 1e-5).rsqrt() * g
    (ref,) = torch.autograd.grad(out, xr, dy)
    dx = torch.empty_like(x)
    _lib.layernorm_channels_bwd(x, g, dy, add, dx, rows, C)
    assert rel(dx, ref + add) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [16, 200, 1024])
def test_linear_attention_fwd_bwd_kernels(hw):
    from diffphycon_b200 import _lib
    torch.manual_seed(2)
    N, heads = 3, 4
    hid = heads * 32
    qkv = torch.randn(N, hw, 3 * hid, device="cuda")
    dout = torch.randn(N, hw, hid, device="cuda")
    vs = 1.0 / hw
    qr = qkv.clone().requires_grad_()
    q, k, v = [t.reshape(N, hw, heads, 32).permute(0, 2, 3, 1) for t in qr.chunk(3, dim=2)]     # [N, heads, d, n]
    qs, ks, vv = q.softmax(dim=-2) * 32 ** -0.5, k.softmax(dim=-1), v * vs
    ctx_ref = torch.einsum('bhdn,bhen->bhde', ks, vv)
    out_ref = torch.einsum('bhde,bhdn->bhen', ctx_ref, qs).permute(0, 3, 1, 2).reshape(N, hw, hid)
    (dref,) = torch.autograd.grad(out_ref, qr, dout)
    ctx, kstat = torch.empty(N * heads * 32 * 32, device="cuda"), torch.empty(N * heads * 32 * 2, device="cuda")
    out = torch.empty(N, hw, hid, device="cuda")
    _lib.spatial_linear_attention_ex(qkv, ctx, kstat, out, N, hw, heads, vs)
    assert rel(out, out_ref.detach()) <= 2e-5
    dctx, dqkv = torch.empty_like(ctx), torch.empty_like(qkv)
    _lib.linattn2d_bwd(qkv, ctx, kstat, dout, dctx, dqkv, N, hw, heads, vs)
    assert rel(dqkv, dref) <= 5e-5


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [4, 64, 256])
def test_attention_bwd_kernel(hw):
    from diffphycon_b200 import _lib
    torch.manual_seed(3)
    N, heads = 2, 4
    hid = heads * 32
    qkv = torch.randn(N, hw, 3 * hid, device="cuda")
    dout = torch.randn(N, hw, hid, device="cuda")
    qr = qkv.clone().requires_grad_()
    q, k, v = [t.reshape(N, hw, heads, 32).permute(0, 2, 1, 3) for t in qr.chunk(3, dim=2)]     # [N, heads, n, d]
    attn = (q * 32 ** -0.5 @ k.transpose(-1, -2)).softmax(-1)
    out_ref = (attn @ v).permute(0, 2, 1, 3).reshape(N, hw, hid)
    (dref,) = torch.autograd.grad(out_ref, qr, dout)
    out = torch.empty(N, hw, hid, device="cuda")
    _lib.spatial_attention(qkv, out, N, hw, heads)
    assert rel(out, out_ref.detach()) <= 2e-5
    dqkv = torch.empty_like(qkv)
    _lib.attention2d_bwd(qkv, out, dout, dqkv, N, hw, heads)
    assert rel(dqkv, dref) <= 5e-5


@pytest.mark.gpu
def test_small_helper_kernels():
    from diffphycon_b200 import _lib
    torch.manual_seed(4)
    N, H, W, C = 2, 5, 6, 64
    dy = torch.randn(N, 2 * H, 2 * W, C, device="cuda")
    dx = torch.empty(N, H, W, C, device="cuda")
    _lib.sumpool2x2(dy, dx, N, H, W, C)
    ref = dy.reshape(N, H, 2, W, 2, C).sum((2, 4))
    assert rel(dx, ref) <= 1e-6
    a, b = torch.randn(1000, device="cuda"), torch.randn(1000, device="cuda")
    o = torch.empty_like(a)
    _lib.add(a, b, o, 1000)
    assert torch.equal(o, a + b)
    x = torch.randn(N, H * W, 512, device="cuda")
    Wt, bias = torch.randn(3, 512, device="cuda"), torch.randn(3, device="cuda")
    out = torch.empty(N, 3, device="cuda")
    _lib.mean_head(x, Wt, bias, out, N, H * W, 512, 3)
    assert rel(out, x.mean(1) @ Wt.t() + bias) <= 2e-5
    dout = torch.randn(N, 3, device="cuda")
    dxh = torch.empty_like(x)
    _lib.mean_head_bwd(dout, Wt, dxh, N, H * W, 512, 3)
    assert rel(dxh, ((dout @ Wt) / (H * W))[:, None, :].expand(N, H * W, 512)) <= 2e-5


@pytest.mark.gpu
def test_time_mlp_fwd_bwd_kernels():
    from diffphycon_b200 import _lib
    import math
    torch.manual_seed(5)
    N, dim, total = 5, 64, 1000
    tdim = 4 * dim
    t = torch.rand(N, device="cuda") * 0.8 + 0.1
    half = dim // 2
    freqs = torch.exp(torch.arange(half, device="cuda") * -(math.log(10000) / (half - 1))).float()
    w1, b1 = torch.randn(tdim, dim, device="cuda") / 8, torch.randn(tdim, device="cuda") * 0.1
    w2, b2 = torch.randn(tdim, tdim, device="cuda") / 16, torch.randn(tdim, device="cuda") * 0.1
    wp, bp = torch.randn(total, tdim, device="cuda") / 16, torch.randn(total, device="cuda") * 0.1
    dss = torch.randn(N, total, device="cuda")
    tr = t.clone().requires_grad_()
    emb = tr[:, None] * freqs[None]
    emb = torch.cat((emb.sin(), emb.cos()), -1)
    te = torch.nn.functional.gelu(emb @ w1.t() + b1) @ w2.t() + b2
    ss = torch.nn.functional.silu(te) @ wp.t() + bp
    (dref,) = torch.autograd.grad(ss, tr, dss)
    hidden, t_emb = torch.empty(N * tdim, device="cuda"), torch.empty(N * tdim, device="cuda")
    _lib.time_embed_f32(t, freqs, w1, b1, w2, b2, hidden, t_emb, N, dim)
    assert rel(t_emb.reshape(N, tdim), te.detach()) <= 2e-5
    ssk = torch.empty(N * total, device="cuda")
    _lib.time_proj(t_emb, wp, bp, ssk, N, tdim, total)
    assert rel(ssk.reshape(N, total), ss.detach()) <= 2e-5
    dt = torch.empty(N, device="cuda")
    _lib.time_mlp_bwd(t, freqs, w1, b1, w2, wp, t_emb, dss, dt, N, dim, total)
    assert rel(dt, dref) <= 1e-4
