"""Development aid for ncu captures: launches the hot kernels a few times at metric-config shapes (batch 8)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "conv"
B, Fr, S = (int(sys.argv[2]) if len(sys.argv) > 2 else 8), 32, 64
if which == "conv":
    xa = torch.randn(B, Fr, S, S, 64, device=dev)
    w = torch.randn(64, 64, 3, 3, 3, device=dev) / (27 * 64) ** 0.5
    wp, _, _ = packing.pack_conv3d(w)
    bias = torch.zeros(64, device=dev); taps = packing.tap_table(3, 3, 3, S, S, dev)
    y = torch.empty_like(xa); stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
    p = _lib.ConvParams()
    p.x1, p.C1, p.C2 = xa.data_ptr(), 64, 0
    p.w, p.bias, p.y, p.taps, p.ntaps = wp.data_ptr(), bias.data_ptr(), y.data_ptr(), taps.data_ptr(), 27
    p.gn_stats, p.gn_groups = stats.data_ptr(), 8
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = B, Fr, S, S, Fr, S, S
    p.st = p.sh = p.sw = 1; p.pt = p.ph = p.pw = 1; p.oh_mul = p.ow_mul = 1; p.Hfull, p.Wfull = S, S
    p.Cout, p.Npad, p.Kpad = 64, wp.shape[0], wp.shape[1]
    for _ in range(4):
        assert _lib.conv(p, tcgen05=True)
elif which == "tattn":
    qkv = torch.randn(B, Fr, S * S, 384, device=dev)
    cs = torch.randn(Fr, 32, device=dev); bias = torch.randn(4, Fr, Fr, device=dev)
    out = torch.empty(B, Fr, S * S, 128, device=dev)
    for _ in range(4):
        _lib.temporal_attention(qkv, cs, cs, bias, out, B, Fr, S * S, 4, True)
elif which == "linblock":
    BF, HW = B * Fr, S * S
    x = torch.randn(BF * HW * 64, device=dev)
    wq = packing.tf32_round(torch.randn(384, 64, device=dev) / 8).contiguous()
    wo = (torch.randn(64, 128, device=dev) / 11).contiguous(); bo = torch.randn(64, device=dev)
    ctx = torch.empty(BF * 4 * 32 * 32, device=dev); mt = torch.empty(BF * 64 * 128, device=dev); y = torch.empty_like(x)
    for _ in range(3):
        assert _lib.spatial_linear_block_fused(x, wq, wo, bo, ctx, mt, y, BF, HW, 64, 4)
elif which == "stem":
    xin = torch.randn(B, Fr, S, S, 8, device=dev)
    ws = packing.pack_stem_conv(torch.randn(64, 6, 7, 7, 7, device=dev) / 45, 8); bias = torch.zeros(64, device=dev)
    y = torch.empty(B, Fr, S, S, 64, device=dev)
    for _ in range(3):
        assert _lib.stem_conv(xin, ws, bias, y, B, Fr, S, S, 8, 64, 7, 7, 7)
elif which == "tblock":
    HW = S * S
    x = torch.randn(B * Fr * HW * 64, device=dev)
    wq = packing.tf32_round(torch.randn(384, 64, device=dev) / 8).contiguous()
    wo = packing.tf32_round(torch.randn(64, 128, device=dev) / 11).contiguous()
    ang = (torch.arange(Fr, dtype=torch.float32)[:, None] * (10000.0 ** (-torch.arange(0, 32, 2, dtype=torch.float32) / 32))[None, :]).repeat_interleave(2, dim=1).to(dev)
    bias = torch.zeros(4, Fr, Fr, device=dev); y = torch.empty_like(x)
    for _ in range(3):
        assert _lib.temporal_block_fused(x, wq, wo, ang.cos().contiguous(), ang.sin().contiguous(), bias, y, B, Fr, HW, 64, 4)
torch.cuda.synchronize()
