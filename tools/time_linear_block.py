"""Development aid: fused spatial-linear-attention block (dpc_spatial_linear_block_fused) at the metric shape."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Fr, HW, Cn, heads = 32, 4096, 64, 4
BF = B * Fr
torch.manual_seed(0)
x = torch.randn(BF * HW * Cn, device="cuda")
wq = packing.tf32_round(torch.randn(384, Cn, device="cuda") / 8).contiguous()
wo = (torch.randn(Cn, 128, device="cuda") / 11).contiguous()
bo = torch.randn(Cn, device="cuda")
ctx = torch.empty(BF * heads * 32 * 32, device="cuda")
mt = torch.empty(BF * Cn * 128, device="cuda")
y = torch.empty_like(x)
for _ in range(2):
    assert _lib.spatial_linear_block_fused(x, wq, wo, bo, ctx, mt, y, BF, HW, Cn, heads)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.spatial_linear_block_fused(x, wq, wo, bo, ctx, mt, y, BF, HW, Cn, heads)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
tok = BF * HW
print(f"fused spatial linear block B={B}: {ms:.3f} ms  ({3 * tok * Cn * 4 / ms / 1e6:.0f} GB/s algorithmic (x twice + y))")
