"""Development aid: fused temporal block (dpc_temporal_block_fused) vs the unfused kernel sequence at the metric shape."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Fr, HW, Cn, heads = 32, 4096, 64, 4
torch.manual_seed(0)
x = torch.randn(B * Fr * HW * Cn, device="cuda")
wq = packing.tf32_round(torch.randn(384, Cn, device="cuda") / 8).contiguous()
wo = packing.tf32_round(torch.randn(Cn, 128, device="cuda") / 11).contiguous()
ang = torch.arange(Fr, dtype=torch.float32)[:, None] * (10000.0 ** (-torch.arange(0, 32, 2, dtype=torch.float32) / 32))[None, :]
ang = ang.repeat_interleave(2, dim=1).cuda()
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
bias = torch.zeros(heads, Fr, Fr, device="cuda")
y = torch.empty_like(x)
for _ in range(2):
    assert _lib.temporal_block_fused(x, wq, wo, cos, sin, bias, y, B, Fr, HW, Cn, heads)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.temporal_block_fused(x, wq, wo, cos, sin, bias, y, B, Fr, HW, Cn, heads)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
tok = B * Fr * HW
print(f"fused temporal block B={B}: {ms:.3f} ms  ({2 * tok * Cn * 4 / ms / 1e6:.0f} GB/s algorithmic, {tok * (2*64*384 + 2*128*64 + 4*2*2*32*32) / ms / 1e9:.1f} TFLOP/s)")
