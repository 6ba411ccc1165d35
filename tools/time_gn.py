"""Development aid: GroupNorm+SiLU kernel at the three levels of the metric shape (CUDA events)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for S, C in ((64, 64), (32, 128), (16, 256)):
    rps = 32 * S * S
    for res in (False, True):
        y = torch.randn(B, rps, C, device="cuda"); out = torch.empty_like(y)
        r = torch.randn(B, rps, C, device="cuda") if res else None
        stats = torch.zeros(B, 8, 2, dtype=torch.float64, device="cuda"); stats[:, :, 1] = rps * C / 8
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        ss = torch.randn(B, 4 * C, device="cuda")
        for _ in range(2): _lib.groupnorm_silu(y, stats, gamma, beta, ss, 4 * C, C, r, out, B, rps, C, 8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): _lib.groupnorm_silu(y, stats, gamma, beta, ss, 4 * C, C, r, out, B, rps, C, 8)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = y.numel() * 4 * (3 if res else 2) / 1e9
        print(f"groupnorm_silu B={B} S={S} C={C} residual={res}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
