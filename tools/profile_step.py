"""Development aid: warm per-kernel-category timing of one denoising step at the metric shape (CUDA events)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, torch
import diffphycon_b200 as dpc
from diffphycon_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
mj = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=6)
mw = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=2)
d = dpc.GaussianDiffusion([mj, mw], image_size=64, frames=32, eval_2ddpm=True, standard_fixed_ratio=1e5, coeff_ratio=0, w_prob_exp=0.97).cuda()
x = torch.randn(B, 32, 6, 64, 64, device="cuda"); init = torch.rand(B, 64, 64, device="cuda")
fn = dpc.StockSmokeGuidance()
for t in (999, 998):
    x, _ = d.p_sample(x.shape, x, t, design_fn=fn, init=init, _impose_init=True)
with _lib.Profiler() as prof:
    x, _ = d.p_sample(x.shape, x, 997, design_fn=fn, init=init, _impose_init=True)
s = prof.summary()
tot = sum(t for _, t in s.values())
for k, (n, t) in sorted(s.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.2f} ms {100*t/tot:5.1f}% n={n:3d} {k}")
print("total", tot)
