"""Development aid: time the jellyfish guidance (ForceUnet + boundary-updater Unet, forward + backward) and update_bd forward at
the BASELINE.json config-3 shape (B=8, 20 frames, 128x128 -> 160 images), with a per-kernel-category breakdown."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffphycon_b200 as dpc
from diffphycon_b200 import _lib
B, Fr, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 20, int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
bd = dpc.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3).cuda()
fm = dpc.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4).cuda()
x = torch.rand(B, Fr, 4, S, S, device="cuda") * 2 - 1
bd_0 = torch.rand(B, Fr, 3, S, S, device="cuda")
fn = dpc.JellyfishGuidance(fm, bd, -1.0, 2.0, 1000.0)
for _ in range(2):
    g = fn(x, bd_0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    g = fn(x, bd_0)
e1.record(); torch.cuda.synchronize()
print(f"guidance fwd+bwd B={B} S={S}: {e0.elapsed_time(e1)/3:.1f} ms")
e0.record()
for _ in range(3):
    y = bd(bd_0.reshape(B * Fr, 3, S, S), x[:, :, 3].mean((-1, -2)).reshape(-1))
e1.record(); torch.cuda.synchronize()
print(f"update_bd fwd: {e0.elapsed_time(e1)/3:.1f} ms")
with _lib.Profiler() as prof:
    g = fn(x, bd_0)
s = prof.summary()
tot = sum(t for _, t in s.values())
for k, (n, t) in sorted(s.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:9.2f} ms {100*t/tot:5.1f}% n={n:3d} {k}")
print("total", tot)
