"""Development aid: time the persistent rollout kernel (B trajectories, 256 frames) with CUDA events."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import smoke_rollout as sr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sim = sr.init_sim_128()
g = torch.Generator().manual_seed(0)
c1 = torch.randn(B, 32, 64, 64, generator=g).cuda() * 0.5
c2 = torch.randn(B, 32, 64, 64, generator=g).cuda() * 0.5
dens = torch.rand(B, 64, 64, generator=g).cuda()
out = sr.solver_batch(sim, sr.init_velocity_(), dens[:2], c1[:2, :4].contiguous(), c2[:2, :4].contiguous(), 8)   # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = sr.solver_batch(sim, sr.init_velocity_(), dens, c1, c2, T)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
its = out["iterations"][:, 1:].float()
print(f"rollout B={B} T={T}: {ms:.1f} ms total, {ms / B:.2f} ms per trajectory (all in parallel), mean CG iterations {its.mean().item():.1f}, "
      f"{its.sum().item() / (ms * 1e-3) / 1e6:.2f} M CG iterations/s")
