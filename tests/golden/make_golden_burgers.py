"""Golden vectors for the Burgers finite-difference rollout from the UNMODIFIED reference
`dataset/apps/generate_burgers.py::burgers_numeric_solve_free` (build container only)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402

gb = ref_import.generate_burgers_module()
rng = np.random.default_rng(0)
N, s, Nt = 4, 128, 10
x = np.linspace(0, 1, s + 2)[1:-1]
u0 = np.stack([a * np.exp(-((x - m) ** 2) / (2 * 0.1 ** 2)) for a, m in zip(rng.uniform(-1, 1, N), rng.uniform(0.2, 0.8, N))])
f = rng.standard_normal((N, Nt, 1)) * np.exp(-((x[None, None] - rng.uniform(0.2, 0.8, (N, Nt, 1))) ** 2) / (2 * 0.15 ** 2))
u0, f = u0.astype(np.float32), f.astype(np.float32)
torch.set_num_threads(4)
traj = gb.burgers_numeric_solve_free(torch.from_numpy(u0), torch.from_numpy(f), visc=0.01, T=1.0, dt=1e-4, num_t=10)
np.savez_compressed(os.path.join(HERE, "burgers_rollout.npz"), u0=u0, f=f, traj=traj.numpy())
print(traj.shape, float(traj.abs().max()))
