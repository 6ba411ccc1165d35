"""B200-native Burgers finite-difference rollout — drop-in for `dataset/apps/generate_burgers.py::burgers_numeric_solve_free`
(generate_burgers.py:207-299), which `inference/inference_1d_burgers.py:294` calls to evaluate the sampled control.
Same signature and return value ([N, Nt+1, s] float32); the 10 000 explicit Euler steps run inside one kernel launch."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


@torch.no_grad()
@_lib.device_guarded
def burgers_numeric_solve_free(u0, f, visc, T, dt=1e-4, num_t=10, mode=None):
    if mode == 'const':
        raise ValueError
    assert f.size()[1] == num_t, 'check number of time interval'
    if not u0.is_cuda:
        raise RuntimeError("diffphycon_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    s = u0.size(-1)
    Nt = f.size(1)
    N = f.size(0)
    assert u0.size(0) == N
    dx = 1.0 / (s + 1)
    steps = math.ceil(T / dt)
    # float32 stencil coefficients exactly as the reference derives them (generate_burgers.py:255-258)
    t = (np.array([-1.0, 1.0]) / (2 * dx)).astype(np.float32)
    d = (visc * np.array([1.0, -2.0, 1.0]) / dx ** 2).astype(np.float32)
    u0c = u0.reshape(N, s).float().contiguous()
    fc = f.reshape(N, Nt, s).float().contiguous()
    traj = torch.empty(N, Nt + 1, s, dtype=torch.float32, device=u0.device)
    _lib.burgers_rollout(u0c, fc, traj, N, s, Nt, steps, float(t[0]), float(t[1]), float(d[0]), float(d[1]), float(d[2]),
                         float(np.float32(dt)))
    return traj
