"""CPU restatement (NumPy float32) of the Burgers finite-difference rollout — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/dataset/apps/generate_burgers.py:207-299 (`burgers_numeric_solve_free`) and :95-110
(`Diff_mat_1D`): explicit Euler, u <- u + dt*(-1/2 * D(u^2)/(2dx) + visc * D2(u)/dx^2 + f_k), homogeneous Dirichlet ends
re-imposed before every step, dx = 1/(s+1), force piecewise constant over floor(steps/Nt) steps, one record per force
window.  Pinned by tests/golden/burgers_rollout.npz (unmodified reference, tests/golden/make_golden_burgers.py)."""
import math

import numpy as np


def coefficients(s: int, visc: float):
    """The float32 stencil coefficients the reference builds from its sparse difference matrices (interior rows)."""
    dx = 1.0 / (s + 1)
    t = (np.array([-1.0, 1.0]) / (2 * dx)).astype(np.float32)
    d = (visc * np.array([1.0, -2.0, 1.0]) / dx ** 2).astype(np.float32)
    return t, d


def burgers_numeric_solve_free(u0: np.ndarray, f: np.ndarray, visc: float, T: float, dt: float = 1e-4, num_t: int = 10):
    u0 = np.asarray(u0, dtype=np.float32)
    f = np.asarray(f, dtype=np.float32)
    N, s = u0.shape
    Nt = f.shape[1]
    assert Nt == num_t
    steps = math.ceil(T / dt)
    rec = math.floor(steps / Nt)
    t, d = coefficients(s, visc)
    u = np.pad(u0, ((0, 0), (1, 1)))
    fp = np.pad(f, ((0, 0), (0, 0), (1, 1)))
    sol = np.zeros((N, Nt, s), dtype=np.float32)
    c, fi = 0, -1
    dt32 = np.float32(dt)
    for j in range(steps):
        u[:, 0] = 0
        u[:, -1] = 0
        us = u * u
        transport = us[:, :-2] * t[0] + us[:, 2:] * t[1]
        diffusion = u[:, :-2] * d[0] + u[:, 1:-1] * d[1] + u[:, 2:] * d[2]
        if j % rec == 0:
            fi += 1
        new = u[:, 1:-1] + dt32 * (np.float32(-0.5) * transport + diffusion + fp[:, fi, 1:-1])
        u = np.pad(new.astype(np.float32), ((0, 0), (1, 1)))
        if (j + 1) % rec == 0:
            sol[:, c] = u[:, 1:-1]
            c += 1
    return np.concatenate([u0[:, None, :], sol], axis=1)
