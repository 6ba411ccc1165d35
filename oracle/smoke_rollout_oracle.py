"""CPU restatement (NumPy, fp64) of the smoke ground-truth rollout — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/dataset/apps/evaluate_solver.py (`es.py:line`) and the vendored PhiFlow 1.0.x under
/root/reference/phi (`flow.py`, `nd.py` = phi/math/nd.py, `scipy_backend.py`, `sparse.py` = phi/solver/sparse.py,
`base.py` = phi/solver/base.py).  Pinned by tests/golden/smoke_rollout.npz, produced by the UNMODIFIED reference
(tests/golden/make_golden_rollout.py) — see tests/test_rollout_oracle.py.

Conventions: fields are indexed [y, x]; the staggered (MAC) velocity is [128, 128, 2] with component 0 = x, 1 = y
(nd.py:312-342); the domain has 127 x 127 cells, open on all four sides (es.py:94-97).
"""
from __future__ import annotations

import numpy as np

N = 127          # cells per side
NS = N + 1       # staggered samples per side
ACCURACY = 1e-8  # es.py:144
MAX_ITERATIONS = 500  # sparse.py:89

# (size_y, size_x), (origin_y, origin_x) of every obstacle of build_obstacles_pi_128 (es.py:32-92)
OBSTACLES_128 = [
    ((1, 96), (16, 16)),
    ((8, 1), (16, 16)), ((16, 1), (40, 16)), ((40, 1), (72, 16)),
    ((8, 1), (16, 112)), ((16, 1), (40, 112)), ((40, 1), (72, 112)),
    ((1, 8), (112, 16)), ((1, 16), (112, 40)), ((1, 16), (112, 72)), ((1, 8), (112, 104)),
    ((16, 1), (64, 48)), ((16, 1), (96, 48)), ((16, 1), (64, 80)), ((16, 1), (96, 80)),
    ((1, 128 - 40 - 40), (40, 40)),
]


def fluid_mask_128() -> np.ndarray:
    """flow.py:172-193 (set_obstacle): 1 = fluid, 0 = obstacle; the active mask is identical (both are cleared)."""
    m = np.ones((N, N), dtype=np.int8)
    for (sy, sx), (oy, ox) in OBSTACLES_128:
        m[oy:oy + sy, ox:ox + sx] = 0
    return m


def velocity_mask(fluid: np.ndarray) -> np.ndarray:
    """flow.py:455-473 (_create_staggered_velocity_mask) with every boundary open: the fluid mask is padded with ones,
    a face is live iff both adjacent cells are fluid.  Returns [128,128,2] (component 0 = x faces, 1 = y faces)."""
    ext = np.pad(fluid.astype(np.float32), 1, constant_values=1)  # [129,129]
    my = np.minimum(ext[1:, 1:], ext[:-1, 1:])   # d = 0 (y)
    mx = np.minimum(ext[1:, 1:], ext[1:, :-1])   # d = 1 (x)
    return np.stack([mx, my], axis=-1)


def divergence(v: np.ndarray) -> np.ndarray:
    """nd.py:367-377: [128,128,2] -> [127,127]."""
    return (v[1:, :-1, 1] - v[:-1, :-1, 1]) + (v[:-1, 1:, 0] - v[:-1, :-1, 0])


def laplace_coefficients(fluid: np.ndarray):
    """sparse.py:27-78 for an all-open domain: off-diagonal (c, nbr) = active[c]*active[nbr] for in-grid neighbours,
    diagonal = min(-(number of fluid neighbours, cells outside the grid count as fluid), -1).
    Returns (lower_y, upper_y, lower_x, upper_x, diag), each [127,127] float64."""
    a = fluid.astype(np.float64)
    act_ext = np.pad(a, 1, constant_values=0)     # pad_active: zeros
    flu_ext = np.pad(a, 1, constant_values=1)     # pad_fluid (open): ones
    c = act_ext[1:-1, 1:-1]
    up_y, lo_y = act_ext[2:, 1:-1] * c, act_ext[:-2, 1:-1] * c
    up_x, lo_x = act_ext[1:-1, 2:] * c, act_ext[1:-1, :-2] * c
    center = -(flu_ext[2:, 1:-1] + flu_ext[:-2, 1:-1]) - (flu_ext[1:-1, 2:] + flu_ext[1:-1, :-2])
    return lo_y, up_y, lo_x, up_x, np.minimum(center, -1.0)


def apply_laplace(coef, p: np.ndarray) -> np.ndarray:
    lo_y, up_y, lo_x, up_x, diag = coef
    out = diag * p
    out[1:, :] += lo_y[1:, :] * p[:-1, :]
    out[:-1, :] += up_y[:-1, :] * p[1:, :]
    out[:, 1:] += lo_x[:, 1:] * p[:, :-1]
    out[:, :-1] += up_x[:, :-1] * p[:, 1:]
    return out


def conjugate_gradient(coef, k: np.ndarray, accuracy=ACCURACY, max_iterations=MAX_ITERATIONS):
    """base.py:56-103 with x0 = 0, including its aliasing quirk: `residual` and `momentum` start as the SAME array and
    `residual -= ...` is in place, so the first momentum update sees the already updated residual:
    p1 = r1 + b*r1 (not r1 + b*r0).  Stop test: max|r| >= accuracy is checked before every iteration."""
    x = np.zeros_like(k)
    r = k.copy()
    p = r            # aliased on purpose (first iteration only)
    Ap = apply_laplace(coef, p)
    it = 0
    while np.max(np.abs(r)) >= accuracy:
        if it == max_iterations:
            break
        tmp = np.sum(p * Ap)
        a = np.sum(p * r) / tmp
        x += a * p
        r -= a * Ap                      # in place: also changes p while it aliases r
        b = -np.sum(r * Ap) / tmp
        p = r + b * p                    # new array from now on
        Ap = apply_laplace(coef, p)
        it += 1
    return x, it


def pressure_gradient(p: np.ndarray) -> np.ndarray:
    """nd.py:602-614 (StaggeredGrid.gradient, symmetric padding): [127,127] -> [128,128,2]."""
    f = np.pad(p, 1, mode="symmetric")            # [129,129]
    gy = f[1:, 1:] - f[:-1, 1:]
    gx = f[1:, 1:] - f[1:, :-1]
    return np.stack([gx, gy], axis=-1)


def divergence_free(v: np.ndarray, vmask: np.ndarray, coef):
    """flow.py:318-327 followed by with_boundary_conditions (es.py:144-145)."""
    v = v * vmask
    p, it = conjugate_gradient(coef, divergence(v))
    v = v - pressure_gradient(p) * vmask
    return v * vmask, p, it


def inject_control(prev_v: np.ndarray, c1f: np.ndarray, c2f: np.ndarray) -> np.ndarray:
    """es.py:128-142: the outer ring of width 16 of the staggered field comes from the control, the interior from the
    previous velocity."""
    ctrl = np.zeros((NS, NS, 2), dtype=np.float64)
    ctrl[:, :, 0] = c1f
    ctrl[:, :, 1] = c2f
    ctrl[16:112, 16:112, :] = 0
    cur = ctrl.copy()
    cur[16:112, 16:112, :] = prev_v[16:112, 16:112, :]
    return cur


def advect(field: np.ndarray, v: np.ndarray, dt: float = 1.0) -> np.ndarray:
    """nd.py:422-427 + scipy_backend.py:58-77, :181-185: semi-Lagrangian back-trace with the cell-centred velocity
    (face sum / 2), coordinates clamped to [0, 127] (not 126), bilinear interpolation on the 127 cell centres and ZERO
    beyond the last centre (interpn fill_value=0).  field: [127,127] float32 -> float32."""
    vy = (v[1:, :-1, 1] + v[:-1, :-1, 1]) / 2
    vx = (v[:-1, 1:, 0] + v[:-1, :-1, 0]) / 2
    yy, xx = np.meshgrid(np.arange(N, dtype=np.float32), np.arange(N, dtype=np.float32), indexing="ij")
    sy = np.clip(yy - vy * dt, 0, N)
    sx = np.clip(xx - vx * dt, 0, N)
    inside = (sy <= N - 1) & (sx <= N - 1)
    y0 = np.clip(np.floor(sy).astype(np.int64), 0, N - 2)
    x0 = np.clip(np.floor(sx).astype(np.int64), 0, N - 2)
    fy, fx = sy - y0, sx - x0
    f = field.astype(np.float64)
    val = (f[y0, x0] * (1 - fy) * (1 - fx) + f[y0 + 1, x0] * fy * (1 - fx) +
           f[y0, x0 + 1] * (1 - fy) * fx + f[y0 + 1, x0 + 1] * fy * fx)
    return np.where(inside, val, 0.0).astype(field.dtype)


def bucket_masks():
    """es.py:150-171."""
    bucket_pos = [(112, 24 - 2, 127 - 112, 16 + 4), (112, 56 - 2, 127 - 112, 16 + 4), (112, 88 - 2, 127 - 112, 16 + 4)]
    bucket_pos_y = [(24 - 2, 0, 16 + 4, 16), (56 - 2, 0, 16 + 4, 16), (24 - 2, 112, 16 + 4, 127 - 112),
                    (56 - 2, 112, 16 + 4, 127 - 112)]
    masks = []
    concat = np.zeros((128, 128))
    keep = np.ones((128, 128))
    for y, x, ly, lx in bucket_pos + bucket_pos_y:
        m = np.zeros((128, 128))
        m[y:y + ly, x:x + lx] = 1
        concat[y:y + ly, x:x + lx] = 1
        keep[y:y + ly, x:x + lx] = 0
        masks.append(m)
    return masks, concat, keep


def solver(fluid: np.ndarray, init_velocity: np.ndarray, init_density: np.ndarray, c1: np.ndarray, c2: np.ndarray,
           per_timelength: int, dt: float = 1.0):
    """es.py:205-310.  init_velocity [128,128,2] (or [1,128,128,2]), init_density [nx,nx], c1/c2 [nt,nx,nx].
    Returns (densitys, zero_densitys, velocitys, c1_tiled, c2_tiled, smoke_out_record[T])."""
    nt, nx = c1.shape[0], c1.shape[1]
    T = per_timelength
    ti, si = int(T / nt), int(128 / nx)
    dens0 = np.tile(init_density.reshape(nx, 1, nx, 1), (1, si, 1, si)).reshape(128, 128)
    c1 = np.tile(c1.reshape(nt, 1, nx, 1, nx, 1), (1, ti, 1, si, 1, si)).reshape(T, 128, 128)
    c2 = np.tile(c2.reshape(nt, 1, nx, 1, nx, 1), (1, ti, 1, si, 1, si)).reshape(T, 128, 128)
    vmask = velocity_mask(fluid).astype(np.float64)
    coef = laplace_coefficients(fluid)
    masks, concat, keep = bucket_masks()
    dens = dens0[:-1, :-1].copy()
    zdens = dens.copy()
    v = np.asarray(init_velocity).reshape(128, 128, 2)
    smoke_outs = np.zeros(7)
    densitys, zero_densitys, velocitys, record = [], [], [], []

    def account(zd):
        full = np.zeros((128, 128))
        full[:-1, :-1] = zd
        if np.sum(full * concat) > 0:
            for i, m in enumerate(masks):
                smoke_outs[i] += np.sum(full * m)
            zd = (zd * keep[:-1, :-1]).astype(zd.dtype)
        full = np.zeros((128, 128))
        full[:-1, :-1] = zd
        return zd, full

    velocitys.append(v.astype(np.float64))
    full = np.zeros((128, 128))
    full[:-1, :-1] = dens
    densitys.append(full)
    zdens, zfull = account(zdens)
    zero_densitys.append(zfull)
    record.append(smoke_outs[1] / (np.sum(smoke_outs) + np.sum(zfull)))
    v = v.astype(np.float64)
    for frame in range(T - 1):
        v, _, _ = divergence_free(inject_control(v, c1[frame], c2[frame]), vmask, coef)
        dens = advect(dens, v, dt)
        zdens = advect(zdens, v, dt)
        zdens, zfull = account(zdens)
        full = np.zeros((128, 128))
        full[:-1, :-1] = dens
        densitys.append(full)
        zero_densitys.append(zfull)
        velocitys.append(v.copy())
        record.append(smoke_outs[1] / (np.sum(smoke_outs) + np.sum(zfull)))
    return np.stack(densitys), np.stack(zero_densitys), np.stack(velocitys), c1, c2, np.array(record)
