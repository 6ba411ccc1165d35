"""B200-native ground-truth rollout of the sampled smoke controls — drop-in for the functions the reference's
`inference/inference_2d_smoke.py` star-imports from `dataset/apps/evaluate_solver.py` (cited as es.py:line):
`init_sim_128` (es.py:94-97), `init_velocity_` (es.py:113-115) and `solver` (es.py:205-310), plus a batched entry point
that keeps everything on the device.  The PhiFlow machinery underneath (MAC-grid masks, divergence, CG pressure solve,
pressure gradient, semi-Lagrangian advection, bucket accounting) runs in ONE persistent kernel per trajectory
(`dpc_smoke_rollout`, csrc/smoke_rollout.cu) in fp64 like the reference's NumPy path.  The reference forks one OS
process per trajectory (inference_2d_smoke.py:339-364); here a batch is one launch, one CTA per trajectory.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

N = 127

# (size_y, size_x), (origin_y, origin_x): build_obstacles_pi_128, es.py:32-63
_OBSTACLES_128 = (
    ((1, 96), (16, 16)),
    ((8, 1), (16, 16)), ((16, 1), (40, 16)), ((40, 1), (72, 16)),
    ((8, 1), (16, 112)), ((16, 1), (40, 112)), ((40, 1), (72, 112)),
    ((1, 8), (112, 16)), ((1, 16), (112, 40)), ((1, 16), (112, 72)), ((1, 8), (112, 104)),
    ((16, 1), (64, 48)), ((16, 1), (96, 48)), ((16, 1), (64, 80)), ((16, 1), (96, 80)),
    ((1, 128 - 40 - 40), (40, 40)),
)


class SmokeSimulation:
    """The part of phi.flow.FluidSimulation the rollout needs: a 127x127 domain, open on all sides, with rectangular
    obstacles (phi/flow.py:47-193).  `fluid_mask` [127,127] int8; `velocity_mask` [128,128,2] float32 is 1 on faces whose
    two adjacent cells are fluid, cells outside the domain counting as fluid (phi/flow.py:455-473)."""

    def __init__(self, size: int = N):
        assert size == N, "the rollout kernel is specialised for the reference's 127x127 domain"
        self.fluid_mask = np.ones((size, size), dtype=np.int8)
        self._dev = {}

    def set_obstacle(self, mask_or_size, origin=None):
        (sy, sx), (oy, ox) = mask_or_size, (origin if origin is not None else (0, 0))
        self.fluid_mask[oy:oy + sy, ox:ox + sx] = 0
        self._dev.clear()

    @property
    def velocity_mask(self) -> np.ndarray:
        ext = np.pad(self.fluid_mask.astype(np.float32), 1, constant_values=1)
        my = np.minimum(ext[1:, 1:], ext[:-1, 1:])
        mx = np.minimum(ext[1:, 1:], ext[1:, :-1])
        return np.stack([mx, my], axis=-1)

    def device_masks(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.fluid_mask.copy()).to(device).contiguous(),
                              torch.from_numpy(self.velocity_mask.copy()).to(device).contiguous())
        return self._dev[key]


def init_sim_128() -> SmokeSimulation:
    sim = SmokeSimulation()
    for size, origin in _OBSTACLES_128:
        sim.set_obstacle(size, origin)
    return sim


def init_velocity_() -> np.ndarray:
    """es.py:103-115: the staggered field of a uniform flow vx = 0, vy = 0.8, shape [1,128,128,2] float32."""
    v = np.empty((1, 128, 128, 2), np.float32)
    v[..., 0] = 0.0
    v[..., 1] = 0.8
    return v


@torch.no_grad()
@_lib.device_guarded
def solver_batch(sim: SmokeSimulation, init_velocity, init_density, c1, c2, per_timelength: int, dt: float = 1.0,
                 accuracy: float = 1e-8, max_iterations: int = 500):
    """Batched rollout on the device.  init_velocity [B,128,128,2] (or [128,128,2] / [1,128,128,2] shared),
    init_density [B,nx,nx], c1/c2 [B,nt,nx,nx] (float32 CUDA tensors).  Returns a dict of CUDA tensors:
    densitys / zero_densitys [B,T,128,128] fp32, velocitys [B,T,128,128,2] fp64, smoke_out [B,T] fp64,
    iterations [B,T] int32."""
    dev = c1.device
    if dev.type != "cuda":
        raise RuntimeError("diffphycon_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    B, nt, nx = c1.shape[0], c1.shape[1], c1.shape[2]
    T = int(per_timelength)
    assert 128 % nx == 0 and T % nt == 0 and c2.shape == c1.shape and init_density.shape == (B, nx, nx)
    v0 = torch.as_tensor(init_velocity, dtype=torch.float32, device=dev).reshape(-1, 128, 128, 2)
    if v0.shape[0] == 1 and B > 1:
        v0 = v0.expand(B, 128, 128, 2)
    v0 = v0.contiguous()
    fluid, vmask = sim.device_masks(dev)
    f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
    out = dict(densitys=torch.empty(B, T, 128, 128, dtype=torch.float32, device=dev),
               zero_densitys=torch.empty(B, T, 128, 128, dtype=torch.float32, device=dev),
               velocitys=torch.empty(B, T, 128, 128, 2, dtype=torch.float64, device=dev),
               smoke_out=torch.empty(B, T, dtype=torch.float64, device=dev),
               iterations=torch.empty(B, T, dtype=torch.int32, device=dev))
    vel_ws = torch.empty(B * 2 * 128 * 128 * 2, dtype=torch.float64, device=dev)
    x_ws = torch.empty(B * N * N, dtype=torch.float64, device=dev)
    dens_ws = torch.empty(B * 4 * N * N, dtype=torch.float32, device=dev)
    _lib.smoke_rollout(fluid, vmask, v0, f32(init_density), f32(c1), f32(c2), vel_ws, x_ws, dens_ws, out["densitys"],
                       out["zero_densitys"], out["velocitys"], out["smoke_out"], out["iterations"], B, nt, nx, T, dt,
                       accuracy, max_iterations)
    return out


def solver(sim: SmokeSimulation, init_velocity, init_density, c1, c2, per_timelength, dt=1):
    """es.py:205-310 with its NumPy call surface: one trajectory in, the reference's six arrays out
    (densitys, zero_densitys, velocitys, c1 tiled, c2 tiled, smoke_out_record tiled to [T,128,128])."""
    dev = torch.device("cuda", torch.cuda.current_device())
    c1 = np.asarray(c1, dtype=np.float32)
    c2 = np.asarray(c2, dtype=np.float32)
    nt, nx = c1.shape[0], c1.shape[1]
    T = int(per_timelength)
    out = solver_batch(sim, np.asarray(init_velocity, dtype=np.float32).reshape(1, 128, 128, 2),
                       torch.from_numpy(np.asarray(init_density, dtype=np.float32).reshape(1, nx, nx)).to(dev),
                       torch.from_numpy(c1[None]).to(dev), torch.from_numpy(c2[None]).to(dev), T, dt)
    ti, si = int(T / nt), int(128 / nx)
    c1t = np.tile(c1.reshape(nt, 1, nx, 1, nx, 1), (1, ti, 1, si, 1, si)).reshape(T, 128, 128)
    c2t = np.tile(c2.reshape(nt, 1, nx, 1, nx, 1), (1, ti, 1, si, 1, si)).reshape(T, 128, 128)
    rec = out["smoke_out"][0].cpu().numpy()
    return (out["densitys"][0].double().cpu().numpy(), out["zero_densitys"][0].double().cpu().numpy(),
            out["velocitys"][0].cpu().numpy(), c1t, c2t, np.tile(rec[:, None, None], (1, 128, 128)))
