"""Multi-GPU sampling: independent trajectories sharded over ranks, one all-gather of the result.

The reference samples on a single GPU (inference/inference_2d_smoke.py:533); its trajectories never interact, so each
rank runs the unmodified sampling loop on its contiguous slice of the batch with a full model replica and there is no
per-step communication.  With `global_noise=True` every rank draws the GLOBAL noise tensor with the same generator
state and keeps its slice, so an N-rank run reproduces the single-rank trajectories exactly (costs B_global/B_local
times the RNG work); with False each rank draws only its own noise (seed it per rank).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slices: the first (batch % world) ranks get one extra trajectory."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _GlobalNoise:
    """Context manager that makes `diffusion.sample_noise` draw the global batch and return this rank's rows."""

    def __init__(self, diffusion, batch_global: int, lo: int, hi: int):
        self.d, self.bg, self.lo, self.hi = diffusion, batch_global, lo, hi

    def __enter__(self):
        self.orig = self.d.sample_noise

        def sample_noise(shape, device):
            full = self.orig([self.bg, *shape[1:]], device)
            return full[self.lo:self.hi].contiguous()
        self.d.sample_noise = sample_noise
        self.d._noise_key = (self.bg, self.lo, self.hi)     # part of the CUDA-graph cache key (the noise draw is captured)
        return self

    def __exit__(self, *a):
        self.d.sample_noise = self.orig
        self.d._noise_key = None


@torch.no_grad()
def sample_sharded(diffusion, batch_size: int, design_fn=None, design_guidance: str = "standard", init=None, init_u=None,
                   control=None, low=None, group: Optional[dist.ProcessGroup] = None, global_noise: bool = True,
                   gather_channels: Optional[slice] = None):
    """`GaussianDiffusion.sample` over all ranks of `group`.  `init` (and `init_u`, `control`, `low` when given) are the
    GLOBAL tensors, identical on every rank.  Returns the gathered samples [batch_size, F, C', H, W] on every rank;
    `gather_channels=slice(3, 5)` gathers the sampled controls only (1.05 MB per trajectory)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(batch_size, world, rank)
    sl = lambda t: None if t is None else t[lo:hi]
    if hi > lo:
        if global_noise and world > 1:
            with _GlobalNoise(diffusion, batch_size, lo, hi):
                local = diffusion.sample(batch_size=hi - lo, design_fn=design_fn, design_guidance=design_guidance,
                                         init=sl(init), init_u=sl(init_u), control=sl(control), low=sl(low))
        else:
            local = diffusion.sample(batch_size=hi - lo, design_fn=design_fn, design_guidance=design_guidance,
                                     init=sl(init), init_u=sl(init_u), control=sl(control), low=sl(low))
    else:
        dev = diffusion.betas.device
        local = torch.empty(0, diffusion.frames, diffusion.channels, diffusion.image_size, diffusion.image_size, device=dev)
    if gather_channels is not None:
        local = local[:, :, gather_channels].contiguous()
    if world == 1:
        return local
    # the single collective of the path: all-gather of the result (uneven shards are padded to the largest)
    max_rows = shard_bounds(batch_size, world, 0)[1]
    padded = torch.zeros(max_rows, *local.shape[1:], dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bucket = torch.empty(world * max_rows, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(bucket, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(batch_size, world, r)
        parts.append(bucket[r * max_rows: r * max_rows + (b - a)])
    return torch.cat(parts, dim=0)
