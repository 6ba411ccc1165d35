"""world_size-2 `gloo` test of the sharded sampling path on CPU: shard bounds, the global-noise slicing that makes an
N-rank run reproduce the 1-rank trajectories, and the single all-gather.  The kernels are replaced by the torch
emulation of the C-ABI (tests/cpu_emulator.py) inside the worker processes; no GPU is involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffphycon_b200.distributed import shard_bounds


def test_shard_bounds_cover_batch():
    for batch in (1, 2, 5, 64, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _build():
    import diffphycon_b200 as dpc
    from oracle import unet3d_oracle as uo
    cj = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6)
    cw = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2)
    mj = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=2)
    mj.load_state_dict(uo.make_params(cj, 11))
    mw.load_state_dict(uo.make_params(cw, 12))
    mj.precision = mw.precision = "3xtf32"
    return dpc.GaussianDiffusion([mj, mw], image_size=16, frames=4, timesteps=3, sampling_timesteps=3, eval_2ddpm=True,
                                 standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97), dpc


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import cpu_emulator
    cpu_emulator.install_raw()
    from diffphycon_b200.distributed import sample_sharded
    d, dpc = _build()
    g = torch.Generator().manual_seed(3)
    init = torch.rand(3, 16, 16, generator=g) / 2
    torch.manual_seed(77)
    y = sample_sharded(d, 3, design_fn=dpc.StockSmokeGuidance(), init=init, global_noise=True)
    torch.manual_seed(77)
    yc = sample_sharded(d, 3, design_fn=dpc.StockSmokeGuidance(), init=init, global_noise=True, gather_channels=slice(3, 5))
    torch.save((y, yc), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sampling_reproduces_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y0, yc0 = torch.load(tmp_path / "rank0.pt")
    y1, yc1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(y0, y1) and torch.equal(yc0, yc1)          # every rank holds the full gathered result
    assert y0.shape == (3, 4, 6, 16, 16) and yc0.shape == (3, 4, 2, 16, 16)
    assert torch.equal(yc0, y0[:, :, 3:5])
    # single-process run with the same seed: identical trajectories (global noise stream is sliced, not re-seeded)
    from tests import cpu_emulator
    mpatch = pytest.MonkeyPatch()
    try:
        cpu_emulator.install(mpatch)
        from diffphycon_b200 import unet3d
        mpatch.setattr(unet3d, "_require_cuda", lambda x: None)
        d, dpc = _build()
        g = torch.Generator().manual_seed(3)
        init = torch.rand(3, 16, 16, generator=g) / 2
        torch.manual_seed(77)
        ref = d.sample(batch_size=3, design_fn=dpc.StockSmokeGuidance(), init=init)
    finally:
        mpatch.undo()
    # (the emulator's BLAS blocks differently for batch 1-2 vs 3; a 3-step schedule amplifies that ~1e3x)
    assert torch.allclose(y0, ref, atol=2e-3, rtol=0)
