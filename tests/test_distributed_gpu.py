"""Two-rank NCCL run of diffphycon_b200.distributed.sample_sharded on two GPUs of one box: the sharded run must reproduce the
single-rank trajectories (global noise stream sliced per rank; trajectories never interact) and every rank must
hold the gathered result.  Skipped on a box with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _build(device):
    import diffphycon_b200 as dpc
    from oracle import unet3d_oracle as uo
    cj = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6)
    cw = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2)
    mj = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=32, dim_mults=(1, 2), channels=2)
    mj.load_state_dict(uo.make_params(cj, 11))
    mw.load_state_dict(uo.make_params(cw, 12))
    d = dpc.GaussianDiffusion([mj, mw], image_size=16, frames=4, timesteps=3, sampling_timesteps=3, standard_fixed_ratio=1e5,
                              coeff_ratio=0.0, eval_2ddpm=True, w_prob_exp=0.97).to(device)
    return d, dpc


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from diffphycon_b200.distributed import sample_sharded
    d, dpc = _build(dev)
    g = torch.Generator().manual_seed(3)
    init = (torch.rand(5, 16, 16, generator=g) / 2).to(dev)
    torch.manual_seed(77)
    y = sample_sharded(d, 5, design_fn=dpc.StockSmokeGuidance(), init=init, global_noise=True)
    torch.manual_seed(77)
    yc = sample_sharded(d, 5, design_fn=dpc.StockSmokeGuidance(), init=init, global_noise=True, gather_channels=slice(3, 5))
    torch.save((y.cpu(), yc.cpu()), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_sampling_reproduces_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y0, yc0 = torch.load(tmp_path / "rank0.pt")
    y1, yc1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(y0, y1) and torch.equal(yc0, yc1)          # every rank holds the full gathered result
    assert y0.shape == (5, 4, 6, 16, 16) and torch.equal(yc0, y0[:, :, 3:5])
    d, dpc = _build(torch.device("cuda", 0))
    g = torch.Generator().manual_seed(3)
    init = (torch.rand(5, 16, 16, generator=g) / 2).cuda()
    torch.manual_seed(77)
    ref = d.sample(batch_size=5, design_fn=dpc.StockSmokeGuidance(), init=init).cpu()
    # bit for bit: the same per-trajectory arithmetic on the same noise rows, whichever rank owns the trajectory
    assert torch.equal(y0, ref), (y0 - ref).abs().max().item()
