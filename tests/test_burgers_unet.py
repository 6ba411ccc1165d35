"""Burgers `Unet2D`: host logic through the CPU emulation of the C-ABI (not gpu) and the CUDA kernels (gpu), both against
golden vectors of the unmodified reference model/burgers_1d/unet.py (tests/golden/make_golden_burgers_unet.py).
Tolerances as for the 3-D U-Net: TF32 contraction class 1e-2 of the output scale, 3xTF32 2e-4."""
import os

import numpy as np
import pytest
import torch

from diffphycon_b200.burgers_unet import Unet2D
from oracle import param_gen

CASES = {
    "burgers_unet_uw": dict(dim=64, dim_mults=(1, 2, 4), channels=2, resnet_block_groups=1),
    "burgers_unet_w": dict(dim=32, dim_mults=(1, 2, 4, 8), channels=2, resnet_block_groups=1),
    "burgers_unet_g8": dict(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=8, out_dim=3),
}
TOL = {"tf32": 1e-2, "3xtf32": 2e-4}


def build(name, seed, precision):
    net = Unet2D(**CASES[name])
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(param_gen.make_params(shapes, seed), strict=True)
    net.precision = precision
    return net


@pytest.mark.parametrize("name", list(CASES))
def test_unet2d_host_logic_matches_reference_golden(name, golden_dir, monkeypatch):
    from diffphycon_b200 import unet3d
    from tests import cpu_emulator
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(unet3d, "_require_cuda", lambda x: None)
    import diffphycon_b200.burgers_unet as bu
    monkeypatch.setattr(bu, "_require_cuda", lambda x: None)
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    net = build(name, int(z["seed"]), "3xtf32")
    y = net(torch.from_numpy(z["x"]), torch.from_numpy(z["t"]))
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
@pytest.mark.parametrize("name", list(CASES))
def test_unet2d_forward_matches_reference_golden(name, precision, golden_dir):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    net = build(name, int(z["seed"]), precision).cuda()
    y = net(torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["t"]).cuda()).cpu()
    ref = torch.from_numpy(z["y"])
    err = (y - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    assert err <= TOL[precision], err
