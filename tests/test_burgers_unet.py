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


@pytest.mark.parametrize("name", ["burgers_unet_uw", "burgers_unet_g8"])
def test_unet2d_tcgen05_host_logic(name, golden_dir, monkeypatch):
    """TF32 mode with a tcgen05 conv that accepts the shapes (emulated): GroupNorm(1) nets take the kernel's GroupNorm(8)
    statistics through dpc_gn_stats_merge, GroupNorm(8) nets take them directly; same result as the reference golden."""
    from diffphycon_b200 import unet3d, _lib
    from tests import cpu_emulator
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(cpu_emulator, "TC_ACCEPTS", True)
    monkeypatch.setattr(unet3d, "_require_cuda", lambda x: None)
    import diffphycon_b200.burgers_unet as bu
    monkeypatch.setattr(bu, "_require_cuda", lambda x: None)
    calls = []
    orig = _lib.gn_stats_merge
    monkeypatch.setattr(_lib, "gn_stats_merge", lambda *a: (calls.append(a[2:]), orig(*a))[1])
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    net = build(name, int(z["seed"]), "tf32")
    y = net(torch.from_numpy(z["x"]), torch.from_numpy(z["t"]))
    ref = torch.from_numpy(z["y"])
    assert (y - ref).abs().max().item() <= TOL["tf32"] * max(1.0, ref.abs().max().item())
    assert (len(calls) > 0) == (CASES[name]["resnet_block_groups"] != 8)


@pytest.mark.gpu
def test_gn_stats_merge_kernel():
    from diffphycon_b200 import _lib
    g = torch.Generator().manual_seed(0)
    a = torch.randn(5, 8, 2, generator=g, dtype=torch.float64).cuda()
    for gout in (1, 2, 4):
        out = torch.full((5, gout, 2), 0.25, dtype=torch.float64, device="cuda")
        _lib.gn_stats_merge(a.reshape(-1), out.reshape(-1), 5, 8, gout)
        ref = 0.25 + a.reshape(5, gout, 8 // gout, 2).sum(2)
        assert (out - ref).abs().max().item() <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_unet2d_tcgen05_path_is_taken_and_matches_igemm(name, golden_dir):
    """In TF32 mode the 3x3 / 1x1 layers with tileable channel counts run on the tcgen05 kernel (profiler categories conv[tc ...]),
    the rest on the mma.sync implicit GEMM; both are TF32 contractions of the same operands: results agree to accumulation order."""
    from diffphycon_b200 import _lib
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    net = build(name, int(z["seed"]), "tf32").cuda()
    x, t = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["t"]).cuda()
    with _lib.Profiler() as prof:
        y_tc = net(x, t)
    labels = list(prof.summary())
    assert any(k.startswith("conv[tc") for k in labels), labels
    net.use_tcgen05 = False
    y_ig = net(x, t)
    scale = max(1.0, y_ig.abs().max().item())
    assert (y_tc - y_ig).abs().max().item() <= 2e-3 * scale


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
