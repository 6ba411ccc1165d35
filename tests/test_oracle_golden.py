"""Pins the CPU oracle (oracle/*.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import smoke_sampler_oracle as so
from oracle import unet3d_oracle as uo

CASES = {
    "unet_small_c6": dict(dim=32, dim_mults=(1, 2), channels=6),
    "unet_small_c2": dict(dim=32, dim_mults=(1, 2), channels=2),
    "unet_smoke_arch": dict(dim=64, dim_mults=(1, 2, 4), channels=6),
    "unet_jelly_arch": dict(dim=32, dim_mults=(1, 2), channels=7, out_dim=4),
}

# The oracle and the reference run the same ATen CPU kernels in the same order, so they agree to rounding noise.
TOL = 2e-5


@pytest.mark.parametrize("name", list(CASES))
def test_unet_oracle_matches_reference_golden(name, golden_dir):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = uo.UnetCfg(**CASES[name])
    params = uo.make_params(cfg, int(z["seed"]))
    taps = {}
    y = uo.forward(params, cfg, torch.from_numpy(z["x"]), torch.from_numpy(z["t"]), taps=taps)
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())
    for k in z.files:
        if k.startswith("act/"):
            a = torch.from_numpy(z[k])
            assert (taps[k[4:]] - a).abs().max().item() <= TOL * max(1.0, a.abs().max().item()), k


def test_param_inventory_counts():
    # 23 044 870 trainable params (ch 6), 22 956 802 (ch 2); rotary freqs are buffers-like
    def count(cfg):
        return sum(int(np.prod(s)) for k, s in uo.param_shapes(cfg).items() if not k.endswith("rotary_emb.freqs"))
    assert count(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=6)) == 23044870
    assert count(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=2)) == 22956802
    assert count(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=7, out_dim=4)) == 23066692
    assert count(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=7, out_dim=1)) == 23066497


def test_schedules_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "schedules.npz"))
    for name in ("sigmoid", "cosine", "linear"):
        for T in (1000, 200):
            s = so.make_schedule(T, name)
            for k, v in s.items():
                ref = z[f"{name}{T}/{k}"]
                assert np.array_equal(v.numpy(), ref), (name, T, k)  # bit-exact: same fp64 math, same cast


def _design(rescaler, w_energy):
    return lambda x: so.guidance_fn(x, rescaler, w_energy)


@pytest.mark.parametrize("tag,guidance", [("std", "standard"), ("alpha", "standard-alpha")])
def test_p_sample_step_matches_reference(tag, guidance, golden_dir):
    z = np.load(os.path.join(golden_dir, f"sampler_step_{tag}.npz"))
    sched = so.make_schedule(1000, "sigmoid")
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    init = torch.from_numpy(z["init"])
    for t in (999, 500, 1, 0):
        g = lambda k: torch.from_numpy(z[f"t{t}/{k}"])
        pred, x_start = so.p_sample_step(
            sched, g("x"), t, g("eps_joint"), g("eps_w"), g("z"), init, _design(R, float(z["w_energy"])),
            design_guidance=guidance, standard_fixed_ratio=float(z["standard_fixed_ratio"]),
            coeff_ratio=float(z["coeff_ratio"]), w_prob_exp=float(z["w_prob_exp"]))
        assert torch.equal(x_start, g("x_start")), t
        assert torch.equal(pred, g("pred")), t


def _nets(golden_seed_j=11, golden_seed_w=12):
    cj = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6)
    cw = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2)
    pj, pw = uo.make_params(cj, golden_seed_j), uo.make_params(cw, golden_seed_w)
    return lambda x, t: (uo.forward(pj, cj, x, t), uo.forward(pw, cw, x[:, :, 3:5], t))


def test_ddpm_loop_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "sampler_loop_ddpm4.npz"))
    sched = so.make_schedule(4, "sigmoid")
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    init = torch.from_numpy(z["init"])
    torch.manual_seed(42)
    y = so.p_sample_loop(sched, _nets(), (2, 4, 6, 16, 16), init, _design(R, 0.0), 4,
                         design_guidance="standard", standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
    ref = torch.from_numpy(z["y"])
    assert (y - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_ddim_loop_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "sampler_loop_ddim3.npz"))
    sched = so.make_schedule(1000, "sigmoid")
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    init = torch.from_numpy(z["init"])
    nets = _nets()
    torch.manual_seed(43)
    shape = (2, 4, 6, 16, 16)
    x = torch.randn(shape)
    x[:, 0, 0] = init
    for time, time_next in so.ddim_times(1000, 3):
        tt = torch.full((2,), time, dtype=torch.long)
        ej, ew = nets(x, tt)
        noise = torch.randn(shape) if time_next >= 0 else None
        x, _ = so.ddim_step(sched, x, time, time_next, ej, ew, noise, init, _design(R, 0.0), eta=1.0,
                            design_guidance="standard", standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
    ref = torch.from_numpy(z["y"])
    assert (x - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
