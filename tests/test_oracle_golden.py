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


# ---- the benchmarked shape (smoke 64x64, 32 frames, dim 64 (1,2,4)): tests/golden/make_golden_metric_shape.py ----------
METRIC_SEEDS = {"joint": (6, 31), "prior": (2, 33)}


def metric_input(channels, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 32, channels, 64, 64, generator=g)


@pytest.mark.parametrize("tag", ["joint", "prior"])
def test_unet_oracle_matches_reference_at_metric_shape(tag, golden_dir):
    """The oracle against strided subsamples of the UNMODIFIED reference's output and stage activations at
    [1,32,C,64,64] — the shape bench.py measures."""
    z = np.load(os.path.join(golden_dir, "metric_shape.npz"))
    ch, seed = METRIC_SEEDS[tag]
    cfg = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=ch)
    taps = {}
    y = uo.forward(uo.make_params(cfg, seed), cfg, metric_input(ch, seed + 1), torch.tensor([int(z[f"{tag}/t"])]), taps=taps)
    ref = torch.from_numpy(z[f"{tag}/y"])
    assert (y[:, ::8, :, ::8, ::8] - ref).abs().max().item() <= TOL * max(1.0, float(z[f"{tag}/y_absmax"]))
    for k in z.files:
        if k.startswith(f"{tag}/act/") and k.split("/", 2)[2] in taps:
            nm = k.split("/", 2)[2]
            a = torch.from_numpy(z[k])
            err = (taps[nm][:, :, ::8, ::8, ::8] - a).abs().max().item()
            assert err <= TOL * max(1.0, float(z[f"{tag}/absmax/{nm}"])), (nm, err)


def test_p_sample_oracle_matches_reference_at_metric_shape(golden_dir):
    z = np.load(os.path.join(golden_dir, "metric_shape.npz"))
    t = int(z["p_sample/t"])
    cj = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=6)
    cw = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=2)
    init = torch.from_numpy(z["p_sample/init"])
    x = metric_input(6, 78)
    x[:, 0, 0] = init
    tt = torch.tensor([t])
    ej = uo.forward(uo.make_params(cj, 31), cj, x, tt)
    ew = uo.forward(uo.make_params(cw, 33), cw, x[:, :, 3:5], tt)
    torch.manual_seed(1234 + t)
    noise = torch.randn(1, 32, 6, 64, 64)
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    sched = so.make_schedule(1000, "sigmoid")
    pred, x_start = so.p_sample_step(sched, x, t, ej, ew, noise, init, _design(R, 0.0), design_guidance="standard",
                                     standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
    # an eps difference d moves x_start by sqrt(1/abar_t - 1) * d (smoke.py:576-580); 2e-5 oracle-vs-reference noise on eps
    amp = float(sched["sqrt_recipm1_alphas_cumprod"][t])
    assert (x_start[:, ::8, :, ::8, ::8] - torch.from_numpy(z["p_sample/x_start"])).abs().max().item() <= TOL * amp + 1e-6
    assert (pred[:, ::8, :, ::8, ::8] - torch.from_numpy(z["p_sample/pred"])).abs().max().item() <= TOL * amp + 1e-6
