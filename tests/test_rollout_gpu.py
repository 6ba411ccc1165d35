"""GPU parity of the persistent rollout kernel (through the C-ABI) against golden vectors of the unmodified reference
solver and against the NumPy oracle on a longer, batched run.  fp64 CG: velocities within 1e-8 of the field scale
(reduction order differs from NumPy's pairwise sums); densities are float32 fields: 2e-6."""
import os

import numpy as np
import pytest
import torch

from diffphycon_b200 import smoke_rollout as sr
from oracle import smoke_rollout_oracle as ro

pytestmark = pytest.mark.gpu


def test_rollout_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "smoke_rollout.npz"))
    sim = sr.init_sim_128()
    assert np.array_equal(sim.fluid_mask, z["fluid_mask"]) and np.array_equal(sim.velocity_mask, z["velocity_mask"])
    d, zd, vs, c1t, c2t, rec = sr.solver(sim, z["init_velocity"], z["init_density"], z["c1"], z["c2"], 4)
    assert d.shape == (4, 128, 128) and vs.shape == (4, 128, 128, 2) and rec.shape == (4, 128, 128)
    assert np.abs(vs - z["velocitys"]).max() <= 1e-8 * np.abs(z["velocitys"]).max()
    assert np.abs(d - z["densitys"]).max() <= 2e-6
    assert np.abs(zd - z["zero_densitys"]).max() <= 2e-6
    assert np.allclose(rec[:, 0, 0], z["smoke_out_record"], rtol=1e-5, atol=1e-12)


def test_batched_rollout_matches_oracle():
    """Three different trajectories in one launch, 6 frames (time tiling 2), checked against the NumPy oracle; also checks
    that every pressure solve hit the reference's 500-iteration cap or converged below 1e-8."""
    rng = np.random.default_rng(3)
    B, nt, nx, T = 3, 3, 64, 6
    c1 = (rng.standard_normal((B, nt, nx, nx)) * 0.5).astype(np.float32)
    c2 = (rng.standard_normal((B, nt, nx, nx)) * 0.5 + 0.2).astype(np.float32)
    dens = rng.random((B, nx, nx)).astype(np.float32)
    sim = sr.init_sim_128()
    out = sr.solver_batch(sim, sr.init_velocity_(), torch.from_numpy(dens).cuda(), torch.from_numpy(c1).cuda(),
                          torch.from_numpy(c2).cuda(), T)
    its = out["iterations"].cpu().numpy()
    assert (its[:, 0] == 0).all() and (its[:, 1:] <= 500).all() and (its[:, 1:] > 0).all()
    fluid = ro.fluid_mask_128()
    for b in range(B):
        d, zd, vs, _, _, rec = ro.solver(fluid, sr.init_velocity_()[0], dens[b], c1[b], c2[b], T)
        assert np.abs(out["velocitys"][b].cpu().numpy() - vs).max() <= 1e-8 * np.abs(vs).max()
        assert np.abs(out["densitys"][b].cpu().numpy() - d).max() <= 2e-6
        assert np.abs(out["zero_densitys"][b].cpu().numpy() - zd).max() <= 2e-6
        assert np.allclose(out["smoke_out"][b].cpu().numpy(), rec, rtol=1e-5, atol=1e-12)


def test_rollout_properties_full_length():
    """Full-size run (256 frames from 32 control frames, like inference_2d_smoke.py:309): finite, mass never created by
    advection + zeroing, masked faces stay zero, zero control on a quiescent field keeps everything at rest."""
    sim = sr.init_sim_128()
    B, nt, nx, T = 2, 32, 64, 256
    g = torch.Generator().manual_seed(0)
    c1 = torch.randn(B, nt, nx, nx, generator=g).cuda() * 0.3
    c2 = torch.randn(B, nt, nx, nx, generator=g).cuda() * 0.3
    c1[1].zero_()
    c2[1].zero_()
    dens = torch.rand(B, nx, nx, generator=g).cuda()
    v0 = np.zeros((1, 128, 128, 2), np.float32)
    out = sr.solver_batch(sim, v0, dens, c1, c2, T)
    for k in ("densitys", "zero_densitys", "velocitys", "smoke_out"):
        assert torch.isfinite(out[k]).all(), k
    vmask = torch.from_numpy(sim.velocity_mask).cuda().double()
    assert (out["velocitys"] * (1 - vmask)).abs().max().item() == 0.0
    assert out["velocitys"][1].abs().max().item() == 0.0                      # trajectory 1: no control, fluid at rest
    assert torch.equal(out["densitys"][1, 0], out["densitys"][1, -1])        # ... so its density never moves
    assert (out["zero_densitys"] <= out["densitys"] + 1e-6).all()
    assert ((out["smoke_out"] >= 0) & (out["smoke_out"] <= 1)).all()


@pytest.mark.parametrize("cs", [2, 4, 8])
def test_rollout_every_cluster_size_matches_golden(cs, golden_dir, monkeypatch):
    """The launcher picks 2, 4 or 8 CTAs per trajectory from the batch size (the widest cluster that keeps all trajectories in one
    wave).  The cluster size only changes the order in which the CTA totals of the dot products are summed: every size must
    hold the reference golden at 1e-8, and a batch must give every trajectory the same result as a single-trajectory launch."""
    monkeypatch.setenv("DPC_ROLLOUT_CLUSTER", str(cs))
    z = np.load(os.path.join(golden_dir, "smoke_rollout.npz"))
    sim = sr.init_sim_128()
    d, zd, vs, _, _, rec = sr.solver(sim, z["init_velocity"], z["init_density"], z["c1"], z["c2"], 4)
    assert np.abs(vs - z["velocitys"]).max() <= 1e-8 * np.abs(z["velocitys"]).max()
    assert np.abs(d - z["densitys"]).max() <= 2e-6 and np.abs(zd - z["zero_densitys"]).max() <= 2e-6
    # batched launch (3 copies + a different 4th trajectory): bit-identical per trajectory, whatever its neighbours are
    c1 = torch.from_numpy(np.stack([z["c1"]] * 3 + [z["c1"][::-1].copy()])).float().cuda()
    c2 = torch.from_numpy(np.stack([z["c2"]] * 3 + [z["c2"][::-1].copy()])).float().cuda()
    dens = torch.from_numpy(np.stack([z["init_density"]] * 4)).float().cuda()
    out = sr.solver_batch(sim, z["init_velocity"][None], dens, c1, c2, 4)
    assert torch.equal(out["velocitys"][0], out["velocitys"][1]) and torch.equal(out["velocitys"][0], out["velocitys"][2])
    assert np.abs(out["velocitys"][0].cpu().numpy() - vs).max() == 0.0
