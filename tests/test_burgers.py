"""Burgers finite-difference rollout: the NumPy oracle against the unmodified reference's golden (CPU), and the CUDA
kernel against both (GPU).  Everything is float32 in the reference's evaluation order: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import burgers_oracle as bo


def test_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "burgers_rollout.npz"))
    tr = bo.burgers_numeric_solve_free(z["u0"], z["f"], 0.01, 1.0)
    assert tr.shape == z["traj"].shape == (4, 11, 128)
    assert np.array_equal(tr, z["traj"])


@pytest.mark.gpu
def test_kernel_matches_reference_golden(golden_dir):
    from diffphycon_b200.burgers import burgers_numeric_solve_free
    z = np.load(os.path.join(golden_dir, "burgers_rollout.npz"))
    tr = burgers_numeric_solve_free(torch.from_numpy(z["u0"]).cuda(), torch.from_numpy(z["f"]).cuda(), visc=0.01, T=1.0,
                                    dt=1e-4, num_t=10)
    assert tr.shape == (4, 11, 128)
    assert np.array_equal(tr.cpu().numpy(), z["traj"])


@pytest.mark.gpu
def test_kernel_matches_oracle_other_sizes():
    from diffphycon_b200.burgers import burgers_numeric_solve_free
    rng = np.random.default_rng(1)
    for N, s, Nt, T in ((3, 64, 5, 0.5), (2, 200, 8, 0.4)):
        u0 = (rng.standard_normal((N, s)) * 0.3).astype(np.float32)
        f = (rng.standard_normal((N, Nt, s)) * 0.5).astype(np.float32)
        ref = bo.burgers_numeric_solve_free(u0, f, 0.01, T, 1e-4, Nt)
        got = burgers_numeric_solve_free(torch.from_numpy(u0).cuda(), torch.from_numpy(f).cuda(), 0.01, T, 1e-4, Nt)
        assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.gpu
def test_burgers_metric_matches_restatement(golden_dir):
    """diffphycon_b200.burgers_metric (utils.py:1203-1284) on the device rollout: the controlled trajectory is the golden one of
    the unmodified reference solver, so J_actual / control_energy equal the reference formulas applied to the golden."""
    from diffphycon_b200.burgers_metric import burgers_metric
    z = np.load(os.path.join(golden_dir, "burgers_rollout.npz"))
    u0, f, traj = (torch.from_numpy(z[k]) for k in ("u0", "f", "traj"))
    u_target = traj.clone()
    u_target[:, -1] += 0.05 * torch.sin(torch.linspace(0, 6.28, traj.shape[-1]))      # target differs from the outcome
    J, E = burgers_metric(u_target.cuda(), f.cuda())
    ref_J = (traj[:, -1] - u_target[:, -1]).square().mean(-1)
    assert torch.allclose(J.cpu(), ref_J, rtol=1e-4, atol=1e-7) and torch.allclose(E.cpu(), f.square().sum((-1, -2)), rtol=1e-6)
    Nx = f.shape[2]
    fz = f.clone()
    fz[:, :, Nx // 4:(3 * Nx) // 4] = 0
    (mse, mse_med, mae, mae_med, nmse, nmae), E2 = burgers_metric(u_target.cuda(), f.cuda(), partial_control='front_rear_quarter',
                                                                  partially_observed='front_rear_quarter', report_all=True)
    assert torch.allclose(E2.cpu(), fz.square().sum((-1, -2)), rtol=1e-6)
    assert mse.shape == (f.shape[0],) and torch.isfinite(torch.stack([mse, mse_med, mae, mae_med, nmse, nmae])).all()
