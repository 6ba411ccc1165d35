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
