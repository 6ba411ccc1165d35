"""Synthetic jellyfish simulations in the on-disk layout of dataset/data_2d.py:45-98 (shared by the golden generator, which
reads them with the unmodified reference reader, and by tests/test_data_jellyfish.py)."""
import os
import pickle

import numpy as np

NORM = dict(vx_max=1.5, vx_min=-1.25, vy_max=2.0, vy_min=-1.75, p_max=3.0, p_min=-2.5)


def write(root, split, sim_id, seed, T=40, S=16):
    rng = np.random.default_rng(seed)
    d = os.path.join(root, split)
    for sub in ("states", "bdry_merged_mask_offsets", "bdry_head_thetas"):
        os.makedirs(os.path.join(d, sub), exist_ok=True)
    with open(os.path.join(d, "normalization_max_min.pkl"), "wb") as fh:
        pickle.dump(NORM, fh)
    states = (rng.standard_normal((T, 3, S, S)) * 2).astype(np.float32)
    states[1, 2, 0, 0] = np.nan                                   # NaNs are zeroed by the reader
    bd = rng.random((T, S - 2, S - 2, 3)).astype(np.float32)
    bd[2, 1, 1, 0] = np.nan
    np.savez(os.path.join(d, "states", "sim_{:06d}.npz".format(sim_id)), a=states)
    np.savez(os.path.join(d, "bdry_merged_mask_offsets", "sim_{:06d}.npz".format(sim_id)), a=bd)
    np.savez(os.path.join(d, "bdry_head_thetas", "sim_{:06d}.npz".format(sim_id)), thetas=rng.random(T).astype(np.float32))
