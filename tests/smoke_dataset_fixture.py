"""Synthetic smoke simulations in the on-disk layout of dataset/data_2d.py:176-207 (shared by the golden generator, which reads
them with the unmodified reference reader, and by tests/test_data_smoke.py, which reads them with the engine's reader)."""
import os

import numpy as np


def synth(seed, T, size=64):
    rng = np.random.default_rng(seed)
    return dict(Density=rng.random((size, size, 1, T), dtype=np.float32),
                Velocity=(rng.standard_normal((size, size, 2, T)) * 3).astype(np.float32),
                Control=(rng.standard_normal((size, size, 2, T)) * 2).astype(np.float32),
                Smoke=(rng.random((T, 8)) + 0.1).astype(np.float32))


def write(root, sub, sim_id, arrays):
    d = os.path.join(root, *sub, "sim_{:06d}".format(sim_id))
    os.makedirs(d)
    for k, v in arrays.items():
        np.save(os.path.join(d, k + ".npy"), v)
