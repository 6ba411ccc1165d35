"""Jellyfish dataset reader (diffphycon_b200/data_jellyfish.py) against what the unmodified reference reader returned for the
same synthetic files (tests/golden/make_golden_jellyfish_dataset.py): bit-exact."""
import os

import numpy as np
import torch

from diffphycon_b200.data_jellyfish import Jellyfish, normalization
from tests.jellyfish_dataset_fixture import NORM, write


def test_jellyfish_reader_matches_reference(tmp_path, golden_dir):
    z = np.load(os.path.join(golden_dir, "jellyfish_dataset.npz"))
    root = str(tmp_path)
    write(root, "train_data", 2, seed=5)
    write(root, "test_data", 1, seed=6)
    cases = {"train": Jellyfish("jellyfish", root, is_train=True)[2 * 20 + 7],
             "test": Jellyfish("jellyfish", root, is_train=False)[1],
             "pipeline": Jellyfish("jellyfish", root, is_train=False, for_pipeline=True)[1 * 20 + 3],
             "train_pressure": Jellyfish("jellyfish", root, is_train=True, only_vis_pressure=True)[2 * 20]}
    for name, item in cases.items():
        n = len([k for k in z.files if k.startswith(name + "/")])
        assert len(item) == n, name
        for i, v in enumerate(item):
            ref = z[f"{name}/{i}"]
            got = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            assert got.shape == ref.shape and np.array_equal(got, ref), (name, i)
    assert len(Jellyfish("jellyfish", root, is_train=True)) == 1000 * 20 and len(Jellyfish("jellyfish", root, is_train=False)) == 100
    assert normalization(root) == NORM
