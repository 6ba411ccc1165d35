"""Smoke dataset reader (diffphycon_b200/data_smoke.py) against what the unmodified reference reader returned for the same
synthetic files (tests/golden/make_golden_smoke_dataset.py): bit-exact (file I/O, permutes, one division)."""
import os

import numpy as np
import torch

from diffphycon_b200.data_smoke import Smoke
from tests.smoke_dataset_fixture import synth, write


def test_smoke_reader_matches_reference(tmp_path, golden_dir):
    z = np.load(os.path.join(golden_dir, "smoke_dataset.npz"))
    root = str(tmp_path)
    write(root, ("train",), 3, synth(int(z["seeds"][0]), 33, 64))
    write(root, ("test", "control"), 1, synth(int(z["seeds"][1]), 257, 64))
    ds = Smoke(root, is_train=True)
    x, sid = ds[3]
    assert x.shape == (32, 6, 64, 64) and sid == int(z["train_id"]) and len(ds) == 20000
    assert torch.equal(x[:, :, ::8, ::8], torch.from_numpy(z["train_item"]))
    assert torch.equal(ds.RESCALER.reshape(-1), torch.tensor([2, 18, 20, 16, 20, 1]))
    dt = Smoke(root, is_train=False)
    x, sid = dt[1]
    assert x.shape == (256, 6, 64, 64) and sid == int(z["test_id"]) and len(dt) == 50
    assert torch.equal(x[::8, :, ::8, ::8], torch.from_numpy(z["test_item"]))
