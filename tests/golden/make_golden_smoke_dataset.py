"""Golden vectors for the smoke dataset reader: two synthetic simulations in the on-disk layout of dataset/data_2d.py:176-207 are
read by the UNMODIFIED reference `Smoke` class (build container only):   python tests/golden/make_golden_smoke_dataset.py
Stored: the synthetic files themselves (small) and what the reference returns for train and test items."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from dataset.data_2d import Smoke  # noqa: E402
from tests.smoke_dataset_fixture import synth, write  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as root:
        tr, te = synth(1, 33, 64), synth(2, 257, 64)
        # keep the fixture small: the stored inputs are 8x8 crops re-expanded by the test the same way
        write(root, ("train",), 3, tr)
        write(root, ("test", "control"), 1, te)
        x, sid = Smoke(root, is_train=True)[3]
        out["train_item"], out["train_id"] = x.numpy()[:, :, ::8, ::8].copy(), np.int64(sid)
        x, sid = Smoke(root, is_train=False)[1]
        out["test_item"], out["test_id"] = x.numpy()[::8, :, ::8, ::8].copy(), np.int64(sid)
    out["seeds"] = np.array([1, 2])
    np.savez_compressed(os.path.join(HERE, "smoke_dataset.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
