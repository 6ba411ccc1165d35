"""Golden vectors for the jellyfish dataset reader: synthetic simulations in the on-disk layout of dataset/data_2d.py:45-98 are
read by the UNMODIFIED reference `Jellyfish` class (build container only):  python tests/golden/make_golden_jellyfish_dataset.py"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from dataset.data_2d import Jellyfish  # noqa: E402
from tests.jellyfish_dataset_fixture import write  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write(root, "train_data", 2, seed=5)
        write(root, "test_data", 1, seed=6)
        with contextlib.redirect_stdout(io.StringIO()):
            cases = {"train": Jellyfish("jellyfish", root, is_train=True)[2 * 20 + 7],
                     "test": Jellyfish("jellyfish", root, is_train=False)[1],
                     "pipeline": Jellyfish("jellyfish", root, is_train=False, for_pipeline=True)[1 * 20 + 3],
                     "train_pressure": Jellyfish("jellyfish", root, is_train=True, only_vis_pressure=True)[2 * 20]}
        for name, item in cases.items():
            for i, v in enumerate(item):
                out[f"{name}/{i}"] = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "jellyfish_dataset.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
