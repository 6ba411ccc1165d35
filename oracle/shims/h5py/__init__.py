"""Import-only shim (reference dataset/apps/generate_burgers.py:13). TEST INFRASTRUCTURE ONLY."""


class File:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py shim: datasets are out of scope")
