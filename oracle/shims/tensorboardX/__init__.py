"""Import-only shim (reference diffusion_1d_burgers.py:24). TEST INFRASTRUCTURE ONLY."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass
