"""Import-only shim (reference diffusion_2d_smoke.py:30). Never touched by sampling. TEST INFRASTRUCTURE ONLY."""


class Accelerator:
    def __init__(self, *a, **k):
        raise RuntimeError("accelerate shim: training is out of scope")
