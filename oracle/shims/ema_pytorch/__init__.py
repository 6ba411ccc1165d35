"""Import-only shim (reference diffusion_2d_smoke.py:28). Never touched by sampling. TEST INFRASTRUCTURE ONLY."""


class EMA:
    def __init__(self, *a, **k):
        raise RuntimeError("ema_pytorch shim: training is out of scope")
