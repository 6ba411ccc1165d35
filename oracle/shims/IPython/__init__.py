"""Import-only shim (reference diffusion_2d_smoke.py:35). TEST INFRASTRUCTURE ONLY."""


def embed(*a, **k):
    raise RuntimeError("IPython shim")
