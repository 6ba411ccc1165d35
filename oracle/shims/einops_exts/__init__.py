"""Import shim for `einops-exts==0.0.4` (reference environment.yaml:40) — TEST INFRASTRUCTURE ONLY.

The reference model imports `rearrange_many` / `check_shape`
(model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py:17).
Both are pure reshapes: `rearrange_many` maps `einops.rearrange` over a sequence.
"""
from einops import rearrange


def rearrange_many(tensors, pattern, **kwargs):
    return tuple(rearrange(t, pattern, **kwargs) for t in tensors)


def check_shape(tensor, pattern, **kwargs):
    return rearrange(tensor, f"{pattern} -> {pattern}", **kwargs)
