"""Import the UNMODIFIED reference modules from /root/reference through the shims in oracle/shims.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not exist on the GPU box.
Used by tests/golden/make_golden.py to pin the restated oracle (oracle/*.py) against the reference itself.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("DPC_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffusion"))


def _prepare():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def unet3d_module():
    """reference model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py"""
    _prepare()
    return importlib.import_module("model.video_diffusion_pytorch.video_diffusion_pytorch_conv3d")


def smoke_diffusion_module():
    """reference diffusion/diffusion_2d_smoke.py"""
    _prepare()
    return importlib.import_module("diffusion.diffusion_2d_smoke")


def burgers_unet_module():
    _prepare()
    return importlib.import_module("model.burgers_1d.unet")


def burgers_diffusion_module():
    _prepare()
    return importlib.import_module("diffusion.diffusion_1d_burgers")
