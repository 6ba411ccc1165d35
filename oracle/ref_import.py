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


# ----------------------------------------------------------------------------------------------------------------
# phi / evaluate_solver: the vendored PhiFlow 1.0.x predates NumPy 1.23 (indexing with a *list* of slices) and
# Python 3.10 (collections.Iterable).  Instead of copying and patching its sources, the UNMODIFIED files are compiled
# through an AST hook at import time: every `x[expr]` whose index is a general expression becomes `x[__phi_idx__(expr)]`
# and __phi_idx__ turns a list that contains slices / None / Ellipsis into the tuple old NumPy treated it as.
# ----------------------------------------------------------------------------------------------------------------
import ast  # noqa: E402
import builtins  # noqa: E402
import collections  # noqa: E402
import collections.abc  # noqa: E402
from importlib.abc import MetaPathFinder  # noqa: E402
from importlib.machinery import PathFinder, SourceFileLoader  # noqa: E402
from unittest import mock  # noqa: E402


def _phi_idx(i):
    if isinstance(i, list) and any(isinstance(e, slice) or e is None or e is Ellipsis for e in i):
        return tuple(i)
    return i


class _IndexFix(ast.NodeTransformer):
    def visit_Subscript(self, node):
        self.generic_visit(node)
        if isinstance(node.slice, (ast.Name, ast.BinOp, ast.List, ast.ListComp, ast.Attribute, ast.Call)):
            node.slice = ast.Call(func=ast.Name(id="__phi_idx__", ctx=ast.Load()), args=[node.slice], keywords=[])
        return node


class _FixLoader(SourceFileLoader):
    def source_to_code(self, data, path, *, _optimize=-1):
        tree = _IndexFix().visit(ast.parse(data, path))
        ast.fix_missing_locations(tree)
        return compile(tree, path, "exec", dont_inherit=True, optimize=_optimize)


class _FixFinder(MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if not (fullname == "phi" or fullname.startswith("phi.") or fullname == "dataset" or fullname.startswith("dataset.")):
            return None
        spec = PathFinder.find_spec(fullname, list(path) if path else [REFERENCE_ROOT])
        if spec is not None and spec.origin and spec.origin.endswith(".py"):
            spec.loader = _FixLoader(fullname, spec.origin)
        return spec


def evaluate_solver_module():
    """reference dataset/apps/evaluate_solver.py (+ phi/) compiled through the index-fix hook; plotting stubbed."""
    _prepare()
    builtins.__phi_idx__ = _phi_idx
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "matplotlib.animation", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "imageio", "IPython", "IPython.display"):
        sys.modules.setdefault(name, mock.MagicMock())
    if not any(isinstance(f, _FixFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _FixFinder())
    return importlib.import_module("dataset.apps.evaluate_solver")


def generate_burgers_module():
    """reference dataset/apps/generate_burgers.py (imports h5py / IPython / matplotlib at module level: stubbed)."""
    _prepare()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "matplotlib.animation", "IPython", "IPython.display",
                 "h5py"):
        sys.modules.setdefault(name, mock.MagicMock())
    return importlib.import_module("dataset.apps.generate_burgers")
