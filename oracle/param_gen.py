"""Deterministic synthetic weights from a state_dict key/shape inventory — TEST INFRASTRUCTURE ONLY.
Every tensor is drawn from its own generator seeded by crc32(key) ^ seed, so the same weights can be rebuilt on any box
from the key names alone (no checkpoint is reachable)."""
import math
import zlib
from collections import OrderedDict

import torch


def make_params(shapes, seed: int = 0):
    out = OrderedDict()
    for key, shape in shapes.items():
        shape = tuple(shape)
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if key.endswith("norm.weight") or key.endswith(".g") or key.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("norm.bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shapes[key[:-5] + ".weight"] if key.endswith(".bias") and (key[:-5] + ".weight") in shapes else shape
            fan_in = 1
            for s in tuple(wshape)[1:]:
                fan_in *= s
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
        out[key] = t.float().contiguous()
    return out
