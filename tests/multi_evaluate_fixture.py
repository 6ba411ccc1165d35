"""Seeded inputs of the multi_evaluate parity case, shared by tests/golden/make_golden_multi_evaluate.py (which feeds them to the
UNMODIFIED reference method) and tests/test_evaluate_gpu.py."""
import torch


def inputs(B, seed=21, F=32, S=64, T=256):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    blob = torch.exp(-((xx - 0.1) ** 2 + (yy + 0.3) ** 2) / 0.05)
    pred = torch.randn(B, F, 6, S, S, generator=g) * 0.5
    pred[:, :, 0] = pred[:, :, 0].abs()
    pred[:, :, 5] = torch.rand(B, F, 1, 1, generator=g).expand(B, F, S, S)
    data = torch.zeros(B, T, 6, S, S)
    data[:, 0, 0] = blob[None] * torch.linspace(1.0, 0.7, B)[:, None, None]
    return pred, data
