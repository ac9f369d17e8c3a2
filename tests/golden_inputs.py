"""Deterministic inputs shared by oracle/make_golden.py and the tests (kept in sync by import)."""
import torch

FPS_SEED = 1234


def golden_inputs():
    g = torch.Generator().manual_seed(7)
    xyz = torch.rand(2, 300, 3, generator=g) - 0.5
    feat = torch.randn(2, 300, 8, generator=g)
    xyz[1, 17] = xyz[1, 3]
    big = torch.rand(1, 11000, 3, generator=torch.Generator().manual_seed(3)) - 0.5
    twist = torch.randn(8, 6, generator=g) * 0.7
    twist[0, :3] *= 1e-3
    twist[1, :3] = 0
    return xyz, feat, big, twist
