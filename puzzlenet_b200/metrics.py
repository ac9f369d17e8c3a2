"""Pose error metrics of the reference's ``metrics.py`` (host-side: the reference evaluates them with numpy/scipy on
the CPU too, ``metrics.py:12-51``; the isotropic errors are also produced on the GPU by ``pz_pair_score``)."""
from __future__ import annotations

import math

import numpy as np
import torch


def inv_R_t(R, t):
    """metrics.py:7-10."""
    inv_R = R.permute(0, 2, 1).contiguous()
    inv_t = -inv_R @ t[..., None]
    return inv_R, torch.squeeze(inv_t, -1)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def anisotropic_R_error(r1, r2, seq="xyz", degrees=True):
    """metrics.py:12-34 -- Euler-angle mse / mae per sample (scipy's Rotation, as in the reference)."""
    from scipy.spatial.transform import Rotation
    r1, r2 = _np(r1), _np(r2)
    assert r1.shape == r2.shape
    e1 = np.stack([Rotation.from_matrix(r).as_euler(seq=seq, degrees=degrees) for r in r1], axis=0)
    e2 = np.stack([Rotation.from_matrix(r).as_euler(seq=seq, degrees=degrees) for r in r2], axis=0)
    return np.mean((e1 - e2) ** 2, axis=-1), np.mean(np.abs(e1 - e2), axis=-1)


def anisotropic_t_error(t1, t2):
    """metrics.py:37-51."""
    t1, t2 = _np(t1), _np(t2)
    assert t1.shape == t2.shape
    return np.mean((t1 - t2) ** 2, axis=1), np.mean(np.abs(t1 - t2), axis=1)


def isotropic_R_error(r1, r2):
    """metrics.py:54-71 (degrees)."""
    r1r2 = torch.matmul(r2.permute(0, 2, 1).contiguous(), r1)
    tr = r1r2[:, 0, 0] + r1r2[:, 1, 1] + r1r2[:, 2, 2]
    return torch.acos(torch.clamp((tr - 1) / 2, -1, 1)) / math.pi * 180


def isotropic_t_error(t1, t2, R2):
    """metrics.py:74-84."""
    R2, t2 = inv_R_t(R2, t2)
    return torch.norm(torch.squeeze(R2 @ t1[..., None], -1) + t2, dim=-1)


def compute_metrics(R, t, igt):
    """model5_b.py:1426-1440."""
    gtR, gtt = igt[:, :3, :3], igt[:, :3, 3]
    inv_R, inv_t = inv_R_t(gtR, gtt)
    r_mse, r_mae = anisotropic_R_error(R, inv_R)
    t_mse, t_mae = anisotropic_t_error(t, inv_t)
    return r_mse, r_mae, t_mse, t_mae, isotropic_R_error(R, inv_R), isotropic_t_error(t, inv_t, inv_R)
