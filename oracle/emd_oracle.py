"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/emd_oracle.c (numpy in / numpy out)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libemd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "emd_oracle.c")
    if force or not os.path.isfile(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, src, "-lm"])
    return SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def approxmatch(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.zeros((b, m, n), np.float32)
    _load().emd_oracle_approxmatch(b, n, m, _p(xyz1), _p(xyz2), _p(match))
    return match


def matchcost(xyz1, xyz2, match):
    xyz1, xyz2, match = _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = np.zeros((b,), np.float32)
    _load().emd_oracle_matchcost(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(cost))
    return cost


def matchcost_grad(grad_cost, xyz1, xyz2, match):
    gc, xyz1, xyz2, match = _f(grad_cost), _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, 3), np.float32)
    g2 = np.zeros((b, m, 3), np.float32)
    _load().emd_oracle_matchcost_grad(b, n, m, _p(gc), _p(xyz1), _p(xyz2), _p(match), _p(g1), _p(g2))
    return g1, g2


def earth_mover_distance(xyz1, xyz2):
    """cost (b,) = matchcost(approxmatch) as PyTorchEMD/emd.py:11-12 composes them."""
    return matchcost(xyz1, xyz2, approxmatch(xyz1, xyz2))
