"""Drop-in for the reference's pybind11 module ``emd_cuda`` (PyTorchEMD/cuda/emd.cpp:23-27).

Same three functions, same argument order and output shapes, same error behaviour for bad
shapes (an exception).  Backed by the sm_100a kernels in ``csrc/emd.cu`` through the C ABI.
"""
from __future__ import annotations

import torch

from . import _lib


def _check(xyz1, xyz2):
    _lib.require_cuda(xyz1, xyz2)
    if xyz1.dtype != torch.float32 or xyz2.dtype != torch.float32:
        raise TypeError("emd_cuda: float32 inputs only")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
        raise RuntimeError(f"emd_cuda: expected (b,n,3) and (b,m,3), got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
    if not (xyz1.is_contiguous() and xyz2.is_contiguous()):      # CHECK_CONTIGUOUS, emd_kernel.cu:17
        raise RuntimeError("emd_cuda: inputs must be contiguous")
    return xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]


def approxmatch_forward(xyz1, xyz2):
    """ApproxMatchForward (emd_kernel.cu:171-193): xyz1 (b,n,3), xyz2 (b,m,3) -> match (b,m,n)."""
    b, n, m = _check(xyz1, xyz2)
    match = torch.empty(b, m, n, device=xyz1.device, dtype=torch.float32)
    lib = _lib.load()
    ws_bytes = lib.pz_emd_workspace_bytes(b, n, m)
    ws = torch.empty(max(ws_bytes, 1), device=xyz1.device, dtype=torch.uint8)
    with torch.cuda.device(xyz1.device):
        _lib.call("pz_emd_approxmatch", xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, match.data_ptr(), ws.data_ptr(),
                  ws_bytes, _lib.stream_ptr())
    return match


def matchcost_forward(xyz1, xyz2, match):
    """MatchCostForward (emd_kernel.cu:257-279) -> cost (b)."""
    b, n, m = _check(xyz1, xyz2)
    _lib.require_cuda(match)
    match = match.contiguous()
    cost = torch.empty(b, device=xyz1.device, dtype=torch.float32)
    with torch.cuda.device(xyz1.device):
        _lib.call("pz_emd_matchcost", xyz1.data_ptr(), xyz2.data_ptr(), match.data_ptr(), b, n, m, cost.data_ptr(),
                  _lib.stream_ptr())
    return cost


def matchcost_backward(grad_cost, xyz1, xyz2, match):
    """MatchCostBackward (emd_kernel.cu:373-398) -> [grad1 (b,n,3), grad2 (b,m,3)]."""
    b, n, m = _check(xyz1, xyz2)
    _lib.require_cuda(grad_cost, match)
    grad_cost = grad_cost.contiguous().float()
    match = match.contiguous()
    g1 = torch.empty(b, n, 3, device=xyz1.device, dtype=torch.float32)
    g2 = torch.empty(b, m, 3, device=xyz1.device, dtype=torch.float32)
    with torch.cuda.device(xyz1.device):
        _lib.call("pz_emd_matchcost_grad", grad_cost.data_ptr(), xyz1.data_ptr(), xyz2.data_ptr(), match.data_ptr(),
                  b, n, m, g1.data_ptr(), g2.data_ptr(), _lib.stream_ptr())
    return [g1, g2]
