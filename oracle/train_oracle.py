"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the training step (BASELINE config 4).

``training_loss`` restates ``TouchedRegraster.training_step`` (model5_b.py:912-1155, the non-pretrain branch) on
top of ``oracle/puzzle_oracle.py`` with train-mode BatchNorm (batch statistics over the point index,
model5_b.py:424-425/:447-448); gradients come from torch autograd on the CPU, the EMD terms from the C restatement
(oracle/emd_oracle.c) wrapped exactly like ``PyTorchEMD/emd.py:5-21`` (``match`` is a constant in backward).

Pinned by ``tests/golden/reference_training.npz`` (oracle/make_golden_training.py): the UNMODIFIED reference
``training_step`` executed on the CPU with ``earth_mover_distance`` replaced by the same C-oracle Function (the
reference's EMD is CUDA-only), its loss and per-parameter gradient digests.
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch
import torch.nn.functional as F

from . import emd_oracle
from . import puzzle_oracle as po

Tensor = torch.Tensor


class EmdFunction(torch.autograd.Function):
    """PyTorchEMD/emd.py:5-21 on the CPU C oracle."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        a, b = xyz1.detach().contiguous().numpy(), xyz2.detach().contiguous().numpy()
        match = emd_oracle.approxmatch(a, b)
        cost = emd_oracle.matchcost(a, b, match)
        ctx.save_for_backward(xyz1, xyz2, torch.from_numpy(match))
        return torch.from_numpy(cost)

    @staticmethod
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = emd_oracle.matchcost_grad(grad_cost.contiguous().numpy(), xyz1.detach().numpy(),
                                           xyz2.detach().numpy(), match.numpy())
        return torch.from_numpy(g1), torch.from_numpy(g2)


def earth_mover_distance(xyz1, xyz2, transpose=True):
    """PyTorchEMD/emd.py:24-45."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1, xyz2 = xyz1.transpose(1, 2), xyz2.transpose(1, 2)
    return EmdFunction.apply(xyz1, xyz2)


def training_loss(sd: Mapping[str, Tensor], batch, starts=None, loss_mode: int = 1, loss_sum: bool = False,
                  use_emd2: bool = False, use_cd2: bool = False, use_emd3: bool = False,
                  bn_state: Dict[str, Tensor] = None, pretrain: bool = False) -> Dict[str, Tensor]:
    """model5_b.py:912-1155: returns every logged term and the total ``loss``.  ``pretrain`` selects the predict6
    branch (:928-931: both clouds through ``Encoder``, the step returns after the pose losses, :1048-1050).
    ``sd`` tensors may require grad; ``bn_state`` (optional dict) receives the updated running statistics."""
    fpc, mrpc, igt, rpc, fpcb, rpcb, fpc_idx, rpc_idx = batch[:8]
    if pretrain:
        o = predict6_train(sd, fpc, mrpc, starts, bn_state)
    else:
        o = po.predict5(sd, fpc, mrpc, need=True, starts=starts, train_bn=True, bn_state=bn_state)
    out, de_fpcb, de_mrpcb = o["out"], o.get("de_fpcb"), o.get("de_mrpcb")
    x2, attention = o["enc_fpc"]["x2"], o["enc_fpc"]["attention"]
    mrpc_x2, mrpc_attention = o["enc_mrpc"]["x2"], o["enc_mrpc"]["attention"]
    att1, att2 = attention.mean(dim=1), mrpc_attention.mean(dim=1)
    x2att1 = x2[:, torch.topk(att1, 32)[1][:, 0]]                  # :939-942 (yields [B,B,3], as in the reference)
    x2att2 = mrpc_x2[:, torch.topk(att2, 32)[1][:, 0]]
    mat = po.se3_exp(out)
    de_mrpc = po.se3_transform(mat, mrpc.permute(0, 2, 1)).permute(0, 2, 1)
    d1, d2 = po.chamfer_loss(rpc, de_mrpc)
    red = torch.sum if loss_sum else torch.mean
    loss_re = red(d1) + red(d2)
    loss_g = po.comp(mat, igt)
    a1, a2 = po.chamfer_loss(x2att1, x2att2)
    emd = earth_mover_distance(de_mrpc, rpc, transpose=False)
    loss_emd = red(emd)
    loss_cd2 = red(a1) + red(a2)
    emd2 = torch.sum(earth_mover_distance(x2att1, x2att2, transpose=False))
    loss = {0: loss_re + loss_g, 1: loss_re + loss_g + loss_emd, 2: loss_emd, 3: loss_emd + loss_g,
            4: loss_emd + loss_re, 5: loss_g, 6: loss_re}[loss_mode]
    if use_emd2:
        loss = loss + emd2
    if use_cd2:
        loss = loss + loss_cd2
    if pretrain:
        return dict(loss=loss, loss_re=loss_re, loss_g=loss_g, loss_emd=loss_emd, loss_cd2=loss_cd2, emd2=emd2,
                    out=out, de_mrpc=de_mrpc, mat=mat, fwd=o)
    ce_f = F.cross_entropy(de_fpcb, fpc_idx.squeeze().long().reshape(de_fpcb.shape[0], -1))
    ce_m = F.cross_entropy(de_mrpcb, rpc_idx.squeeze().long().reshape(de_mrpcb.shape[0], -1))
    loss = loss + ce_f + ce_m
    idx_f = torch.topk(torch.softmax(de_fpcb, dim=1)[:, 1, :], 128, 1)[1]
    idx_m = torch.topk(torch.softmax(de_mrpcb, dim=1)[:, 1, :], 128, 1)[1]
    bnd_f = torch.gather(fpc, 1, idx_f.unsqueeze(-1).repeat(1, 1, 3))
    bnd_m = torch.gather(mrpc, 1, idx_m.unsqueeze(-1).repeat(1, 1, 3))
    c1, c2 = po.chamfer_loss(bnd_f, fpcb)
    loss_fpcb = torch.mean(c1) + torch.mean(c2)
    inv_bnd_m = po.se3_transform(mat, bnd_m.permute(0, 2, 1)).permute(0, 2, 1)
    c1, c2 = po.chamfer_loss(inv_bnd_m, rpcb)
    loss_mrpcb = torch.mean(c1) + torch.mean(c2)
    emd_fpcb = torch.mean(earth_mover_distance(bnd_f, fpcb, transpose=False))
    emd_mrpcb = torch.mean(earth_mover_distance(inv_bnd_m, rpcb, transpose=False))
    loss = loss + loss_mrpcb + loss_fpcb
    if use_emd3:
        loss = loss + emd_fpcb + emd_mrpcb
    return dict(loss=loss, loss_re=loss_re, loss_g=loss_g, loss_emd=loss_emd, loss_cd2=loss_cd2, emd2=emd2,
                ce_f=ce_f, ce_m=ce_m, loss_fpcb=loss_fpcb, loss_mrpcb=loss_mrpcb, emd_fpcb=emd_fpcb,
                emd_mrpcb=emd_mrpcb, out=out, de_fpcb=de_fpcb, de_mrpcb=de_mrpcb, de_mrpc=de_mrpc, mat=mat,
                idx_f=idx_f, idx_m=idx_m, fwd=o)


def predict6_train(sd, fpc, mrpc, starts=None, bn_state=None):
    """predict6(training=True, need=True, pretrain=True), model5_b.py:612-658: ``Encoder`` on both clouds (two
    train-mode passes: its BatchNorm running statistics are updated twice), pose head only."""
    st1, st2 = (None, None) if starts is None else starts
    e1 = po.encoder_forward(sd, "Encoder", fpc, st1, True, bn_state)
    sd2 = dict(sd)
    if bn_state:       # the second pass starts from the running statistics the first one left behind
        sd2.update({k: v for k, v in bn_state.items() if k.startswith("Encoder.")})
    e2 = po.encoder_forward(sd2, "Encoder", mrpc, st2, True, bn_state)
    out6 = po._seq(sd, "tfMLP", torch.cat([e1["f_global"], e2["f_global"]], dim=-1), (0, 2, 4, 6, 8))
    return dict(out=out6, enc_fpc=e1, enc_mrpc=e2)


def grad_digest(g: Tensor) -> np.ndarray:
    """[sum, sum of squares, first 4 entries, last 4 entries] of a gradient tensor (float64) -- a compact pin."""
    f = g.detach().double().reshape(-1)
    head = torch.zeros(4, dtype=torch.float64)
    tail = torch.zeros(4, dtype=torch.float64)
    head[: min(4, f.numel())] = f[:4]
    tail[: min(4, f.numel())] = f[-4:]
    return torch.cat([f.sum()[None], (f * f).sum()[None], head, tail]).numpy()
