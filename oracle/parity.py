"""TEST INFRASTRUCTURE ONLY -- parity figures of a ``predict5`` result against the CPU oracle.

Used by ``tests/`` and by ``bench.py``'s ``parity`` leg (the oracle as the CHECKER of the outputs the timed steps
produced); never on the product path.  The tolerances are north_star's: features / boundary logits within 1e-4
(fp32-tolerance paths) or 2e-2 (bf16 path), rotation within 0.01 degree, translation within 1e-4.

Two error figures per tensor:
* ``rel``      = max|got - ref| / max|ref|                      (relative to the tensor's largest magnitude)
* ``rel_elem`` = max_i |got_i - ref_i| / max(|ref_i|, rms(ref))   (element-wise, guarded: an element smaller than the
                 tensor's RMS is measured against the RMS instead of against itself, so a logit that happens to be
                 ~0 does not turn an absolute error of 1e-7 into a relative error of 1)
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import puzzle_oracle as po

BOUNDS = {
    # precision: (bound on rel, bound on rel_elem, rotation bound [deg], translation bound)
    "fp32": (1e-4, 1e-4, 0.01, 1e-4),
    "split": (1e-4, 1e-4, 0.01, 1e-4),
    # north_star bounds the bf16 path's features / logits at 2e-2 and says nothing about its pose; the pose error that
    # path reaches is held to a looser figure of its own (1.5 deg / 5e-2) so that a regression is still caught
    "bf16": (2e-2, 2e-2, 1.5, 5e-2),
}
POSE_CLAIMED = {"fp32": True, "split": True, "bf16": False}   # which paths claim north_star's 0.01 deg / 1e-4


def rel(got, ref) -> float:
    ref = torch.as_tensor(ref).double()
    got = torch.as_tensor(got).detach().cpu().double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def rel_elem(got, ref) -> float:
    ref = torch.as_tensor(ref).double()
    got = torch.as_tensor(got).detach().cpu().double()
    guard = ref.pow(2).mean().sqrt().clamp_min(1e-30)
    return ((got - ref).abs() / torch.maximum(ref.abs(), guard)).max().item()


def pose_errors(twist_got, twist_ref):
    """(max rotation error [deg], max translation error) of se3.exp(twist) -- se_math/se3.py:57-80 and
    metrics.py:54-84 (isotropic errors), evaluated with the oracle's own exp on both twists IN FLOAT64: rotation
    matrices rounded to fp32 are orthonormal only to ~1e-7, which the trace formula turns into a 0.02-0.05 degree noise
    floor -- above the 0.01 degree bound being checked.  In float64 the figure is the pose error the twist difference
    causes, nothing else."""
    g = po.se3_exp(torch.as_tensor(twist_got).detach().cpu().double())
    r = po.se3_exp(torch.as_tensor(twist_ref).double())
    return (po.rotation_error_deg(g[:, :3, :3], r[:, :3, :3]).max().item(),
            po.translation_error(g[:, :3, 3], r[:, :3, 3]).max().item())


def predict5_parity(state_dict, fpc, mrpc, starts, got_out, got_de_fpcb, got_de_mrpcb,
                    pairs: Optional[Sequence[int]] = None, ref: Optional[dict] = None) -> Dict[str, float]:
    """Runs the oracle on ``pairs`` of the batch (pairs are independent in eval mode, model5_b.py:672-759) and returns
    the parity figures of the three ``need=False`` outputs.  ``starts`` is the [4,B] FPS start table of the call."""
    idx = torch.arange(fpc.shape[0]) if pairs is None else torch.as_tensor(list(pairs))
    if ref is None:
        ref = oracle_subset(state_dict, fpc, mrpc, starts, idx)
    out, de_f, de_m = (torch.as_tensor(t).detach().cpu()[idx] for t in (got_out, got_de_fpcb, got_de_mrpcb))
    rot, trans = pose_errors(out, ref["out"])
    return {
        "pairs_checked": int(idx.numel()),
        "rel_out": rel(out, ref["out"]), "rel_elem_out": rel_elem(out, ref["out"]),
        "rel_logits": max(rel(de_f, ref["de_fpcb"]), rel(de_m, ref["de_mrpcb"])),
        "rel_elem_logits": max(rel_elem(de_f, ref["de_fpcb"]), rel_elem(de_m, ref["de_mrpcb"])),
        "rot_deg": rot, "trans": trans,
    }


def oracle_subset(state_dict, fpc, mrpc, starts, idx) -> dict:
    fpc, mrpc, starts = fpc.cpu().float(), mrpc.cpu().float(), starts.cpu().long()
    st = ((starts[0][idx], starts[1][idx]), (starts[2][idx], starts[3][idx]))
    with torch.no_grad():
        return po.predict5(state_dict, fpc[idx], mrpc[idx], starts=st)


def within(p: Dict[str, float], precision: str) -> bool:
    feat, feat_elem, rot, trans = BOUNDS[precision]
    return bool(max(p["rel_out"], p["rel_logits"]) < feat and max(p["rel_elem_out"], p["rel_elem_logits"]) < feat_elem
                and p["rot_deg"] < rot and p["trans"] < trans)
