"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the reference's (unused) PointNet++ blocks, pointnet_util.py:159-315,
run unmodified on the CPU in eval mode with seeded parameters and randomised BatchNorm statistics.
    python oracle/make_golden_pointnet.py -> tests/golden/reference_pointnet.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests.golden_inputs import pointnet_block_inputs, seed_block  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_pointnet.npz")


def main():
    pu = ref_shim.load_reference().pointnet_util
    xyz, feat = pointnet_block_inputs()
    out = {}
    with torch.no_grad():
        sa = seed_block(pu.PointNetSetAbstraction(32, 0.3, 16, 8 + 3, [16, 32], False, knn=True), 1).eval()
        torch.manual_seed(21)
        nx, f = sa(xyz, feat)
        out.update(sa_xyz=nx.numpy(), sa_feat=f.numpy())
        sa_ball = seed_block(pu.PointNetSetAbstraction(32, 0.3, 16, 8 + 3, [16, 32], False, knn=False), 2).eval()
        torch.manual_seed(22)
        nx, f = sa_ball(xyz, feat)
        out.update(sab_xyz=nx.numpy(), sab_feat=f.numpy())
        sa_all = seed_block(pu.PointNetSetAbstraction(None, None, None, 8 + 3, [16, 24], True), 3).eval()
        nx, f = sa_all(xyz, feat)
        out.update(saa_xyz=nx.numpy(), saa_feat=f.numpy())
        msg = seed_block(pu.PointNetSetAbstractionMsg(32, [0.2, 0.4], [8, 16], 8, [[16, 16], [16, 32]], knn=True), 4).eval()
        torch.manual_seed(23)
        nx, f = msg(xyz, feat)
        out.update(msg_xyz=nx.numpy(), msg_feat=f.numpy())
        fp = seed_block(pu.PointNetFeaturePropagation(8 + 32, [32, 16]), 5).eval()
        o = fp(xyz.permute(0, 2, 1), torch.from_numpy(out["sa_xyz"]).permute(0, 2, 1),
               feat.permute(0, 2, 1), torch.from_numpy(out["sa_feat"]).permute(0, 2, 1))
        out["fp_out"] = o.numpy()
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, len(out), "arrays")


if __name__ == "__main__":
    main()
