"""GPU parity: the PointNet++ blocks of the reference's pointnet_util (dead code on the live path, mirrored for API
completeness) vs the frozen outputs of the unmodified reference classes (tests/golden/reference_pointnet.npz)."""
import os

import numpy as np
import pytest
import torch

from tests.golden_inputs import pointnet_block_inputs, seed_block

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_pointnet.npz")


def _close(got, ref, tol=1e-4):
    ref = np.asarray(ref)
    assert got.shape == ref.shape
    assert np.abs(got.cpu().numpy() - ref).max() <= tol * max(np.abs(ref).max(), 1e-6)


def test_pointnet_blocks_match_reference():
    from puzzlenet_b200 import pointnet_util as pu
    gold = dict(np.load(GOLDEN))
    xyz, feat = pointnet_block_inputs()
    xyz, feat = xyz.to(DEV), feat.to(DEV)
    sa = seed_block(pu.PointNetSetAbstraction(32, 0.3, 16, 8 + 3, [16, 32], False, knn=True), 1).to(DEV).eval()
    torch.manual_seed(21)
    nx, f = sa(xyz, feat)
    np.testing.assert_array_equal(nx.cpu().numpy(), gold["sa_xyz"])
    _close(f, gold["sa_feat"])
    sab = seed_block(pu.PointNetSetAbstraction(32, 0.3, 16, 8 + 3, [16, 32], False, knn=False), 2).to(DEV).eval()
    torch.manual_seed(22)
    nx, f = sab(xyz, feat)
    np.testing.assert_array_equal(nx.cpu().numpy(), gold["sab_xyz"])
    _close(f, gold["sab_feat"])
    saa = seed_block(pu.PointNetSetAbstraction(None, None, None, 8 + 3, [16, 24], True), 3).to(DEV).eval()
    nx, f = saa(xyz, feat)
    np.testing.assert_array_equal(nx.cpu().numpy(), gold["saa_xyz"])
    _close(f, gold["saa_feat"])
    msg = seed_block(pu.PointNetSetAbstractionMsg(32, [0.2, 0.4], [8, 16], 8, [[16, 16], [16, 32]], knn=True), 4).to(DEV).eval()
    torch.manual_seed(23)
    nx, f = msg(xyz, feat)
    np.testing.assert_array_equal(nx.cpu().numpy(), gold["msg_xyz"])
    _close(f, gold["msg_feat"])
    fp = seed_block(pu.PointNetFeaturePropagation(8 + 32, [32, 16]), 5).to(DEV).eval()
    o = fp(xyz.permute(0, 2, 1), torch.from_numpy(gold["sa_xyz"]).to(DEV).permute(0, 2, 1), feat.permute(0, 2, 1),
           torch.from_numpy(gold["sa_feat"]).to(DEV).permute(0, 2, 1))
    _close(o, gold["fp_out"])
    with pytest.raises(NotImplementedError):
        sa.train()(xyz, feat)
