"""CPU: the oracle restatement must reproduce the frozen outputs of the UNMODIFIED reference
(tests/golden/reference_goldens.npz, made by oracle/make_golden.py from /root/reference)."""
import numpy as np
import torch

from oracle import puzzle_oracle as po
from puzzlenet_b200.weights import synthetic_pairs
from tests.golden_inputs import FPS_SEED, golden_inputs


def _eq(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_pointnet_util_goldens(goldens):
    xyz, feat, big, _ = golden_inputs()
    torch.manual_seed(11)
    _eq(po.farthest_point_sample(xyz, 64).numpy(), goldens["fps_small"])
    q = xyz[:, :40]
    _eq(po.square_distance(q, xyz).numpy(), goldens["sqdist"])
    _eq(po.query_ball_point(0.25, 12, xyz, q).numpy(), goldens["ball"])
    idx = torch.randint(0, 300, (2, 9, 5), generator=torch.Generator().manual_seed(13))
    _eq(po.index_points(feat, idx).numpy(), goldens["index_points"])


def test_fps_11000_to_1024_golden(goldens):
    _, _, big, _ = golden_inputs()
    torch.manual_seed(12)
    _eq(po.farthest_point_sample(big, 1024).numpy(), goldens["fps_11000_1024"])


def test_sample_and_group_goldens(goldens):
    xyz, feat, _, _ = golden_inputs()
    torch.manual_seed(14)
    nx, npts, gx, fi = po.sample_and_group(32, 0, 16, xyz, feat, returnfps=True, knn=True)
    _eq(nx.numpy(), goldens["sg_new_xyz"])
    _eq(fi.numpy(), goldens["sg_fps_idx"])
    # the duplicated point (xyz[1,17] == xyz[1,3]) makes the reference's unstable argsort order
    # ambiguous between the twins; compare each group as a multiset of rows
    a = np.sort(npts.numpy().reshape(2, 32, 16, -1), axis=2)
    b = np.sort(goldens["sg_new_points"].reshape(2, 32, 16, -1), axis=2)
    assert np.array_equal(a[..., :3], b[..., :3])
    _eq(npts.numpy()[0], goldens["sg_new_points"][0])      # cloud 0 has no duplicates: exact incl. order
    _eq(gx.numpy()[0], goldens["sg_grouped_xyz"][0])
    torch.manual_seed(15)
    nx, npts = po.sample_and_group(32, 0.3, 16, xyz, feat, knn=False)
    _eq(nx.numpy(), goldens["sgb_new_xyz"])
    _eq(npts.numpy(), goldens["sgb_new_points"])


def test_se3_exp_golden(goldens):
    _, _, _, twist = golden_inputs()
    np.testing.assert_allclose(po.se3_exp(twist).numpy(), goldens["se3_exp"], rtol=0, atol=1e-6)


def test_predict5_goldens(goldens, state_dict):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    torch.manual_seed(FPS_SEED)
    o = po.predict5(state_dict, fpc, mrpc)
    e1, e2 = o["enc_fpc"], o["enc_mrpc"]
    _eq(e1["fps1"].numpy(), goldens["enc_fps1"])
    _eq(e1["knn1"].numpy(), goldens["enc_knn1"])
    _eq(e1["fps2"].numpy(), goldens["enc_fps2"])
    _eq(e1["knn2"].numpy(), goldens["enc_knn2"])
    _eq(e1["x2"].numpy(), goldens["p5_x2_fpc"])
    _eq(e2["x2"].numpy(), goldens["p5_x2_mrpc"])
    # same ATen ops in the same order -> bit-identical on the same machine; keep a hair of slack for
    # a different CPU's GEMM blocking on the GPU box's host
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o["out"].numpy(), goldens["p5_out"], **tol)
    np.testing.assert_allclose(o["de_fpcb"].numpy(), goldens["p5_de_fpcb"], **tol)
    np.testing.assert_allclose(o["de_mrpcb"].numpy(), goldens["p5_de_mrpcb"], **tol)
    np.testing.assert_allclose(e1["attention"][:, ::16].numpy(), goldens["p5_attn_fpc_rows"], **tol)
    np.testing.assert_allclose(e2["attention"][:, ::16].numpy(), goldens["p5_attn_mrpc_rows"], **tol)
    np.testing.assert_allclose(e1["f_global"].numpy(), goldens["enc_f_global"], **tol)
    np.testing.assert_allclose(e1["out"][:, ::32].numpy(), goldens["enc_out_rows"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(e1["x_feature"][:, ::16].numpy(), goldens["enc_x_feature_rows"], **tol)
    np.testing.assert_allclose(po.se3_exp(o["out"]).numpy(), goldens["p5_mat"], rtol=1e-5, atol=1e-6)
