"""CPU: the oracle's restatement of the rows around the forward (test_step epilogue, chamfer_loss, comp, predict6,
dataset-side FPS / plane_split / get_boundary) against the frozen outputs of the UNMODIFIED reference
(tests/golden/reference_epilogue.npz, made by oracle/make_golden_epilogue.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from puzzlenet_b200.weights import synthetic_pairs
from tests.golden_inputs import FPS_SEED, dataset_inputs, epilogue_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_epilogue.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


def test_chamfer_and_comp(gold):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    ep = epilogue_inputs(2, po.se3_exp)
    c1, c2 = po.chamfer_loss(ep["fpcb"], ep["rpcb"])
    np.testing.assert_array_equal(c1.numpy(), gold["chamfer_128_d1"])
    np.testing.assert_array_equal(c2.numpy(), gold["chamfer_128_d2"])
    c1, c2 = po.chamfer_loss(fpc, mrpc)
    np.testing.assert_array_equal(c1.numpy(), gold["chamfer_1024_d1"])
    np.testing.assert_array_equal(c2.numpy(), gold["chamfer_1024_d2"])
    np.testing.assert_allclose(po.comp(po.se3_exp(ep["twist2"]), ep["igt"]).numpy(), gold["comp"], rtol=1e-6)


def test_test_step_scores(gold, state_dict):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    ep = epilogue_inputs(2, po.se3_exp)
    torch.manual_seed(FPS_SEED)
    o = po.predict5(state_dict, fpc, mrpc)
    s = po.test_step_scores(o["out"], o["de_fpcb"], o["de_mrpcb"], fpc, ep["rpc"], ep["fpcb"], ep["rpcb"],
                            ep["fpc_idx"], ep["rpc_idx"], ep["igt"])
    ref = gold["test_step"][0]          # r_mse r_mae t_mse t_mae r_iso t_iso iou_f iou_m cd_fpc cd_rpc
    got = [s["t_mse"].mean(), s["t_mae"].mean(), s["r_iso"].mean(), s["t_iso"].mean(),
           s["inter_f"].sum() / s["union_f"].sum(), s["inter_m"].sum() / s["union_m"].sum(),
           s["cd_fpc"].mean(), s["cd_rpc"].mean()]
    np.testing.assert_allclose(np.array([float(v) for v in got]), ref[2:], rtol=2e-6)


def test_predict6(gold, state_dict):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    torch.manual_seed(FPS_SEED)
    o = po.predict6(state_dict, fpc, mrpc)
    np.testing.assert_array_equal(o["out"].numpy(), gold["predict6"])


def test_dataset_side(gold):
    cloud = dataset_inputs()
    np.random.seed(21)
    up, down = po.plane_split(cloud)
    assert up.shape[0] == gold["split_up_n"] and down.shape[0] == gold["split_down_n"]
    np.testing.assert_array_equal(up[:64], gold["split_up_head"])
    np.testing.assert_array_equal(down[:64], gold["split_down_head"])
    np.random.seed(22)
    up_s = po.dataset_fps(up, 1024)
    down_s = po.dataset_fps(down, 1024)
    np.testing.assert_array_equal(up_s, gold["ds_fps_up"])
    np.testing.assert_array_equal(down_s, gold["ds_fps_down"])
    fb, rb, fi, ri = po.get_boundary(torch.from_numpy(down_s).float(), torch.from_numpy(up_s).float())
    np.testing.assert_array_equal(fb.numpy(), gold["gb_fpcb"])
    np.testing.assert_array_equal(rb.numpy(), gold["gb_rpcb"])
    np.testing.assert_array_equal(fi.numpy(), gold["gb_fpc_idx"])
    np.testing.assert_array_equal(ri.numpy(), gold["gb_rpc_idx"])


def test_dataset_fps_equals_pointnet_fps():
    """A14 == A1 given the same start: the float64 running distance of the numpy version holds only
    fp32-representable values (1e10 included), so its comparisons are the fp32 ones."""
    cloud = dataset_inputs()[:3000]
    pts = po.dataset_fps(cloud, 256, start=17)
    idx = po.farthest_point_sample(torch.from_numpy(cloud)[None], 256, start=torch.tensor([17]))
    np.testing.assert_array_equal(pts, cloud[idx[0].numpy()])
