"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the UNMODIFIED reference for the rows around the forward:
test_step (model5_b.py:1279-1358), chamfer_loss / comp (:1495-1519), predict6 (:612-668) and the dataset-side
numpy FPS / get_boundary / plane_split (dataset.py:1147-1163, :1357-1367, :761-775).

Build container only (needs /root/reference).  Writes tests/golden/reference_epilogue.npz; inputs are regenerated
from seeds by tests/golden_inputs.py::epilogue_inputs.

    python oracle/make_golden_epilogue.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from puzzlenet_b200.weights import make_batch, synthetic_pairs, synthetic_state_dict  # noqa: E402
from tests.golden_inputs import FPS_SEED, dataset_inputs, epilogue_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_epilogue.npz")


def main():
    ns = ref_shim.load_reference()
    m5, se3 = ns.model5_b, ns.se3
    import dataset as ref_dataset   # the reference's dataset.py (on sys.path through the shim)
    out = {}
    sd = synthetic_state_dict(0)
    model = m5.TouchedRegraster(ref_shim.reference_config())
    model.load_state_dict(sd, strict=True)
    model.eval()
    model.device = torch.device("cpu")
    fpc, mrpc = synthetic_pairs(2, seed=64)
    ep = epilogue_inputs(2, se3.exp)
    batch = [fpc, mrpc, ep["igt"], ep["rpc"], ep["fpcb"], ep["rpcb"], ep["fpc_idx"], ep["rpc_idx"]]
    with torch.no_grad():
        torch.manual_seed(FPS_SEED)
        out["test_step"] = model.test_step(batch, 0).numpy()
        a, b = ep["fpcb"], ep["rpcb"]
        c1, c2 = model.chamfer_loss(a, b)
        out.update(chamfer_128_d1=c1.numpy(), chamfer_128_d2=c2.numpy())
        c1, c2 = model.chamfer_loss(fpc, mrpc)
        out.update(chamfer_1024_d1=c1.numpy(), chamfer_1024_d2=c2.numpy())
        out["comp"] = model.comp(se3.exp(ep["twist2"]), ep["igt"]).numpy()
        torch.manual_seed(FPS_SEED)
        out["predict6"] = model.predict6(make_batch(fpc, mrpc), 0, need=False, training=False, pretrain=True).numpy()
    # ---- dataset side: the methods only use self for self.chamfer_loss
    ds = ref_dataset.CADDataset.__new__(ref_dataset.CADDataset)
    cloud = dataset_inputs()
    np.random.seed(21)
    up, down = ref_dataset.plane_split(cloud)
    out.update(split_up_n=np.int64(up.shape[0]), split_down_n=np.int64(down.shape[0]),
               split_up_head=up[:64].copy(), split_down_head=down[:64].copy())
    np.random.seed(22)
    up_s = ds.fps(up, 1024)
    down_s = ds.fps(down, 1024)
    out.update(ds_fps_up=up_s.copy(), ds_fps_down=down_s.copy())
    fb, rb, fi, ri = ds.get_boundary(torch.from_numpy(down_s).float(), torch.from_numpy(up_s).float())
    out.update(gb_fpcb=fb.numpy(), gb_rpcb=rb.numpy(), gb_fpc_idx=fi.numpy(), gb_rpc_idx=ri.numpy())
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, f"{os.path.getsize(GOLDEN) / 1024:.0f} KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
