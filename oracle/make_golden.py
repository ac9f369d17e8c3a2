"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the UNMODIFIED reference as fixtures.

Runs in the build container only (needs /root/reference).  Executes the reference's own torch
code on the CPU through oracle/ref_shim.py and writes tests/golden/reference_goldens.npz.
Inputs and weights are NOT stored: they are regenerated deterministically by
puzzlenet_b200.weights (synthetic_state_dict(0), synthetic_pairs(2, seed=64)) and the seeds below.

    python oracle/make_golden.py
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

from tests.golden_inputs import FPS_SEED, golden_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_goldens.npz")


def main():
    ns = ref_shim.load_reference()
    pu, m5, se3 = ns.pointnet_util, ns.model5_b, ns.se3
    out = {}
    xyz, feat, big, twist = golden_inputs()

    # ---- pointnet_util
    torch.manual_seed(11)
    out["fps_small"] = pu.farthest_point_sample(xyz, 64).numpy().astype(np.int32)
    torch.manual_seed(12)
    out["fps_11000_1024"] = pu.farthest_point_sample(big, 1024).numpy().astype(np.int32)
    q = xyz[:, :40]
    out["sqdist"] = pu.square_distance(q, xyz).numpy()
    out["ball"] = pu.query_ball_point(0.25, 12, xyz, q).numpy().astype(np.int32)
    idx = torch.randint(0, 300, (2, 9, 5), generator=torch.Generator().manual_seed(13))
    out["index_points"] = pu.index_points(feat, idx).numpy()
    torch.manual_seed(14)
    nx, npts, gx, fi = pu.sample_and_group(32, 0, 16, xyz, feat, True, True)
    out.update(sg_new_xyz=nx.numpy(), sg_new_points=npts.numpy(), sg_grouped_xyz=gx.numpy(),
               sg_fps_idx=fi.numpy().astype(np.int32))
    torch.manual_seed(15)
    nx, npts = pu.sample_and_group(32, 0.3, 16, xyz, feat, False, False)
    out.update(sgb_new_xyz=nx.numpy(), sgb_new_points=npts.numpy())

    # ---- se3
    out["se3_exp"] = se3.exp(twist).numpy()

    # ---- model5_b: encoder + predict5, B=2
    sd = synthetic_state_dict(0)
    model = m5.TouchedRegraster(ref_shim.reference_config())
    model.load_state_dict(sd, strict=True)
    model.eval()
    fpc, mrpc = synthetic_pairs(2, seed=64)
    with torch.no_grad():
        torch.manual_seed(FPS_SEED)
        r = model.predict5(make_batch(fpc, mrpc), 0, need=True)
        out.update(p5_out=r[0].numpy(), p5_x2_fpc=r[2].numpy(), p5_attn_fpc_rows=r[3][:, ::16].numpy(),
                   p5_x2_mrpc=r[4].numpy(), p5_attn_mrpc_rows=r[5][:, ::16].numpy(),
                   p5_de_fpcb=r[6].numpy(), p5_de_mrpcb=r[7].numpy())
        out["p5_mat"] = se3.exp(r[0]).numpy()
        torch.manual_seed(FPS_SEED)
        e = model.Encoder(fpc)
        out.update(enc_f_global=e[0].numpy(), enc_x2=e[1].numpy(), enc_out_rows=e[3][:, ::32].numpy(),
                   enc_x_feature_rows=e[4][:, ::16].numpy())
        # indices the encoder used (same seed -> same two draws)
        torch.manual_seed(FPS_SEED)
        fps1 = pu.farthest_point_sample(fpc, 512)
        x1 = pu.index_points(fpc, fps1)
        d1 = pu.square_distance(x1, fpc)
        knn1 = torch.sort(d1, dim=-1, stable=True).indices[:, :, :32]
        fps2 = pu.farthest_point_sample(x1, 256)
        x2 = pu.index_points(x1, fps2)
        knn2 = torch.sort(pu.square_distance(x2, x1), dim=-1, stable=True).indices[:, :, :32]
        assert torch.equal(x2, e[1])
        # the reference's own (unstable) argsort must select the same neighbour SETS
        ref_knn1 = d1.argsort()[:, :, :32]
        assert torch.equal(ref_knn1.sort(-1).values, knn1.sort(-1).values)
        out.update(enc_fps1=fps1.numpy().astype(np.int16), enc_knn1=knn1.numpy().astype(np.int16),
                   enc_fps2=fps2.numpy().astype(np.int16), enc_knn2=knn2.numpy().astype(np.int16))
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, f"{os.path.getsize(GOLDEN) / 1024:.0f} KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
