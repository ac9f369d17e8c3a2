"""BASELINE config 3 microbench: sample_and_group at the dataset shape (FPS 11000 -> 1024, kNN k=32,
gather + MLP 67->128->128 + max-pool), B clouds on one GPU.  Prints one JSON line.

    python scripts/bench_c3.py [--batch 64] [--iters 10]

Algorithmic (compulsory) bytes per cloud follow SURVEY.md §8(d); FPS / kNN are latency / issue bound, so their
fraction of the HBM peak is reported truthfully (it is tiny by construction)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from puzzlenet_b200 import pointnet_util as pu
from puzzlenet_b200.weights import synthetic_state_dict


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    B, N, S, K, D, C1, C2 = a.batch, 11000, 1024, 32, 64, 128, 128
    dev = "cuda:0"
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    xyz = (torch.rand(B, N, 3, generator=torch.Generator().manual_seed(3)) - 0.5).to(dev)
    feat = torch.randn(B, N, D, generator=torch.Generator().manual_seed(4)).to(dev)
    sd = synthetic_state_dict(0)
    w1, b1 = sd["Encoder.mlp3.weight"].to(dev), sd["Encoder.mlp3.bias"].to(dev)
    w2, b2 = sd["Encoder.mlp4.weight"].to(dev), sd["Encoder.mlp4.bias"].to(dev)
    torch.manual_seed(5)
    t_fps, fps_idx = timed(lambda: pu.farthest_point_sample(xyz, S), a.iters)
    new_xyz = pu.index_points(xyz, fps_idx)
    t_knn, idx = timed(lambda: pu.knn_point(K, xyz, new_xyz), a.iters)
    res = {"config": "C3 sample_and_group microbench", "B": B, "N": N, "S": S, "K": K, "hbm_peak_gbs": hbm}
    stages = {
        "fps": (t_fps, B * (12 * N + 8 * S)),
        "knn": (t_knn, B * (12 * N + 12 * S + 8 * S * K)),
    }
    for prec in ("fp32", "bf16"):
        try:
            t, out = timed(lambda: pu.group_mlp_maxpool(xyz, feat, new_xyz, idx, w1, b1, w2, b2,
                                                        precision=0 if prec == "fp32" else 1), a.iters)
            stages[f"group_mlp_maxpool_{prec}"] = (t, B * (12 * N + 4 * N * D + 8 * S + 8 * S * K + 12 * S + 4 * S * C2))
        except RuntimeError as e:          # precision not available for this entry point
            res[f"group_mlp_maxpool_{prec}"] = f"unavailable: {e}"
    t_sg, _ = timed(lambda: pu.sample_and_group(S, 0, K, xyz, feat, False, True), max(2, a.iters // 3))
    stages["sample_and_group_materialising"] = (t_sg, B * (12 * N + 4 * N * D + 12 * S + 4 * S * K * (3 + D)))
    for k, (ms, byts) in stages.items():
        gbs = byts / (ms / 1e3) / 1e9
        res[k] = {"ms": round(ms, 4), "algorithmic_MB": round(byts / 1e6, 2), "GBps": round(gbs, 1), "frac_hbm_peak": round(gbs / hbm, 5)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
