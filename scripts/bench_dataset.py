"""Dataset-side preprocessing throughput: samples/s of the batched GPU pipeline
(puzzlenet_b200.dataset.make_pair_batch: plane cut, FPS of both halves to 1024, boundaries, rigid motion) against the
reference's per-sample numpy/torch-CPU pipeline (oracle port, one host thread -- the reference runs it on DataLoader
workers).  Pieces: 11 000-point synthetic clouds.

    python scripts/bench_dataset.py [--pieces 64] [--iters 5]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pieces", type=int, default=64)
    ap.add_argument("--points", type=int, default=11000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-samples", type=int, default=4)
    a = ap.parse_args()
    from puzzlenet_b200 import dataset as D
    g = torch.Generator().manual_seed(1)
    pieces = [(torch.randn(a.points, 3, generator=g) * 0.3).numpy() for _ in range(a.pieces)]
    np.random.seed(0); torch.manual_seed(0)
    D.make_pair_batch(pieces)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        out = D.make_pair_batch(pieces)
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / a.iters
    t0 = time.perf_counter()
    rt = D.RandomTransformSE3(0.8)
    for p in pieces[:8]:
        D.make_pair(p, rt)
    torch.cuda.synchronize()
    single_s = (time.perf_counter() - t0) / 8
    # CPU reference port, one thread
    from oracle import puzzle_oracle as po
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    for p in pieces[:a.cpu_samples]:
        u, d = po.plane_split(p)
        while u.shape[0] < 1024 or d.shape[0] < 1024:
            u, d = po.plane_split(p)
        us, ds = po.dataset_fps(u, 1024), po.dataset_fps(d, 1024)
        po.get_boundary(torch.from_numpy(ds), torch.from_numpy(us))
    cpu_s = (time.perf_counter() - t0) / a.cpu_samples
    print(json.dumps({"metric": "dataset samples/s (plane cut + 2x FPS ->1024 + boundaries + SE3)", "pieces": a.pieces,
                      "points_per_piece": a.points, "gpu_batched_samples_per_s": a.pieces / gpu_s,
                      "gpu_batched_ms_per_batch": gpu_s * 1e3, "gpu_per_sample_api_samples_per_s": 1 / single_s,
                      "cpu_port_samples_per_s_one_thread": 1 / cpu_s, "cpu_sample": f"{a.cpu_samples} samples, numpy FPS loop",
                      "timing": "host wall clock incl. H2D of the raw pieces"}))


if __name__ == "__main__":
    main()
