"""BASELINE config 5: multi-piece greedy assembly -- all-pairs matching of 32 DublinCity-like 11000-point pieces
(496 candidate pairs) sharded over the GPUs of one box.

    python scripts/bench_assembly.py [--pieces 32] [--points 11000] [--iters 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_assembly.py

Pieces are synthetic (the union of 3-6 random planar rectangles per piece, seed 5+piece); weights
are synthetic_state_dict(0).  One JSON line from rank 0: time per full assembly (FPS down-sampling of the pieces,
pair scoring = predict5 + pz_pair_score, the all_gather of [n_pairs, 8] rows, host-side greedy merge)."""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dublin_like_piece(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(5 + seed)
    k = int(rng.integers(3, 7))
    per = np.full(k, n // k)
    per[: n - per.sum()] += 1
    parts = []
    for c in per:
        origin = rng.uniform(-0.4, 0.4, 3)
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        v = rng.normal(size=3); v -= u * (u @ v); v /= np.linalg.norm(v)
        ext = rng.uniform(0.2, 0.6, 2)
        ab = rng.uniform(0, 1, (c, 2)) * ext
        parts.append(origin + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, 0.002, (c, 3)))
    pts = np.concatenate(parts).astype(np.float32)
    pts -= pts.mean(0)
    return pts / np.abs(pts).max() * 0.5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pieces", type=int, default=32)
    ap.add_argument("--points", type=int, default=11000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--rescore", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from puzzlenet_b200 import assembly
    from puzzlenet_b200.model5_b import TouchedRegraster
    from puzzlenet_b200.weights import synthetic_state_dict
    model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
    model.load_state_dict(synthetic_state_dict(0))
    model.to(dev).eval()
    model.precision = a.precision
    scorer = assembly.ModelScorer(model)
    pieces = [dublin_like_piece(i, a.points) for i in range(a.pieces)]
    starts = [int(np.random.default_rng(100 + i).integers(0, a.points)) for i in range(a.pieces)]
    n_pairs = a.pieces * (a.pieces - 1) // 2

    def once():
        torch.manual_seed(1234)
        clouds = assembly.downsample_pieces(pieces, 1024, starts=starts, device=dev)
        if a.rescore:
            return assembly.assemble(clouds, scorer, batch=64, rescore=True)
        pairs, rows = assembly.score_all_pairs(clouds, scorer, batch=64)
        return assembly.greedy_assemble(a.pieces, pairs, rows)[::2]

    for _ in range(a.warmup):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        poses, merges = once()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = dt.item() / a.iters * 1e3
        print(json.dumps({"metric": "assemblies/sec (config 5)", "pieces": a.pieces, "points_per_piece": a.points,
                          "pairs": n_pairs, "n_gpus": world, "ms_per_assembly": ms,
                          "pairs_per_s": n_pairs / ms * 1e3, "merges": len(merges), "precision": a.precision,
                          "rescore": a.rescore, "timing": "host wall clock around whole assemblies (includes "
                          "H2D of the raw pieces, the NCCL all_gathers and the host-side greedy merge), max over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
