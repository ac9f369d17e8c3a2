"""BASELINE config 4: PuzzleNet training step (pose + boundary losses + EMD) data-parallel, 64 pairs per GPU
(batch 512 on 8 GPUs), one NCCL all-reduce of the flat gradient buffer per step.

    python scripts/bench_train.py [--pairs 64] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_train.py

Synthetic batch (tests/golden_inputs-style: mrpc = igt . rpc), synthetic_state_dict(0) weights, loss_mode 1.  One JSON
line from rank 0; timing = CUDA events around whole steps (forward, losses, backward, all-reduce, Adam), max over
ranks.  --phases additionally times forward / backward / all-reduce+Adam separately (with a sync between them)."""
import argparse
import json
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_training_batch(B, seed, dev):
    from puzzlenet_b200 import se3
    g = torch.Generator().manual_seed(seed)
    fpc = torch.rand(B, 1024, 3, generator=g) - 0.5
    rpc = torch.rand(B, 1024, 3, generator=g) - 0.5
    twist = (torch.randn(B, 6, generator=g) * 0.3).to(dev)
    igt = se3.exp(twist)
    rpc_d = rpc.to(dev)
    mrpc = (igt[:, :3, :3] @ rpc_d.permute(0, 2, 1) + igt[:, :3, 3:]).permute(0, 2, 1).contiguous()
    fpcb = torch.rand(B, 128, 3, generator=g) - 0.5
    rpcb = torch.rand(B, 128, 3, generator=g) - 0.5
    fi = (torch.rand(B, 1024, generator=g) < 0.125).float()
    ri = (torch.rand(B, 1024, generator=g) < 0.125).float()
    return [fpc.to(dev), mrpc, igt, rpc_d, fpcb.to(dev), rpcb.to(dev), fi.to(dev), ri.to(dev)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"])
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from puzzlenet_b200 import _lib
    from puzzlenet_b200.model5_b import TouchedRegraster
    from puzzlenet_b200.training import Trainer
    from puzzlenet_b200.weights import synthetic_state_dict
    model = TouchedRegraster(types.SimpleNamespace(dataset="vase", loss_mode=1, loss_sum=False, lr=a.lr))
    model.load_state_dict(synthetic_state_dict(0))
    model.to(dev)
    tr = Trainer(model, precision=a.precision)
    batch = make_training_batch(a.pairs, 64 + rank, dev)
    losses = []
    for _ in range(a.warmup):
        losses.append(tr.training_step(batch)["loss"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = _lib.load().pz_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses.append(tr.training_step(batch)["loss"])
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.load().pz_launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # phase split (one extra step, synchronised between phases)
    ph = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    tr.forward_backward(batch)
    ev[1].record()
    w = tr.all_reduce_grads()
    ev[2].record()
    tr.optimizer_step(w)
    ev[3].record()
    torch.cuda.synchronize()
    ph = {"forward_backward_ms": ev[0].elapsed_time(ev[1]), "allreduce_ms": ev[1].elapsed_time(ev[2]),
          "adam_ms": ev[2].elapsed_time(ev[3])}
    if rank == 0:
        t = ms.item()
        # dense FLOPs: forward 7.347 GFLOP/pair, backward = 2x forward
        gflop = 3 * 7.347 * a.pairs
        print(json.dumps({"metric": "pairs/sec PuzzleNet training step (config 4)", "value": a.pairs * world / t * 1e3,
                          "unit": "pairs/s", "n_gpus": world, "pairs_per_gpu": a.pairs, "ms_per_step": t,
                          "steps": a.steps, "warmup": a.warmup, "dtype": "f32" if a.precision == "fp32" else "tf32 tensor-core GEMMs, fp32 storage/accumulate", "scaling": "weak",
                          "achieved_tflops_per_gpu": gflop / t, "gpu_launches_per_step": launches / a.steps,
                          "grad_elems_allreduced": tr.flat.n, "phases": ph,
                          "loss_first_last": [losses[0], losses[-1]], "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
