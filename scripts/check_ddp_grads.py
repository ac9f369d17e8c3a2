"""Numerical check of the data-parallel training step on N GPUs (run under torchrun): every rank back-propagates its
own batch, the bucketed all-reduce (started from inside backward) sums the flat gradient buffer, and the result is
compared with the sum of the per-batch gradients computed locally, one batch after the other, without any collective.
Also checks that all ranks hold identical parameters after the Adam step.  One JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/check_ddp_grads.py"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from puzzlenet_b200.model5_b import TouchedRegraster
    from puzzlenet_b200.training import Trainer
    from puzzlenet_b200.weights import synthetic_state_dict
    from scripts.bench_train import make_training_batch

    def fresh():
        m = TouchedRegraster(types.SimpleNamespace(dataset="vase", loss_mode=1, loss_sum=False, lr=1e-5))
        m.load_state_dict(synthetic_state_dict(0))
        return m.to(dev)

    B = 8
    batches = [make_training_batch(B, 100 + r, dev) for r in range(world)]
    g = torch.Generator().manual_seed(3)
    starts = [torch.stack([torch.randint(0, n, (B,), generator=g) for n in (1024, 512, 1024, 512)]) for _ in range(world)]
    out = {}
    for prec in ("fp32", "tf32"):
        # distributed: own batch, overlapped buckets
        tr = Trainer(fresh(), precision=prec)
        tr.forward_backward(batches[rank], starts[rank])
        tr.all_reduce_grads()
        torch.cuda.synchronize()
        got = tr.flat.grads.clone()
        # local reference: every batch in turn, summed, no collective
        ref_tr = Trainer(fresh(), precision=prec)
        ref_tr.overlap_allreduce = False
        total = torch.zeros_like(got)
        for r in range(world):
            ref_tr.forward_backward(batches[r], starts[r])
            total += ref_tr.flat.grads
        err = ((got - total).norm() / total.norm()).item()
        # parameters after one optimizer step must agree across ranks bit for bit
        tr.optimizer_step(world)
        mine = tr.flat.params.clone()
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out[prec] = {"grad_rel_l2_vs_local_sum": err, "params_identical_across_ranks": bool(torch.equal(lo, hi))}
    if rank == 0:
        print(json.dumps({"check": "bucketed all-reduce == sum of per-batch gradients", "world": world, "pairs_per_rank": B,
                          **out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
