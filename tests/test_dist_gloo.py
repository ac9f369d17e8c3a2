"""CPU, world_size 2, gloo: the pair-sharding host logic (partition + the single all_gather of results)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from puzzlenet_b200 import sharding


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 7, 64, 496, 513):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_all_pairs_count():
    p = sharding.all_pairs(32)
    assert p.shape == (496, 2) and (p[:, 0] < p[:, 1]).all() and len({tuple(x) for x in p.tolist()}) == 496


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def score(lo, hi):                       # stand-in for predict5 on pairs [lo, hi): row = f(pair id)
            ids = torch.arange(lo, hi, dtype=torch.float32)
            return torch.stack([ids, ids * 2 + 1, torch.full_like(ids, float(rank))], dim=1)
        out = sharding.run_sharded(n_items, score)
        q.put((rank, out.tolist()))          # plain lists: tensors through a Queue need the sender alive
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [496, 7, 1])
def test_run_sharded_world2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_items) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = torch.arange(n_items, dtype=torch.float32)
    for rank in (0, 1):
        out = torch.tensor(results[rank], dtype=torch.float32).reshape(n_items, 3)
        assert out.shape == (n_items, 3)
        assert torch.equal(out[:, 0], ids) and torch.equal(out[:, 1], ids * 2 + 1)
        lo, hi = sharding.shard_bounds(n_items, 0, 2)
        assert (out[lo:hi, 2] == 0).all() and (out[hi:, 2] == 1).all()     # rows came from the right rank


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from puzzlenet_b200.training import all_reduce_bucket
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        works = [all_reduce_bucket(flat, lo, hi) for lo, hi in ((0, 320), (320, 704), (704, 1000), (10, 10))]
        assert works[-1] is None                                   # empty bucket: nothing to do
        for w in works:
            if w is not None:
                w.wait()
        q.put((rank, flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_all_reduce_world2_gloo():
    """the training step's bucketed gradient all-reduce (heads / Encoder / Encoder2 segments of the flat buffer)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(1000, dtype=torch.float32) * 3).tolist()
    assert res[0] == want and res[1] == want


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        from puzzlenet_b200.model5_b import TouchedRegraster
        from puzzlenet_b200.training import Trainer
        torch.manual_seed(100 + rank)                  # every rank draws DIFFERENT initial weights
        model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
        with torch.no_grad():
            model.Encoder.bn1.running_mean.fill_(float(rank + 1))
        tr = Trainer(model)                            # broadcasts rank 0's parameters / buffers / Adam state
        first = (float(model.Encoder.mlp3.weight.double().sum()), float(model.dt.sum()),
                 float(model.Encoder.bn1.running_mean.sum()), float(tr.flat.params.double().sum()))
        # resume where only rank 0 holds the checkpoint's optimizer state
        sd = tr.state_dict()
        if rank == 0:
            sd["step"], sd["exp_avg"] = 7, torch.full_like(sd["exp_avg"], 0.5)
        tr.load_state_dict(sd)
        q.put((rank, first, tr.step_count, float(tr.flat.exp_avg.double().sum())))
    finally:
        dist.destroy_process_group()


def test_trainer_broadcasts_replica_state_world2_gloo():
    """Trainer construction and load_state_dict leave every rank with rank 0's parameters, BN buffers and Adam state
    (what Lightning DDP does for the reference); ranks start from different seeds here."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 25500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]
    assert res[0][1] == 7 and res[0][0][2] == 1024.0       # step from rank 0, running_mean = rank 0's fill (1.0 * 1024)
