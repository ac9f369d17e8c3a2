"""CPU: host logic of the assembly driver -- greedy merge / pose composition on synthetic score tables and the
sharded all-pairs scoring with a stand-in scorer over gloo (world size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from puzzlenet_b200 import assembly, sharding


def test_se3_exp_np_matches_oracle():
    from oracle import puzzle_oracle as po
    rng = np.random.default_rng(0)
    for scale in (1e-3, 0.3, 2.0):
        x = rng.normal(size=6) * scale
        ref = po.se3_exp(torch.tensor(x, dtype=torch.float64)[None])[0].numpy()
        np.testing.assert_allclose(assembly._se3_exp_np(x), ref, atol=1e-12)


def test_greedy_assemble_chain_recovers_poses():
    """Pieces 0-1-2-3 in a chain with exact pair twists and low scores on the chain edges: the composed poses must
    equal the products of the pair poses; high-score edges are never used."""
    rng = np.random.default_rng(1)
    P = 4
    pairs = sharding.all_pairs(P)
    tw = {e: rng.normal(size=6) * 0.4 for e in [(0, 1), (1, 2), (2, 3)]}
    rows = np.zeros((pairs.shape[0], assembly.ROW_COLS))
    for r, (i, j) in enumerate(pairs.tolist()):
        rows[r, 7] = 1
        if (i, j) in tw:
            rows[r, :6] = tw[(i, j)]
            rows[r, 6] = 0.01 * (1 + i)
        else:
            rows[r, :6] = rng.normal(size=6)
            rows[r, 6] = 5.0
    poses, comp, merges = assembly.greedy_assemble(P, pairs, rows)
    assert [m[:2] for m in merges] == [(0, 1), (1, 2), (2, 3)] and len(set(comp.tolist())) == 1
    g01, g12, g23 = (assembly._se3_exp_np(tw[e]) for e in [(0, 1), (1, 2), (2, 3)])
    np.testing.assert_allclose(poses[0], np.eye(4), atol=1e-12)
    np.testing.assert_allclose(poses[1], g01, atol=1e-12)
    np.testing.assert_allclose(poses[2], g01 @ g12, atol=1e-12)
    np.testing.assert_allclose(poses[3], g01 @ g12 @ g23, atol=1e-12)


def test_greedy_assemble_merges_components_in_any_order():
    """(2,3) first, then (0,1), then (1,2): the second component is re-expressed as a whole."""
    rng = np.random.default_rng(2)
    pairs = sharding.all_pairs(4)
    tw = {(2, 3): rng.normal(size=6) * 0.3, (0, 1): rng.normal(size=6) * 0.3, (1, 2): rng.normal(size=6) * 0.3}
    score = {(2, 3): 0.1, (0, 1): 0.2, (1, 2): 0.3}
    rows = np.zeros((6, assembly.ROW_COLS))
    for r, e in enumerate(map(tuple, pairs.tolist())):
        rows[r, 7] = 1
        rows[r, 6] = score.get(e, 9.0)
        rows[r, :6] = tw.get(e, np.zeros(6))
    poses, comp, merges = assembly.greedy_assemble(4, pairs, rows, max_score=1.0)
    assert [m[:2] for m in merges] == [(2, 3), (0, 1), (1, 2)]
    g = {e: assembly._se3_exp_np(t) for e, t in tw.items()}
    np.testing.assert_allclose(poses[2], g[(0, 1)] @ g[(1, 2)], atol=1e-12)
    np.testing.assert_allclose(poses[3], g[(0, 1)] @ g[(1, 2)] @ g[(2, 3)], atol=1e-12)
    # threshold: nothing below it -> nothing merges
    poses, comp, merges = assembly.greedy_assemble(4, pairs, rows, max_score=0.05)
    assert merges == [] and len(set(comp.tolist())) == 4
    # invalid / non-finite rows are skipped
    rows[:, 7] = 0
    assert assembly.greedy_assemble(4, pairs, rows)[2] == []


def _fake_scorer(fpc, mrpc):
    """deterministic stand-in for the network: the 'pose' and score depend only on the two clouds"""
    rows = torch.zeros(fpc.shape[0], assembly.ROW_COLS)
    rows[:, :3] = fpc.mean(1)
    rows[:, 3:6] = mrpc.mean(1)
    rows[:, 6] = (fpc.mean((1, 2)) - mrpc.mean((1, 2))).abs()
    rows[:, 7] = 1
    return rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clouds = torch.rand(7, 1024, 3, generator=torch.Generator().manual_seed(3))
        pairs, rows = assembly.score_all_pairs(clouds, _fake_scorer, batch=4)
        q.put((rank, pairs.tolist(), rows.tolist()))
    finally:
        dist.destroy_process_group()


def test_score_all_pairs_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    clouds = torch.rand(7, 1024, 3, generator=torch.Generator().manual_seed(3))
    pairs = sharding.all_pairs(7)
    want = assembly.score_pairs(clouds, pairs, _fake_scorer, batch=5)          # single process, other batching
    for _, got_pairs, got_rows in res:
        assert got_pairs == pairs.tolist()
        assert torch.equal(torch.tensor(got_rows), want)
