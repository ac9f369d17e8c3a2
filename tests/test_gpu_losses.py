"""GPU parity: loss-side kernels, the fused test_step epilogue (pz_pair_score), predict6 and the dataset-side
preprocessing, through the C ABI, vs the CPU oracle and the frozen outputs of the unmodified reference
(tests/golden/reference_epilogue.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from puzzlenet_b200.weights import make_batch, synthetic_pairs
from tests.golden_inputs import FPS_SEED, dataset_inputs, epilogue_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_epilogue.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def L():
    from puzzlenet_b200 import losses
    return losses


def _chamfer_any(a, b):
    """the oracle formula for n != m (the reference's expand_as only admits n == m)."""
    rx = (a * a).sum(-1)
    ry = (b * b).sum(-1)
    P = rx[:, :, None] + ry[:, None, :] - 2 * torch.bmm(a, b.transpose(2, 1))
    return P.min(1)[0], P.min(2)[0]


# the expanded form cancels |x|^2 + |y|^2 (~0.5) against 2 x.y: absolute error ~ a few ulp of 0.5
CH_ATOL = 3e-7


@pytest.mark.parametrize("B,n,m", [(1, 1, 1), (2, 128, 128), (3, 1024, 1024), (2, 100, 300), (1, 2500, 1030)])
def test_chamfer_vs_oracle(L, B, n, m):
    g = torch.Generator().manual_seed(n + m)
    a, b = torch.rand(B, n, 3, generator=g) - 0.5, torch.rand(B, m, 3, generator=g) - 0.5
    if n == m and n > 4:
        b[0, 3] = a[0, 1]                         # an exact hit: P ~ 0 (may come out slightly negative)
    r1, r2 = po.chamfer_loss(a, b) if n == m else _chamfer_any(a, b)
    d1, d2 = L.chamfer_loss(a.to(DEV), b.to(DEV))
    assert d1.shape == (B, m) and d2.shape == (B, n)
    np.testing.assert_allclose(d1.cpu().numpy(), r1.numpy(), rtol=0, atol=CH_ATOL)
    np.testing.assert_allclose(d2.cpu().numpy(), r2.numpy(), rtol=0, atol=CH_ATOL)


def test_chamfer_goldens_and_empty(L, gold):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    ep = epilogue_inputs(2, po.se3_exp)
    d1, d2 = L.chamfer_loss(ep["fpcb"].to(DEV), ep["rpcb"].to(DEV))
    np.testing.assert_allclose(d1.cpu().numpy(), gold["chamfer_128_d1"], rtol=0, atol=CH_ATOL)
    np.testing.assert_allclose(d2.cpu().numpy(), gold["chamfer_128_d2"], rtol=0, atol=CH_ATOL)
    d1, d2 = L.chamfer_loss(fpc.to(DEV), mrpc.to(DEV))
    np.testing.assert_allclose(d1.cpu().numpy(), gold["chamfer_1024_d1"], rtol=0, atol=CH_ATOL)
    np.testing.assert_allclose(d2.cpu().numpy(), gold["chamfer_1024_d2"], rtol=0, atol=CH_ATOL)
    e1, e2 = L.chamfer_loss(torch.empty(0, 8, 3, device=DEV), torch.empty(0, 8, 3, device=DEV))
    assert e1.shape == (0, 8) and e2.shape == (0, 8)
    with pytest.raises((ValueError, RuntimeError)):
        L.chamfer_loss(torch.empty(1, 0, 3, device=DEV), torch.rand(1, 4, 3, device=DEV))
    with pytest.raises(RuntimeError):
        L.chamfer_loss(torch.rand(1, 4, 3), torch.rand(1, 4, 3))          # CPU tensors: no fallback


def test_chamfer_backward_vs_autograd(L):
    g = torch.Generator().manual_seed(3)
    a = (torch.rand(2, 200, 3, generator=g) - 0.5).requires_grad_()
    b = (torch.rand(2, 200, 3, generator=g) - 0.5).requires_grad_()
    w1, w2 = torch.rand(2, 200, generator=g), torch.rand(2, 200, generator=g)
    r1, r2 = po.chamfer_loss(a, b)
    ((r1 * w1).sum() + (r2 * w2).sum()).backward()
    ac, bc = a.detach().to(DEV).requires_grad_(), b.detach().to(DEV).requires_grad_()
    d1, d2 = L.chamfer_loss(ac, bc)
    ((d1 * w1.to(DEV)).sum() + (d2 * w2.to(DEV)).sum()).backward()
    np.testing.assert_allclose(ac.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bc.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_chamfer_full_size_symmetry(L):
    """B=64 x 1024^2 (the training shape): swapping the arguments swaps the outputs bit for bit, and the diagonal
    hit of chamfer(x, x) is ~0."""
    fpc, mrpc = synthetic_pairs(64, seed=65)
    a, b = fpc.to(DEV), mrpc.to(DEV)
    d1, d2 = L.chamfer_loss(a, b)
    e1, e2 = L.chamfer_loss(b, a)
    assert torch.equal(d1, e2) and torch.equal(d2, e1)
    s1, s2 = L.chamfer_loss(a, a)
    assert s1.abs().max().item() < 1e-6 and torch.equal(s1, s2)


def test_comp_and_transform(L, gold):
    ep = epilogue_inputs(2, po.se3_exp)
    got = L.comp(po.se3_exp(ep["twist2"]).to(DEV), ep["igt"].to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), gold["comp"], rtol=2e-6)
    g = po.se3_exp(ep["twist"])
    pts = ep["rpc"]
    ref = po.se3_transform(g, pts.permute(0, 2, 1)).permute(0, 2, 1)
    got = L.transform_points(g.to(DEV), pts.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=0, atol=2e-7)


def _same_selection(idx_got, idx_ref, key_ref, k):
    """Equal as ordered lists, or -- when the reference keys tie at the selection threshold or inside it --
    equal key multisets (torch.topk leaves tie order unspecified)."""
    idx_got, idx_ref = idx_got.cpu(), idx_ref.cpu()
    for b in range(idx_ref.shape[0]):
        if torch.equal(idx_got[b], idx_ref[b]):
            continue
        assert torch.equal(key_ref[b][idx_got[b]].sort().values, key_ref[b][idx_ref[b]].sort().values), b
        assert len(set(idx_got[b].tolist())) == k


def test_boundary_topk_and_topk(L):
    g = torch.Generator().manual_seed(8)
    logits = torch.randn(5, 2, 1024, generator=g) * 3
    logits[1, :, 10] = logits[1, :, 700]                  # tie in probability
    logits[2] = 0                                          # everything ties: lowest indices win
    idx = L.boundary_topk(logits.to(DEV), 128)
    assert idx.dtype == torch.int64 and idx.shape == (5, 128)
    ref = po.boundary_topk(logits, 128)
    # the CUDA softmax uses expf, torch's CPU kernel a vectorised exp: probabilities agree to ~1 ulp, so compare
    # selections through the reference probabilities with that slack at the threshold
    p = po.boundary_prob(logits)
    assert torch.equal(idx[2].cpu(), torch.arange(128))
    for b in range(5):
        got, want = set(idx[b].tolist()), set(ref[b].tolist())
        thr = p[b][ref[b][-1]]
        for i in got ^ want:
            assert abs(p[b][i] - thr) <= 2e-7 * max(thr, 1e-30) + 1e-12, (b, i)
    v = torch.randn(4, 1000, generator=g)
    v[0, 5] = v[0, 900]
    for largest in (True, False):
        vals, idx = L.topk(v.to(DEV), 128, largest=largest)
        rv, ri = torch.sort(v if not largest else -v, dim=1, stable=True)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri[:, :128].numpy())
        np.testing.assert_array_equal(vals.cpu().numpy(), torch.gather(v, 1, ri[:, :128]).numpy())
    vals, idx = L.topk(torch.tensor([[3.0, -1.0, 2.0]], device=DEV), 3, largest=False)
    assert idx.tolist() == [[1, 2, 0]]


def _scores_vs_oracle(L, o, fpc, ep, B):
    s, idx_f, idx_m, bnd_f, bnd_m = L.pair_score(
        o["out"].to(DEV), o["de_fpcb"].to(DEV), o["de_mrpcb"].to(DEV), fpc.to(DEV), ep["rpc"].to(DEV),
        ep["fpcb"].to(DEV), ep["rpcb"].to(DEV), ep["fpc_idx"].to(DEV), ep["rpc_idx"].to(DEV), ep["igt"].to(DEV),
        return_boundaries=True)
    r = po.test_step_scores(o["out"], o["de_fpcb"], o["de_mrpcb"], fpc, ep["rpc"], ep["fpcb"], ep["rpcb"],
                            ep["fpc_idx"], ep["rpc_idx"], ep["igt"])
    s = s.cpu()
    assert torch.equal(idx_f.cpu(), r["idx_f"]) and torch.equal(idx_m.cpu(), r["idx_m"])
    np.testing.assert_array_equal(bnd_f.cpu().numpy(), r["bnd_f"].numpy())
    np.testing.assert_allclose(bnd_m.cpu().numpy(), r["bnd_m"].numpy(), rtol=1e-6, atol=1e-6)
    # fp32 acos((tr-1)/2) has a 0.03 deg granularity near 0 but is well conditioned at these angles
    np.testing.assert_allclose(s[:, 0].numpy(), r["r_iso"].numpy(), rtol=1e-5, atol=1e-4)
    for col, key in ((1, "t_iso"), (2, "t_mse"), (3, "t_mae")):
        np.testing.assert_allclose(s[:, col].numpy(), r[key].numpy(), rtol=1e-5, atol=1e-6)
    for col, key in ((4, "inter_f"), (5, "union_f"), (6, "inter_m"), (7, "union_m")):
        np.testing.assert_array_equal(s[:, col].numpy(), r[key].numpy())
    for col, key in ((8, "cd_fpc"), (9, "cd_rpc"), (10, "cd_pair")):
        np.testing.assert_allclose(s[:, col].numpy(), r[key].numpy(), rtol=1e-5, atol=5e-7)
    return s


def test_pair_score_vs_oracle(L, state_dict):
    B = 3
    fpc, mrpc = synthetic_pairs(B, seed=66)
    ep = epilogue_inputs(B, po.se3_exp)
    torch.manual_seed(FPS_SEED)
    o = po.predict5(state_dict, fpc, mrpc)
    _scores_vs_oracle(L, o, fpc, ep, B)


def test_pair_score_optional_inputs(L):
    """assembly mode: no ground truth -> only cd_pair (and the selections) are produced."""
    g = torch.Generator().manual_seed(4)
    B = 2
    out6 = torch.randn(B, 6, generator=g) * 0.2
    lf, lm = torch.randn(B, 2, 1024, generator=g), torch.randn(B, 2, 1024, generator=g)
    fpc, mrpc = synthetic_pairs(B, seed=67)
    s = L.pair_score(out6.to(DEV), lf.to(DEV), lm.to(DEV), fpc.to(DEV), mrpc.to(DEV)).cpu()
    assert torch.all(s[:, :4] == 0) and torch.all(s[:, 8:10] == 0)
    mat = po.se3_exp(out6)
    idx_f, idx_m = po.boundary_topk(lf), po.boundary_topk(lm)
    bf = torch.gather(fpc, 1, idx_f.unsqueeze(-1).repeat(1, 1, 3))
    bm = po.se3_transform(mat, torch.gather(mrpc, 1, idx_m.unsqueeze(-1).repeat(1, 1, 3)).permute(0, 2, 1)).permute(0, 2, 1)
    c1, c2 = po.chamfer_loss(bf, bm)
    np.testing.assert_allclose(s[:, 10].numpy(), (c1.mean(1) + c2.mean(1)).numpy(), rtol=1e-5, atol=5e-7)
    assert L.pair_score(out6[:0].to(DEV), lf[:0].to(DEV), lm[:0].to(DEV), fpc[:0].to(DEV), mrpc[:0].to(DEV)).shape == (0, 12)


def test_test_step_golden(cuda_model, gold):
    """model.test_step (one pz_predict5 + one pz_pair_score) against the reference's own test_step output."""
    fpc, mrpc = synthetic_pairs(2, seed=64)
    ep = epilogue_inputs(2, po.se3_exp)
    batch = [t.to(DEV) for t in (fpc, mrpc, ep["igt"], ep["rpc"], ep["fpcb"], ep["rpcb"], ep["fpc_idx"], ep["rpc_idx"])]
    cuda_model.precision = "fp32"
    torch.manual_seed(FPS_SEED)
    got = cuda_model.test_step(batch, 0).cpu().numpy()[0]
    ref = gold["test_step"][0]
    assert got.shape == (10,)
    np.testing.assert_allclose(got[[6, 7]], ref[[6, 7]], rtol=1e-6)                 # IoU: integer counts
    np.testing.assert_allclose(got[[8, 9]], ref[[8, 9]], rtol=1e-4)                 # boundary chamfer
    np.testing.assert_allclose(got[[2, 3, 5]], ref[[2, 3, 5]], rtol=1e-4)           # translation errors
    np.testing.assert_allclose(got[4], ref[4], atol=0.01)                            # rotation (deg)
    np.testing.assert_allclose(got[[0, 1]], ref[[0, 1]], rtol=1e-3)                 # Euler-angle errors (scipy)


def test_predict6_golden(cuda_model, gold):
    fpc, mrpc = synthetic_pairs(2, seed=64)
    cuda_model.precision = "fp32"
    torch.manual_seed(FPS_SEED)
    out = cuda_model.predict6(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0, pretrain=True)
    ref = gold["predict6"]
    assert (np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()) < 1e-4
    torch.manual_seed(FPS_SEED)
    r = cuda_model.predict6(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0, need=True, pretrain=True)
    assert len(r) == 6 and r[1] == [0] and r[2].shape == (2, 256, 3) and r[3].shape == (2, 256, 256)
    assert torch.equal(r[0], out)
    with pytest.raises(AttributeError):
        cuda_model.predict6(make_batch(fpc.to(DEV), mrpc.to(DEV)), 0, pretrain=False)


def test_pair_score_full_batch_properties(L):
    """B=64: permuting the pairs permutes the rows; deterministic run to run."""
    g = torch.Generator().manual_seed(12)
    B = 64
    fpc, mrpc = synthetic_pairs(B, seed=68)
    ep = epilogue_inputs(B, po.se3_exp)
    out6 = torch.randn(B, 6, generator=g) * 0.3
    lf, lm = torch.randn(B, 2, 1024, generator=g), torch.randn(B, 2, 1024, generator=g)
    args = [out6, lf, lm, fpc, ep["rpc"], ep["fpcb"], ep["rpcb"], ep["fpc_idx"], ep["rpc_idx"], ep["igt"]]
    s1 = L.pair_score(*[t.to(DEV) for t in args])
    s2 = L.pair_score(*[t.to(DEV) for t in args])
    assert torch.equal(s1, s2)
    perm = torch.randperm(B, generator=g)
    s3 = L.pair_score(*[t[perm].to(DEV) for t in args])
    assert torch.equal(s3, s1[perm.to(DEV)])
    assert torch.all(s1[:, 5] >= 128) and torch.all(s1[:, 4] <= 128)


# ------------------------------------------------------------------ dataset side (rows A14 / F1)

def test_dataset_fps_golden(gold):
    from puzzlenet_b200 import dataset as D
    cloud = dataset_inputs()
    np.random.seed(21)
    up, down = D.plane_split(cloud)
    assert up.shape[0] == gold["split_up_n"] and down.shape[0] == gold["split_down_n"]
    np.testing.assert_array_equal(up[:64], gold["split_up_head"])
    np.testing.assert_array_equal(down[:64], gold["split_down_head"])
    np.random.seed(22)
    up_s, down_s = D.fps(up, 1024), D.fps(down, 1024)
    assert isinstance(up_s, np.ndarray)
    np.testing.assert_array_equal(up_s, gold["ds_fps_up"])
    np.testing.assert_array_equal(down_s, gold["ds_fps_down"])
    assert D.fps(cloud[:100], 1024) is None                       # dataset.py:1148-1149
    fb, rb, fi, ri = D.get_boundary(torch.from_numpy(down_s).to(DEV), torch.from_numpy(up_s).to(DEV))
    # chamfer values carry ~1e-7 rounding differences vs MKL's bmm: compare the selected sets through the
    # oracle's chamfer values with that slack at the 128th place
    cd1, cd2 = po.chamfer_loss(torch.from_numpy(down_s)[None], torch.from_numpy(up_s)[None])
    for mask_got, mask_ref, cd in ((fi, gold["gb_fpc_idx"], cd2[0]), (ri, gold["gb_rpc_idx"], cd1[0])):
        got, ref = mask_got.cpu().numpy() > 0, mask_ref > 0
        assert got.sum() == 128
        thr = np.sort(cd.numpy())[127]
        for i in np.nonzero(got ^ ref)[0]:
            assert abs(cd[i].item() - thr) < 5e-7
    assert fb.shape == (128, 3) and rb.shape == (128, 3)


def test_fps_batch_ragged():
    from puzzlenet_b200 import dataset as D
    g = torch.Generator().manual_seed(31)
    pieces = [(torch.rand(n, 3, generator=g) - 0.5).numpy() for n in (1500, 11000, 1024, 4097)]
    starts = [7, 10999, 0, 123]
    got = D.fps_batch(pieces, 1024, starts=starts).cpu().numpy()
    for i, p in enumerate(pieces):
        if p.shape[0] > 5000:
            ref = p[po.farthest_point_sample(torch.from_numpy(p)[None], 1024, start=torch.tensor([starts[i]]))[0].numpy()]
        else:
            ref = po.dataset_fps(p, 1024, start=starts[i])
        np.testing.assert_array_equal(got[i], ref)


def test_make_pair_matches_oracle_pipeline():
    from puzzlenet_b200 import dataset as D
    cloud = dataset_inputs()
    np.random.seed(40)
    torch.manual_seed(41)
    down, mup, igt, up, fpcb, rpcb, fpc_idx, rpc_idx = D.make_pair(cloud, D.RandomTransformSE3(0.8))
    np.random.seed(40)
    torch.manual_seed(41)
    u, d = po.plane_split(cloud)
    while u.shape[0] < 1024 or d.shape[0] < 1024:
        u, d = po.plane_split(cloud)
    u, d = po.dataset_fps(u, 1024), po.dataset_fps(d, 1024)
    np.testing.assert_array_equal(up.cpu().numpy(), u)
    np.testing.assert_array_equal(down.cpu().numpy(), d)
    x = torch.randn(1, 6)
    x = x / x.norm(p=2, dim=1, keepdim=True) * 0.8
    g = po.se3_exp(x)
    np.testing.assert_allclose(igt.cpu().numpy(), g[0].numpy(), atol=1e-6)
    ref_mup = po.se3_transform(g, torch.from_numpy(u).T[None])[0].T
    np.testing.assert_allclose(mup.cpu().numpy(), ref_mup.numpy(), atol=1e-6)
    assert fpc_idx.sum().item() == 128 and rpc_idx.sum().item() == 128
    # the boundary points are rows of the sampled halves
    assert set(map(tuple, fpcb.cpu().numpy().tolist())) == set(map(tuple, down[fpc_idx.bool()].cpu().numpy().tolist()))
    assert set(map(tuple, rpcb.cpu().numpy().tolist())) == set(map(tuple, up[rpc_idx.bool()].cpu().numpy().tolist()))


def test_make_pair_batch_matches_per_sample_pipeline():
    """the batched dataset pipeline (one cut / FPS / chamfer / top-k / transform launch for P pieces) against the
    oracle's per-sample pipeline fed with the same random draws in the documented stage order"""
    from puzzlenet_b200 import dataset as D
    g = torch.Generator().manual_seed(50)
    pieces = [(torch.randn(n, 3, generator=g) * 0.3).numpy() for n in (6000, 4500, 7001)]
    np.random.seed(60)
    torch.manual_seed(61)
    down, mup, igt, up, downb, upb, fpc_idx, rpc_idx = D.make_pair_batch(pieces, 0.8)
    assert down.shape == (3, 1024, 3) and igt.shape == (3, 4, 4) and downb.shape == (3, 128, 3)
    np.random.seed(60)
    torch.manual_seed(61)
    def cut(p):
        nrm = np.random.rand(3, 1)
        z = np.random.rand(1) / 3
        dis = np.dot(p, nrm) + z
        return p[(dis >= 0)[:, 0]], p[(dis < 0)[:, 0]]

    halves = [cut(p) for p in pieces]                     # stage 1: every piece once, in order ...
    bad = [i for i, (u, d) in enumerate(halves) if u.shape[0] < 1024 or d.shape[0] < 1024]
    rounds = 0
    while bad:                                            # ... then re-draws for the failed ones, in piece order
        rounds += 1
        for i in bad:
            halves[i] = cut(pieces[i])
        bad = [i for i in bad if halves[i][0].shape[0] < 1024 or halves[i][1].shape[0] < 1024]
    assert rounds >= 1                                    # this seed exercises the re-draw path
    starts = [(np.random.randint(0, u.shape[0]), np.random.randint(0, d.shape[0])) for u, d in halves]
    for i, ((u, d), (su, sd)) in enumerate(zip(halves, starts)):
        us, ds = po.dataset_fps(u, 1024, start=su), po.dataset_fps(d, 1024, start=sd)
        np.testing.assert_array_equal(up[i].cpu().numpy(), us)
        np.testing.assert_array_equal(down[i].cpu().numpy(), ds)
        x = torch.randn(1, 6)
        x = x / x.norm(p=2, dim=1, keepdim=True) * 0.8
        torch.randn(1, 6)
        gi = po.se3_exp(x)
        np.testing.assert_allclose(igt[i].cpu().numpy(), gi[0].numpy(), atol=1e-6)
        ref_mup = po.se3_transform(gi, torch.from_numpy(us).T[None])[0].T
        np.testing.assert_allclose(mup[i].cpu().numpy(), ref_mup.numpy(), atol=1e-6)
        fb, rb, fi, ri = po.get_boundary(torch.from_numpy(ds), torch.from_numpy(us))
        cd1, cd2 = po.chamfer_loss(torch.from_numpy(ds)[None], torch.from_numpy(us)[None])
        for got, ref, cd in ((fpc_idx[i], fi, cd2[0]), (rpc_idx[i], ri, cd1[0])):
            gm, rm = got.cpu().numpy() > 0, ref.numpy() > 0
            assert gm.sum() == 128
            thr = np.sort(cd.numpy())[127]
            for j in np.nonzero(gm ^ rm)[0]:
                assert abs(cd[j].item() - thr) < 5e-7
    # the tuple feeds the model directly
    assert fpc_idx.sum().item() == 3 * 128 and torch.isfinite(mup).all()
