"""GPU parity: geometry kernels (through the C ABI via the pointnet_util mirror) vs the CPU oracle and
the frozen reference outputs.  Indices must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import puzzle_oracle as po
from tests.golden_inputs import golden_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pu():
    from puzzlenet_b200 import pointnet_util
    return pointnet_util


def _knn_sets_equal(idx_a, d_a, idx_b, d_b):
    """Ordered equality unless distances tie; with ties compare neighbour sets."""
    idx_a, idx_b = idx_a.cpu(), idx_b.cpu()
    if torch.equal(idx_a, idx_b):
        return True
    assert torch.equal(d_a.cpu(), d_b.cpu()), "kNN distances differ"
    return torch.equal(idx_a.sort(-1).values, idx_b.sort(-1).values)


def test_fps_goldens(pu, goldens):
    xyz, _, big, _ = golden_inputs()
    torch.manual_seed(11)
    got = pu.farthest_point_sample(xyz.to(DEV), 64)
    assert got.dtype == torch.int64
    assert np.array_equal(got.cpu().numpy(), goldens["fps_small"])
    torch.manual_seed(12)
    got = pu.farthest_point_sample(big.to(DEV), 1024)          # C3 shape: 11000 -> 1024
    assert np.array_equal(got.cpu().numpy(), goldens["fps_11000_1024"])


@pytest.mark.parametrize("B,N,S", [(1, 1, 1), (2, 33, 33), (3, 257, 64), (4, 512, 256), (8, 1024, 512),
                                   (2, 2049, 100), (1, 5000, 300), (1, 16384, 64)])
def test_fps_vs_oracle(pu, B, N, S):
    g = torch.Generator().manual_seed(N * 7 + S)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    if N > 40:
        xyz[0, 5] = xyz[0, 20]                                  # duplicates -> argmax ties
        xyz[-1, 30:36] = xyz[-1, 2]
    torch.manual_seed(N)
    ref = po.farthest_point_sample(xyz, S)
    torch.manual_seed(N)
    got = pu.farthest_point_sample(xyz.to(DEV), S)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("B,N,S", [(3, 11000, 1024), (2, 4097, 200), (1, 14336, 512), (2, 8000, 2048)])
def test_fps_large_cloud_block_pruning(pu, B, N, S):
    """4096 < N <= 14336 takes fps_grid_kernel (Morton-binned cloud, blocks skipped when the centroid's distance to their
    bounding box is >= their maximum running minimum): the sampled indices must stay identical to the oracle's, including
    dense clusters, exact duplicates (arg-max ties -> lowest index) and a cloud whose points all coincide."""
    g = torch.Generator().manual_seed(N * 3 + S)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    xyz[0, : N // 3] = xyz[0, : N // 3] * 0.02 + 0.4              # a dense cluster
    xyz[-1, 100:400] = xyz[-1, 7]                                 # 300 duplicates of one point
    if B > 2:
        xyz[1] = 0.25                                             # degenerate cloud
    torch.manual_seed(N + 1)
    ref = po.farthest_point_sample(xyz, S)
    torch.manual_seed(N + 1)
    got = pu.farthest_point_sample(xyz.to(DEV), S)
    assert torch.equal(got.cpu(), ref)


def test_fps_degenerate_cloud(pu):
    xyz = torch.zeros(2, 64, 3)                                  # all points identical: argmax of zeros -> 0
    torch.manual_seed(3)
    ref = po.farthest_point_sample(xyz, 16)
    torch.manual_seed(3)
    assert torch.equal(pu.farthest_point_sample(xyz.to(DEV), 16).cpu(), ref)


def test_fps_too_large_fails_loudly(pu):
    with pytest.raises(RuntimeError, match="16384"):
        pu.farthest_point_sample(torch.zeros(1, 20000, 3, device=DEV), 4)


def test_square_distance(pu, goldens):
    xyz, _, _, _ = golden_inputs()
    got = pu.square_distance(xyz[:, :40].to(DEV), xyz.to(DEV))
    assert np.array_equal(got.cpu().numpy(), goldens["sqdist"])
    a = torch.rand(2, 130, 3) - 0.5
    b = torch.rand(2, 1000, 3) - 0.5
    assert torch.equal(pu.square_distance(a.to(DEV), b.to(DEV)).cpu(), po.square_distance(a, b))
    assert pu.square_distance(a[:, :0].to(DEV), b.to(DEV)).shape == (2, 0, 1000)


@pytest.mark.parametrize("B,S,N,K", [(2, 64, 300, 16), (3, 512, 1024, 32), (2, 256, 512, 32), (1, 40, 32, 32),
                                     (1, 7, 5000, 32), (2, 100, 2049, 1), (2, 33, 1024, 5), (1, 50, 33, 32), (2, 64, 1500, 20)])
def test_knn_vs_oracle(pu, B, S, N, K):
    g = torch.Generator().manual_seed(S + N)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    if N >= 300:
        xyz[0, 100:110] = xyz[0, 7]                               # ties
    q = xyz[:, :S].clone() if S <= N else torch.rand(B, S, 3, generator=g) - 0.5
    ref_idx, ref_d = po.knn_select(po.square_distance(q, xyz), K)
    idx, d2 = pu.knn_point(K, xyz.to(DEV), q.to(DEV), return_dist=True)
    assert idx.dtype == torch.int64 and idx.shape == (B, S, K)
    assert torch.equal(d2.cpu(), ref_d), "distances must be bit-exact (fp32, no FMA)"
    assert torch.equal(idx.cpu(), ref_idx), "order is (distance, index) ascending, same as the stable oracle"


@pytest.mark.parametrize("N,dups", [(1024, 200), (1024, 1024), (3000, 70), (512, 40)])
def test_knn_heavy_ties(pu, N, dups):
    """Many coincident points: more than 64 candidates tie under the selection bound (the incremental path of
    knn_kernel); order among equal distances is by index, as in the stable oracle."""
    g = torch.Generator().manual_seed(N + dups)
    xyz = torch.rand(2, N, 3, generator=g) - 0.5
    xyz[0, :dups] = xyz[0, 0]
    xyz[1, N - dups:] = xyz[1, N - 1]
    q = torch.cat([xyz[:, :40], xyz[:, N - 40:]], 1).clone()
    ref_idx, ref_d = po.knn_select(po.square_distance(q, xyz), 32)
    idx, d2 = pu.knn_point(32, xyz.to(DEV), q.to(DEV), return_dist=True)
    assert torch.equal(d2.cpu(), ref_d)
    assert torch.equal(idx.cpu(), ref_idx)


@pytest.mark.parametrize("B,S,N,K", [(2, 1024, 11000, 32), (1, 200, 14336, 32), (2, 64, 2049, 7), (1, 96, 4100, 32)])
def test_knn_large_cloud_block_pruning(pu, B, S, N, K):
    """N > 2048 takes knn_grid_kernel (cloud binned into a Morton-ordered grid, 128-point blocks pruned by their bounding
    boxes): indices and distances must stay bit-identical to the exhaustive oracle, including clustered points, exact
    duplicates and queries far outside the cloud."""
    g = torch.Generator().manual_seed(N + S)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    xyz[:, : N // 4] = xyz[:, : N // 4] * 0.05 + 0.3             # a dense cluster
    xyz[0, 500:560] = xyz[0, 17]                                  # duplicates (ties broken by index)
    q = xyz[:, :S].clone()
    q[:, -3:] = torch.tensor([[5.0, 5.0, 5.0], [-4.0, 0.0, 0.0], [0.3, 0.3, 0.3]])
    ref_idx, ref_d = po.knn_select(po.square_distance(q, xyz), K)
    idx, d2 = pu.knn_point(K, xyz.to(DEV), q.to(DEV), return_dist=True)
    assert torch.equal(d2.cpu(), ref_d)
    assert torch.equal(idx.cpu(), ref_idx)


def test_knn_rejects_k_above_32(pu):
    x = torch.rand(1, 100, 3, device=DEV)
    with pytest.raises(RuntimeError, match="K="):
        pu.knn_point(33, x, x)


def test_query_ball_point(pu, goldens):
    xyz, _, _, _ = golden_inputs()
    got = pu.query_ball_point(0.25, 12, xyz.to(DEV), xyz[:, :40].to(DEV))
    assert np.array_equal(got.cpu().numpy(), goldens["ball"])
    # radius 0 around non-member queries: nothing in range -> N everywhere, as the reference
    q = torch.full((2, 3, 3), 9.0)
    assert torch.equal(pu.query_ball_point(0.0, 4, xyz.to(DEV), q.to(DEV)).cpu(), po.query_ball_point(0.0, 4, xyz, q))


def test_index_points(pu, goldens):
    _, feat, _, _ = golden_inputs()
    idx = torch.randint(0, 300, (2, 9, 5), generator=torch.Generator().manual_seed(13))
    got = pu.index_points(feat.to(DEV), idx.to(DEV))
    assert np.array_equal(got.cpu().numpy(), goldens["index_points"])
    # 2-D index, int64 payload, odd row width (byte path)
    pts = torch.randint(0, 255, (2, 50, 3), dtype=torch.uint8)
    idx2 = torch.randint(0, 50, (2, 11))
    assert torch.equal(pu.index_points(pts.to(DEV), idx2.to(DEV)).cpu(), po.index_points(pts, idx2))
    pts64 = torch.arange(2 * 50 * 2).view(2, 50, 2)
    assert torch.equal(pu.index_points(pts64.to(DEV), idx2.to(DEV)).cpu(), po.index_points(pts64, idx2))


def test_sample_and_group_goldens(pu, goldens):
    xyz, feat, _, _ = golden_inputs()
    torch.manual_seed(14)
    nx, npts, gx, fi = pu.sample_and_group(32, 0, 16, xyz.to(DEV), feat.to(DEV), True, True)
    assert np.array_equal(nx.cpu().numpy(), goldens["sg_new_xyz"])
    assert np.array_equal(fi.cpu().numpy(), goldens["sg_fps_idx"])
    assert np.array_equal(npts.cpu().numpy()[0], goldens["sg_new_points"][0])
    assert np.array_equal(gx.cpu().numpy()[0], goldens["sg_grouped_xyz"][0])
    a = np.sort(npts.cpu().numpy()[1][..., :3], axis=1)           # cloud 1 holds a duplicated point
    b = np.sort(goldens["sg_new_points"][1][..., :3], axis=1)
    assert np.array_equal(a, b)
    torch.manual_seed(15)
    nx, npts = pu.sample_and_group(32, 0.3, 16, xyz.to(DEV), feat.to(DEV), False, False)
    assert np.array_equal(nx.cpu().numpy(), goldens["sgb_new_xyz"])
    assert np.array_equal(npts.cpu().numpy(), goldens["sgb_new_points"])
    # points=None
    torch.manual_seed(16)
    ref = po.sample_and_group(8, 0, 4, xyz, None, knn=True)
    torch.manual_seed(16)
    got = pu.sample_and_group(8, 0, 4, xyz.to(DEV), None, False, True)
    assert torch.equal(got[1].cpu(), ref[1])


def test_full_size_properties_c3(pu):
    """C3 shape (B=4 here for time, N=11000 -> 1024, K=32): size-independent properties at full size."""
    g = torch.Generator().manual_seed(3)
    xyz = (torch.rand(4, 11000, 3, generator=g) - 0.5).to(DEV)
    torch.manual_seed(5)
    fps = pu.farthest_point_sample(xyz, 1024)
    assert all(len(set(r.tolist())) == 1024 for r in fps.cpu())   # FPS never repeats while distinct points remain
    new_xyz = pu.index_points(xyz, fps)
    idx, d2 = pu.knn_point(32, xyz, new_xyz, return_dist=True)
    assert torch.equal(idx[:, :, 0], fps)                         # nearest neighbour of a member is itself
    assert (d2[:, :, 0] == 0).all() and (d2[:, :, 1:] >= d2[:, :, :-1]).all()
    # every returned distance is the true distance to the returned index
    chk = (pu.index_points(xyz, idx) - new_xyz[:, :, None]).pow(2)
    chk = (chk[..., 0] + chk[..., 1]) + chk[..., 2]
    assert torch.equal(chk, d2)
    # the 32nd distance bounds everything that was not selected (checked on a sample of queries)
    full = pu.square_distance(new_xyz[:, :16], xyz)
    kth = d2[:, :16, -1:]
    assert ((full < kth).sum(-1) <= 31).all() and ((full <= kth).sum(-1) >= 32).all()


def test_fps_large_cloud_variants_agree():
    """FPS at the dataset shape (11000 -> 1024) takes three code paths depending on how many clouds share the GPU: a
    4-CTA cluster per cloud (4B <= 148 SMs), a 2-CTA cluster (2B <= 148) and one CTA per cloud.  All three must pick
    the same indices for the same cloud and start (and B = 2 is pinned to the reference goldens above)."""
    from puzzlenet_b200 import _lib
    g = torch.Generator().manual_seed(11)
    xyz = (torch.rand(80, 11000, 3, generator=g) - 0.5).to(DEV)
    start = torch.randint(0, 11000, (80,), generator=g).to(DEV)

    def run(b):
        out = torch.empty(b, 1024, dtype=torch.int64, device=DEV)
        _lib.call("pz_fps", xyz[:b].contiguous().data_ptr(), b, 11000, start[:b].contiguous().data_ptr(), 1024, out.data_ptr(), None,
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        return out.cpu()

    four, two, one = run(3), run(40), run(80)
    assert torch.equal(four, two[:3]) and torch.equal(two, one[:40])
    ref = po.farthest_point_sample(xyz[:1].cpu(), 1024, start[:1].cpu())
    assert torch.equal(four[:1], ref)
