"""Drop-in for the reference's ``pointnet_util`` on the PuzzleNet hot path.

Same function names, argument order, defaults and return shapes as
``pointnet_util.py`` of the reference (file:line cited per function); every
function launches hand-written sm_100a kernels through the C ABI
(``include/puzzlenet_b200.h``).  CUDA tensors only -- there is no CPU fallback.

Index outputs are ``torch.int64`` like the reference.  FPS consumes exactly one
``torch.randint(0, N, (B,))`` from the CPU default generator per call
(pointnet_util.py:65), so seeding reproduces the reference's centroids.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()


_WS_CACHE = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch per (device, stream): the fused ops need hundreds of MB at dataset shapes and a fresh
    cudaMalloc per call would dominate their run time."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), device=device, dtype=torch.uint8)
        _WS_CACHE[key] = ws
    return ws


def _i64c(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t)
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64 (got {t.dtype})")
    return t.contiguous()


def square_distance(src, dst):
    """pointnet_util.py:22-36 -- src [B,S,3], dst [B,N,3] -> [B,S,N] squared distances."""
    src, dst = _f32c(src, "src"), _f32c(dst, "dst")
    B, S, C = src.shape
    if C != 3 or dst.shape[0] != B or dst.shape[2] != 3:
        raise ValueError(f"square_distance expects [B,S,3] and [B,N,3] (got {tuple(src.shape)}, {tuple(dst.shape)})")
    N = dst.shape[1]
    out = torch.empty(B, S, N, device=src.device, dtype=torch.float32)
    with torch.cuda.device(src.device):
        _lib.call("pz_sqdist", src.data_ptr(), dst.data_ptr(), B, S, N, out.data_ptr(), _lib.stream_ptr())
    return out


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _gather_rows(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    B, N, Cc = points.shape
    M = idx.numel() // B if B else 0
    out = torch.empty(B, M, Cc, device=points.device, dtype=points.dtype)
    with torch.cuda.device(points.device):
        _lib.call("pz_gather", points.data_ptr(), idx.data_ptr(), B, N, Cc, M, points.element_size(), out.data_ptr(),
                  _lib.stream_ptr())
    return out.reshape(*idx.shape, Cc)


def _scatter_rows(grad: torch.Tensor, c0: int, C: int, idx: torch.Tensor, N: int, dst: torch.Tensor) -> None:
    """dst[b, idx[b, m], 0:C] += grad[b, m, c0:c0+C] -- the backward of a row gather (pz_scatter_add_rows)."""
    B = idx.shape[0]
    M = idx.numel() // B
    with torch.cuda.device(grad.device):
        _lib.call("pz_scatter_add_rows", grad.data_ptr(), grad.shape[-1], c0, C, idx.data_ptr(), B * M, M, N,
                  dst.data_ptr(), dst.shape[-1], _lib.stream_ptr())


class _IndexPoints(torch.autograd.Function):
    """index_points with the gradient the reference's advanced indexing has (pointnet_util.py:39-50):
    d points[b, idx[b, m]] += d out[b, m]."""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.n = points.shape[1]
        return _gather_rows(points, idx)

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        grad = grad.contiguous().float()
        Cc = grad.shape[-1]
        dpoints = torch.zeros(idx.shape[0], ctx.n, Cc, device=grad.device, dtype=torch.float32)
        _scatter_rows(grad.reshape(idx.shape[0], -1, Cc), 0, Cc, idx, ctx.n, dpoints)
        return dpoints, None


def index_points(points, idx):
    """pointnet_util.py:39-50 -- points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,(K,)C].  Differentiable in ``points``
    (float32) like the reference's indexing."""
    _lib.require_cuda(points, idx)
    points = points.contiguous()
    idx = _i64c(idx, "idx")
    if _needs_grad(points):
        if points.dtype != torch.float32:
            raise TypeError("index_points: gradients are implemented for float32 points only")
        return _IndexPoints.apply(points, idx)
    return _gather_rows(points, idx)


def _draw_start(B: int, N: int, device) -> torch.Tensor:
    # pointnet_util.py:65: drawn on the CPU generator, then moved to the device
    return torch.randint(0, N, (B,), dtype=torch.long).to(device)


def _fps(xyz: torch.Tensor, npoint: int, start: torch.Tensor, want_xyz: bool):
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, device=xyz.device, dtype=torch.int64)
    new_xyz = torch.empty(B, npoint, 3, device=xyz.device, dtype=torch.float32) if want_xyz else None
    with torch.cuda.device(xyz.device):
        _lib.call("pz_fps", xyz.data_ptr(), B, N, start.data_ptr(), npoint, idx.data_ptr(),
                  new_xyz.data_ptr() if want_xyz else None, _lib.stream_ptr())
    return idx, new_xyz


def farthest_point_sample(xyz, npoint):
    """pointnet_util.py:53-73 -- xyz [B,N,3] -> centroids [B,npoint] (int64), bit-exact with the
    reference's CPU result for the same start draw."""
    xyz = _f32c(xyz, "xyz")
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise ValueError(f"farthest_point_sample expects [B,N,3] (got {tuple(xyz.shape)})")
    start = _draw_start(xyz.shape[0], xyz.shape[1], xyz.device)
    return _fps(xyz, int(npoint), start, False)[0]


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet_util.py:76-96 -- first ``nsample`` in-radius indices per query, padded with the first."""
    xyz, new_xyz = _f32c(xyz, "xyz"), _f32c(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, int(nsample), device=xyz.device, dtype=torch.int64)
    with torch.cuda.device(xyz.device):
        _lib.call("pz_ball_query", xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, ctypes.c_float(float(radius)),
                  int(nsample), out.data_ptr(), _lib.stream_ptr())
    return out


def knn_point(nsample, xyz, new_xyz, return_dist=False):
    """The kNN branch of sample_and_group (pointnet_util.py:118-119) as one call:
    ``square_distance(new_xyz, xyz).argsort()[:, :, :nsample]`` without the [B,S,N] matrix.
    Ascending by (distance, index)."""
    xyz, new_xyz = _f32c(xyz, "xyz"), _f32c(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, int(nsample), device=xyz.device, dtype=torch.int64)
    d2 = torch.empty(B, S, int(nsample), device=xyz.device, dtype=torch.float32) if return_dist else None
    with torch.cuda.device(xyz.device):
        _lib.call("pz_knn", new_xyz.data_ptr(), xyz.data_ptr(), B, S, N, int(nsample), idx.data_ptr(),
                  d2.data_ptr() if return_dist else None, _lib.stream_ptr())
    return (idx, d2) if return_dist else idx


def _sample_and_group_forward(S, radius, K, xyz, feat, returnfps, knn):
    B, N, _ = xyz.shape
    start = _draw_start(B, N, xyz.device)
    fps_idx, new_xyz = _fps(xyz, S, start, True)
    idx = knn_point(K, xyz, new_xyz) if knn else query_ball_point(radius, K, xyz, new_xyz)
    D = feat.shape[-1] if feat is not None else 0
    new_points = torch.empty(B, S, K, 3 + D, device=xyz.device, dtype=torch.float32)
    grouped_xyz = torch.empty(B, S, K, 3, device=xyz.device, dtype=torch.float32) if returnfps else None
    with torch.cuda.device(xyz.device):
        _lib.call("pz_group_concat", xyz.data_ptr(), feat.data_ptr() if feat is not None else None,
                  new_xyz.data_ptr(), idx.data_ptr(), B, N, D, S, K, new_points.data_ptr(),
                  grouped_xyz.data_ptr() if returnfps else None, _lib.stream_ptr())
    return new_xyz, new_points, grouped_xyz, fps_idx, idx


class _SampleAndGroup(torch.autograd.Function):
    """sample_and_group with the reference's gradients (pointnet_util.py:115-130): the indices are constants;
    ``new_points = cat(xyz[idx] - new_xyz[:, :, None], points[idx])`` and ``new_xyz = xyz[fps_idx]`` are differentiable
    in ``points`` and ``xyz``.  In the model the gradient of mlp1/mlp2/bn1/bn2 (stage 1) and of mlp3/mlp4 (stage 2)
    flows ONLY through this function (model5_b.py:449-461)."""

    @staticmethod
    def forward(ctx, xyz, feat, S, radius, K, returnfps, knn):
        new_xyz, new_points, grouped_xyz, fps_idx, idx = _sample_and_group_forward(S, radius, K, xyz, feat, returnfps, knn)
        ctx.save_for_backward(idx, fps_idx)
        ctx.shape = (xyz.shape[1], feat.shape[-1] if feat is not None else 0, returnfps)
        ctx.mark_non_differentiable(fps_idx)
        if returnfps:
            return new_xyz, new_points, grouped_xyz, fps_idx
        return new_xyz, new_points, fps_idx

    @staticmethod
    def backward(ctx, g_new_xyz, g_new_points, *rest):
        idx, fps_idx = ctx.saved_tensors
        N, D, returnfps = ctx.shape
        B, S, K = idx.shape
        dev = idx.device
        g_grouped = rest[0] if returnfps else None
        need_xyz, need_feat = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and D > 0
        d_xyz = d_feat = None
        gp = g_new_points.contiguous().float() if g_new_points is not None else None
        if need_feat and gp is not None:
            d_feat = torch.zeros(B, N, D, device=dev, dtype=torch.float32)
            _scatter_rows(gp.reshape(B, S * K, 3 + D), 3, D, idx, N, d_feat)
        elif need_feat:
            d_feat = torch.zeros(B, N, D, device=dev, dtype=torch.float32)
        if need_xyz:
            d_xyz = torch.zeros(B, N, 3, device=dev, dtype=torch.float32)
            d_centre = torch.zeros(B, S, 3, device=dev, dtype=torch.float32)
            if gp is not None:
                _scatter_rows(gp.reshape(B, S * K, 3 + D), 0, 3, idx, N, d_xyz)            # through xyz[idx]
                d_centre -= gp[..., :3].sum(dim=2)                                          # - sum_k through new_xyz
            if g_new_xyz is not None:
                d_centre += g_new_xyz.float()
            _scatter_rows(d_centre, 0, 3, fps_idx, N, d_xyz)                                # new_xyz = xyz[fps_idx]
            if g_grouped is not None:
                _scatter_rows(g_grouped.contiguous().float().reshape(B, S * K, 3), 0, 3, idx, N, d_xyz)
        return d_xyz, d_feat, None, None, None, None, None


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, knn=False):
    """pointnet_util.py:99-136.

    Returns ``new_xyz [B,npoint,3]``, ``new_points [B,npoint,nsample,3+D]`` (xyz-relative first,
    then features) and, with ``returnfps=True``, also ``grouped_xyz`` and ``fps_idx``.  Differentiable in ``points``
    and ``xyz`` exactly as the reference's gather / subtract / cat sequence is (the indices are constants).
    """
    xyz = _f32c(xyz, "xyz")
    S, K = int(npoint), int(nsample)
    feat = _f32c(points, "points") if points is not None else None
    if _needs_grad(xyz, feat):
        r = _SampleAndGroup.apply(xyz, feat, S, radius, K, bool(returnfps), bool(knn))
        return r if returnfps else r[:2]
    new_xyz, new_points, grouped_xyz, fps_idx, _ = _sample_and_group_forward(S, radius, K, xyz, feat, returnfps, knn)
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet_util.py:139-156 -- a pure view/concat (no kernel needed)."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device, dtype=xyz.dtype)
    grouped = xyz.view(B, 1, N, C)
    new_points = grouped if points is None else torch.cat([grouped, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, new_points


def group_mlp_maxpool(xyz, points, new_xyz, idx, w1, b1, w2, b2, precision=_lib.PZ_PREC_FP32):
    """Fused form of ``sample_and_group``'s grouping + ``relu(mlp_b(relu(mlp_a(.)))).max(-2)``
    (model5_b.py:449-454): never materialises [B,S,K,3+D].  Returns [B,S,C2]."""
    if _needs_grad(xyz, points, new_xyz, w1, b1, w2, b2):
        raise RuntimeError("group_mlp_maxpool is the fused INFERENCE op and records no autograd graph: call it under "
                           "torch.no_grad() / with detached tensors, or use sample_and_group (differentiable) followed by "
                           "the layers; the train-mode forward/backward of the model lives in TouchedRegraster.training_step")
    xyz, points, new_xyz = _f32c(xyz, "xyz"), _f32c(points, "points"), _f32c(new_xyz, "new_xyz")
    idx = _i64c(idx, "idx")
    w1, b1, w2, b2 = (_f32c(t, "weight") for t in (w1, b1, w2, b2))
    B, N, _ = xyz.shape
    D = points.shape[-1]
    S, K = idx.shape[1], idx.shape[2]
    C1, C2 = w1.shape[0], w2.shape[0]
    lib = _lib.load()
    ws_bytes = lib.pz_group_mlp_workspace_bytes(B, N, D, S, K, C1, C2)
    ws = _workspace(ws_bytes, xyz.device)
    out = torch.empty(B, S, C2, device=xyz.device, dtype=torch.float32)
    with torch.cuda.device(xyz.device):
        _lib.call("pz_group_mlp_maxpool", xyz.data_ptr(), points.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(),
                  w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), B, N, D, S, K, C1, C2, int(precision),
                  out.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr())
    return out


# ----------------------------------------------------------------------------------------------------------------
# PointNet++ blocks of the reference (pointnet_util.py:159-315).  model5_b never instantiates them (SURVEY.md §2:
# dead code on the live path); they are mirrored for API completeness (row F4) on the same kernels: FPS / kNN /
# ball query / gathers for the geometry, pz_linear for every 1x1 convolution with its eval-mode BatchNorm folded in,
# pz_maxpool_forward for the neighbourhood max.  Inference only: train-mode BatchNorm2d raises.
# ----------------------------------------------------------------------------------------------------------------
import torch.nn as nn  # noqa: E402


def _folded_linear(conv, bn):
    """1x1 conv (weight [O,I,1(,1)]) followed by eval-mode BatchNorm -> (W [O,I], b [O]) of the equivalent Linear."""
    if bn.training:
        raise NotImplementedError("puzzlenet_b200 PointNet++ blocks run in eval mode only (call .eval())")
    w = conv.weight.detach().reshape(conv.weight.shape[0], -1).float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return (w * scale[:, None]).contiguous(), ((b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()).contiguous()


def _mlp_rows(x, convs, bns):
    """relu(bn(conv(.))) for every layer on rows x [M, C] (the reference applies them as 1x1 convolutions)."""
    M = x.shape[0]
    for conv, bn in zip(convs, bns):
        w, b = _folded_linear(conv, bn)
        y = torch.empty(M, w.shape[0], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.call("pz_linear", x.data_ptr(), x.shape[1], w.data_ptr(), b.data_ptr(), M, w.shape[0], w.shape[1], 1,
                      None, 0, y.data_ptr(), w.shape[0], _lib.PZ_PREC_FP32, _lib.stream_ptr())
        x = y
    return x


def _max_over_neighbours(x, G, K):
    """x [G*K, C] -> [G, C] = max over the K consecutive rows of a group"""
    C = x.shape[1]
    y = torch.empty(G, C, device=x.device, dtype=torch.float32)
    arg = torch.empty(G, C, device=x.device, dtype=torch.int32)
    with torch.cuda.device(x.device):
        _lib.call("pz_maxpool_forward", x.data_ptr(), G, K, C, y.data_ptr(), arg.data_ptr(), _lib.stream_ptr())
    return y


class PointNetSetAbstraction(nn.Module):
    """pointnet_util.py:159-197."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, knn=False):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.knn, self.group_all = npoint, radius, nsample, knn, group_all
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel

    def forward(self, xyz, points):
        if self.group_all:
            new_xyz, new_points = sample_and_group_all(xyz, points)
        else:
            new_xyz, new_points = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points, knn=self.knn)
        B, S, K, C = new_points.shape
        rows = _mlp_rows(_f32c(new_points, "new_points").reshape(B * S * K, C), self.mlp_convs, self.mlp_bns)
        return new_xyz, _max_over_neighbours(rows, B * S, K).view(B, S, -1)


class PointNetSetAbstractionMsg(nn.Module):
    """pointnet_util.py:200-264 (multi-scale grouping; features first, then relative xyz)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, knn=False):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list, self.knn = npoint, radius_list, nsample_list, knn
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel + 3
            for out_channel in mlp:
                convs.append(nn.Conv2d(last, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points, seed_idx=None):
        xyz = _f32c(xyz, "xyz")
        B, N, C = xyz.shape
        S = self.npoint
        new_xyz = index_points(xyz, farthest_point_sample(xyz, S) if seed_idx is None else seed_idx)
        outs = []
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            idx = knn_point(K, xyz, new_xyz) if self.knn else query_ball_point(radius, K, xyz, new_xyz)
            grouped_xyz = index_points(xyz, idx) - new_xyz.view(B, S, 1, C)
            grouped = grouped_xyz if points is None else torch.cat([index_points(_f32c(points, "points"), idx), grouped_xyz], dim=-1)
            rows = _mlp_rows(grouped.reshape(B * S * K, -1).contiguous(), self.conv_blocks[i], self.bn_blocks[i])
            outs.append(_max_over_neighbours(rows, B * S, K).view(B, S, -1))
        return new_xyz, torch.cat(outs, dim=-1)


class PointNetFeaturePropagation(nn.Module):
    """pointnet_util.py:268-315 -- inverse-distance interpolation from the 3 nearest sampled points + an MLP.
    Channel-first tensors ([B,C,N]) in and out, like the reference."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        xyz1, xyz2 = _f32c(xyz1.permute(0, 2, 1), "xyz1"), _f32c(xyz2.permute(0, 2, 1), "xyz2")
        points2 = _f32c(points2.permute(0, 2, 1), "points2")
        B, N, _ = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated = points2.repeat(1, N, 1)
        else:
            idx, d2 = knn_point(3, xyz2, xyz1, return_dist=True)            # the 3 nearest of xyz2 for every xyz1 point
            recip = 1.0 / (d2 + 1e-8)
            weight = recip / recip.sum(dim=2, keepdim=True)
            interpolated = (index_points(points2, idx) * weight.view(B, N, 3, 1)).sum(dim=2)
        new_points = interpolated if points1 is None else torch.cat([_f32c(points1.permute(0, 2, 1), "points1"), interpolated], dim=-1)
        rows = _mlp_rows(new_points.reshape(B * N, -1).contiguous(), self.mlp_convs, self.mlp_bns)
        return rows.view(B, N, -1).permute(0, 2, 1)
