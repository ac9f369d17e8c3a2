"""GPU versions of the dataset-side preprocessing that feeds the forward.

The reference does this per sample in numpy on DataLoader workers (``dataset.py``): cut a piece with a random plane
(:761-775), farthest-point-sample each half 11000 -> 1024 with a 1024-iteration python loop (:1147-1163), find the
128 boundary points of each half with a chamfer matrix + top-k (:1357-1367) and move one half by a random rigid
motion (``se_math/transforms.py:151-196``).  Here the same steps run on the GPU through the C ABI:

* :func:`fps`              -- ``CADDataset.fps(points, npoints)``; returns the selected POINTS like the reference
* :func:`fps_batch`        -- the same for a ragged list of pieces in one launch (what assembly config 5 needs)
* :func:`plane_split`      -- ``plane_split(points, z=None)`` (host RNG as in the reference, partition by ``pz_plane_split``)
* :func:`get_boundary`     -- ``CADDataset.get_boundary(fpc, de_mrpc)``
* :class:`RandomTransformSE3` -- ``transforms.RandomTransformSE3``
* :func:`make_pair`        -- ``CADDataset.getitem_non_random`` + ``MovedCADDataset2.__getitem__`` for one piece

Random numbers are drawn from the same host generators in the same order as the reference
(``np.random`` for the cut and the FPS start, the torch CPU generator for the twist).  CUDA only, no fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, losses
from . import se3

MAX_FPS_POINTS = 16384   # pz_fps keeps the cloud and its running distances in one CTA's shared memory


def _to_cuda(points, device=None):
    if isinstance(points, np.ndarray):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return torch.from_numpy(np.ascontiguousarray(points)).to(device), True
    _lib.require_cuda(points)
    return points, False


def _fps_indices(xyz: torch.Tensor, starts: torch.Tensor, npoints: int) -> torch.Tensor:
    B, N, _ = xyz.shape
    if N > MAX_FPS_POINTS:
        raise ValueError(f"fps: at most {MAX_FPS_POINTS} points per cloud (got {N})")
    idx = torch.empty(B, npoints, device=xyz.device, dtype=torch.int64)
    with torch.cuda.device(xyz.device):
        _lib.call("pz_fps", xyz.data_ptr(), B, N, starts.data_ptr(), npoints, idx.data_ptr(), None, _lib.stream_ptr())
    return idx


def fps(points, npoints: int, start=None, device=None):
    """dataset.py:1147-1163.  points [N,D] (numpy or CUDA tensor; the first three columns are xyz) ->
    the ``npoints`` selected rows in selection order, same container type as the input; ``None`` when N < npoints.
    ``start`` None draws ``np.random.randint(0, N)`` exactly like the reference."""
    if points.shape[0] < npoints:
        return None
    N = points.shape[0]
    if start is None:
        start = np.random.randint(0, N)
    pts, was_numpy = _to_cuda(points, device)
    xyz = pts[:, :3].contiguous().float().unsqueeze(0)
    idx = _fps_indices(xyz, torch.tensor([int(start)], dtype=torch.int64, device=pts.device), npoints)[0]
    sel = pts[idx]
    return sel.cpu().numpy() if was_numpy else sel


def fps_batch(pieces, npoints: int, starts=None, device=None) -> torch.Tensor:
    """FPS of a ragged list of pieces ([N_i, 3] numpy arrays or tensors, every N_i >= npoints) in ONE launch ->
    ``[P, npoints, 3]`` on the GPU.  Shorter pieces are padded with copies of their own point 0: a copy has the same
    running distance as the original and a higher index, so the first-arg-max rule never selects it.
    ``starts`` None draws ``np.random.randint(0, N_i)`` per piece in order, like P successive reference calls."""
    P = len(pieces)
    if P == 0:
        raise ValueError("fps_batch: empty list")
    sizes = [int(p.shape[0]) for p in pieces]
    if min(sizes) < npoints:
        raise ValueError(f"fps_batch: a piece has {min(sizes)} < {npoints} points (the reference returns None there)")
    if starts is None:
        starts = [np.random.randint(0, n) for n in sizes]
    if device is None:
        first = pieces[0]
        device = first.device if isinstance(first, torch.Tensor) and first.is_cuda else \
            torch.device("cuda", torch.cuda.current_device())
    nmax = max(sizes)
    host = torch.empty(P, nmax, 3, dtype=torch.float32)
    for i, p in enumerate(pieces):
        t = (p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)))[:, :3].float().cpu()
        host[i, :sizes[i]] = t
        host[i, sizes[i]:] = t[0]
    xyz = host.to(device, non_blocking=True)
    idx = _fps_indices(xyz, torch.tensor([int(s) for s in starts], dtype=torch.int64, device=device), npoints)
    return torch.gather(xyz, 1, idx.unsqueeze(-1).expand(-1, -1, 3))


def _split_device(pts, sizes, planes, pad):
    """pz_plane_split on [P, nmax, C] clouds -> (up [P,nmax,C], down [P,nmax,C], counts [P,2] int32), all on the GPU"""
    P, nmax, C = pts.shape
    up, down = torch.empty_like(pts), torch.empty_like(pts)
    counts = torch.empty(P, 2, device=pts.device, dtype=torch.int32)
    with torch.cuda.device(pts.device):
        _lib.call("pz_plane_split", pts.data_ptr(), None if sizes is None else sizes.data_ptr(), P, nmax, C,
                  planes.data_ptr(), up.data_ptr(), down.data_ptr(), counts.data_ptr(), int(pad), _lib.stream_ptr())
    return up, down, counts


def plane_split(points, z=None, device=None):
    """dataset.py:761-775: cut with the plane ``points . normal + z = 0``, ``normal ~ U[0,1)^3`` and (unless given)
    ``z ~ U[0,1)/3`` from ``np.random`` in the reference's order -> (up, down), order preserved.  The signed
    distance is evaluated in float64 like numpy does for a float32 cloud and float64 normal."""
    normal = np.random.rand(3, 1)
    if z is None:
        z = np.random.rand(1) / 3
    pts, was_numpy = _to_cuda(points, device)
    pts = pts.contiguous().float()
    n, C = pts.shape
    plane = torch.tensor([[normal[0, 0], normal[1, 0], normal[2, 0], float(np.asarray(z).reshape(-1)[0])]],
                         dtype=torch.float64).to(pts.device)
    up_buf, down_buf, counts = _split_device(pts.unsqueeze(0), None, plane, pad=False)
    n_up, n_down = counts[0].tolist()
    up, down = up_buf[0, :n_up], down_buf[0, :n_down]
    if was_numpy:
        return up.cpu().numpy(), down.cpu().numpy()
    return up, down


def get_boundary(fpc: torch.Tensor, de_mrpc: torch.Tensor):
    """dataset.py:1357-1367.  fpc, de_mrpc [1024,3] -> (fpc boundary [128,3], mrpc boundary [128,3],
    fpc_idx [1024], rpc_idx [1024]) -- the 128 points of each cloud nearest to the other one (chamfer matrix never
    materialised: one pz_chamfer launch, two pz_topk launches)."""
    _lib.require_cuda(fpc, de_mrpc)
    cd1, cd2 = losses.chamfer_loss(fpc.unsqueeze(0), de_mrpc.unsqueeze(0))   # cd1: per de_mrpc point, cd2: per fpc point
    _, top1 = losses.topk(cd1, 128, largest=False)
    _, top2 = losses.topk(cd2, 128, largest=False)
    cdxyz1 = de_mrpc[top1[0]]
    cdxyz2 = fpc[top2[0]]
    fpc_idx = torch.zeros(fpc.shape[0], device=fpc.device)
    fpc_idx[top2[0]] = 1
    rpc_idx = torch.zeros(de_mrpc.shape[0], device=fpc.device)
    rpc_idx[top1[0]] = 1
    return cdxyz2, cdxyz1, fpc_idx, rpc_idx


class RandomTransformSE3:
    """se_math/transforms.py:151-196 -- random rigid motion; the twist comes from the torch CPU generator exactly
    as in the reference, ``exp`` and the point transform run on the GPU."""

    def __init__(self, mag=1, mag_randomly=False):
        self.mag = mag
        self.randomly = mag_randomly
        self.gt = None
        self.igt = None

    def generate_transform(self):
        amp = self.mag
        if self.randomly:
            amp = torch.rand(1, 1) * self.mag
        x = torch.randn(1, 6)
        x = x / x.norm(p=2, dim=1, keepdim=True) * amp
        self.x = x
        return x

    def apply_transform(self, p0, x):
        _lib.require_cuda(p0)
        x = x.to(p0.device)
        g = se3.exp(x)
        gt = se3.exp(-x)
        p1 = losses.transform_points(g, p0.unsqueeze(0))[0]
        self.gt = gt.squeeze(0)
        self.igt = g.squeeze(0)
        return p1

    def transform(self, tensor):
        return self.apply_transform(tensor, self.generate_transform())

    def __call__(self, tensor):
        return self.transform(tensor)

    def get_x(self):
        return self.x


def make_pair(piece, rigid_transform: RandomTransformSE3, device=None, max_tries: int = 100):
    """One training / test sample from one raw piece, all on the GPU:
    ``CADDataset.getitem_non_random`` (dataset.py:1165-1190: split until both halves have >= 1024 points, FPS both
    to 1024, boundaries) followed by ``MovedCADDataset2.__getitem__`` (dataset.py:88-105) ->
    ``(down, mup, igt, up, downb, upb, fpc_idx, rpc_idx)``: the 8-tuple ``predict5`` / ``test_step`` take."""
    pts, _ = _to_cuda(np.asarray(piece, dtype=np.float32) if isinstance(piece, np.ndarray) else piece, device)
    up, down = plane_split(pts)
    tries = 0
    while up.shape[0] < 1024 or down.shape[0] < 1024:
        tries += 1
        if tries > max_tries:
            raise RuntimeError("make_pair: no cut leaves 1024 points on both sides")
        up, down = plane_split(pts)
    up = fps(up, 1024).float()
    down = fps(down, 1024).float()
    fpcb, rpcb, fpc_idx, rpc_idx = get_boundary(down, up)
    mup = rigid_transform(up)
    igt = rigid_transform.igt
    rigid_transform(rpcb)            # the reference also moves the boundary (and discards it), consuming one twist draw
    return down, mup, igt, up, fpcb, rpcb, fpc_idx, rpc_idx


def make_pair_batch(pieces, mag: float = 0.8, device=None, max_tries: int = 100):
    """``make_pair`` for a whole batch of raw pieces with a handful of launches instead of ~25 per sample:
    one batched plane cut (re-drawn only for the pieces whose cut leaves fewer than 1024 points on a side), ONE FPS
    launch for all 2P halves, one chamfer + two top-k launches for the boundaries, one ``se3.exp`` + one transform.
    Returns the batched 8-tuple ``(down, mup, igt, up, downb, upb, fpc_idx, rpc_idx)`` = ``[P,1024,3]``,
    ``[P,1024,3]``, ``[P,4,4]``, ``[P,1024,3]``, ``[P,128,3]``, ``[P,128,3]``, ``[P,1024]``, ``[P,1024]`` that
    ``predict5`` / ``test_step`` / ``training_step`` take.

    Random numbers come from the reference's generators but are drawn per STAGE, not per sample: first the cut of every
    piece in order (``np.random.rand(3,1)``, ``np.random.rand(1)/3``; failed cuts are re-drawn afterwards, in piece
    order, until they succeed), then the two FPS starts of every piece (``np.random.randint(0, n_up)``,
    ``randint(0, n_down)``), then per piece the twist of ``mup`` and the discarded one of ``mupb`` (``torch.randn(1,6)``
    twice).  Sample by sample the arithmetic is that of :func:`make_pair`."""
    P = len(pieces)
    if P == 0:
        raise ValueError("make_pair_batch: empty list")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    sizes_h = [int(p.shape[0]) for p in pieces]
    nmax = max(sizes_h)
    if nmax > MAX_FPS_POINTS:
        raise ValueError(f"make_pair_batch: at most {MAX_FPS_POINTS} points per piece (got {nmax})")
    host = torch.zeros(P, nmax, 3, dtype=torch.float32)
    for i, p in enumerate(pieces):
        host[i, :sizes_h[i]] = (p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)))[:, :3].float().cpu()
    pts = host.to(device)
    sizes = torch.tensor(sizes_h, dtype=torch.int32, device=device)

    def draw_planes(k):
        out = np.empty((k, 4))
        for i in range(k):
            out[i, :3] = np.random.rand(3, 1)[:, 0]
            out[i, 3] = np.random.rand(1)[0] / 3
        return out

    planes_h = draw_planes(P)
    up, down, counts = _split_device(pts, sizes, torch.from_numpy(planes_h).to(device), pad=True)
    cnt = counts.cpu().numpy()
    tries = 0
    bad = np.nonzero((cnt < 1024).any(axis=1))[0]
    while len(bad):
        tries += 1
        if tries > max_tries:
            raise RuntimeError("make_pair_batch: no cut leaves 1024 points on both sides of a piece")
        idx = torch.from_numpy(bad).to(device)
        u2, d2, c2 = _split_device(pts[idx].contiguous(), sizes[idx].contiguous(),
                                   torch.from_numpy(draw_planes(len(bad))).to(device), pad=True)
        up[idx], down[idx] = u2, d2
        cnt[bad] = c2.cpu().numpy()
        bad = bad[(cnt[bad] < 1024).any(axis=1)]
    starts = torch.tensor([[np.random.randint(0, int(cnt[i, 0])), np.random.randint(0, int(cnt[i, 1]))] for i in range(P)],
                          dtype=torch.int64)
    # one FPS launch over the 2P padded halves ([up_0, down_0, up_1, ...]); padding rows duplicate row 0 of the half
    halves = torch.stack([up, down], dim=1).reshape(2 * P, nmax, 3)
    idx = _fps_indices(halves, starts.reshape(-1).to(device), 1024)
    sel = torch.gather(halves, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).reshape(P, 2, 1024, 3)
    up_s, down_s = sel[:, 0].contiguous(), sel[:, 1].contiguous()
    # boundaries: get_boundary(fpc=down, de_mrpc=up) for every piece
    cd1, cd2 = losses.chamfer_loss(down_s, up_s)              # cd1 per up point, cd2 per down point
    _, top1 = losses.topk(cd1, 128, largest=False)
    _, top2 = losses.topk(cd2, 128, largest=False)
    upb = torch.gather(up_s, 1, top1.unsqueeze(-1).expand(-1, -1, 3))
    downb = torch.gather(down_s, 1, top2.unsqueeze(-1).expand(-1, -1, 3))
    fpc_idx = torch.zeros(P, 1024, device=device).scatter_(1, top2, 1.0)
    rpc_idx = torch.zeros(P, 1024, device=device).scatter_(1, top1, 1.0)
    # rigid motion of the `up` half: per piece one twist for mup and one (discarded) for the boundary
    tw = torch.empty(P, 6)
    for i in range(P):
        x = torch.randn(1, 6)
        tw[i] = (x / x.norm(p=2, dim=1, keepdim=True) * mag)[0]
        torch.randn(1, 6)
    igt = se3.exp(tw.to(device))
    mup = losses.transform_points(igt, up_s)
    return down_s, mup, igt, up_s, downb, upb, fpc_idx, rpc_idx
