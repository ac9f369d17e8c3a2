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
    up_buf, down_buf = torch.empty_like(pts), torch.empty_like(pts)
    counts = torch.empty(2, device=pts.device, dtype=torch.int32)
    with torch.cuda.device(pts.device):
        _lib.call("pz_plane_split", pts.data_ptr(), n, C, float(normal[0, 0]), float(normal[1, 0]), float(normal[2, 0]),
                  float(np.asarray(z).reshape(-1)[0]), up_buf.data_ptr(), down_buf.data_ptr(), counts.data_ptr(),
                  _lib.stream_ptr())
    n_up, n_down = counts.tolist()
    up, down = up_buf[:n_up], down_buf[:n_down]
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
