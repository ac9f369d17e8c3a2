"""Data-parallel sharding of independent piece pairs over ranks.

The forward has no data-path collective: rank r processes pairs [lo, hi) and the per-pair results
(twist [6] + scores) are gathered once at the end.  Works with any torch.distributed backend
(nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first n % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_pairs(n_pieces: int) -> torch.Tensor:
    """[P(P-1)/2, 2] unordered piece pairs (i < j) in lexicographic order -- the candidate set of the
    multi-piece assembly (BASELINE config 5: 32 pieces -> 496 pairs)."""
    i, j = torch.triu_indices(n_pieces, n_pieces, offset=1)
    return torch.stack([i, j], dim=1)


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather per-item rows computed under ``shard_bounds`` back into [n_total, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    width = -(-n_total // world)                      # pad every shard to the largest one
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r, b in enumerate(bufs):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(b[: hi - lo])
    return torch.cat(parts, dim=0)


def run_sharded(n_items: int, fn: Callable[[int, int], torch.Tensor]) -> torch.Tensor:
    """Run ``fn(lo, hi)`` on this rank's shard and gather the rows from all ranks."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_items, rank, world)
    return gather_rows(fn(lo, hi), n_items)
