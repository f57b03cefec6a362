"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

north_star: "MSM shards by SRS point ranges, with each GPU's tiny bucket sums combined on the host, and batched
column NTTs shard by column".  Neither needs a data-path collective: the only exchange is an all_gather of one
96-byte G1 per rank (MSM) or nothing at all (independent columns).  The fold is the
`results.iter().fold(identity, |a, b| a + b)` of halo2's best_multiexp, applied to per-GPU partial sums.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from . import halo2


def point_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous SRS point range [offset, offset+len) owned by `rank` (ragged tail goes to the last ranks)."""
    base, rem = divmod(n, world)
    off = rank * base + min(rank, rem)
    return off, base + (1 if rank < rem else 0)


def columns_for_rank(ncols: int, rank: int, world: int) -> list[int]:
    """Round-robin column ownership for batched NTTs / batched commits."""
    return list(range(rank, ncols, world))


def all_gather_g1(partial: np.ndarray, group=None, device=None) -> np.ndarray:
    """all_gather one 12-limb G1 per rank -> (world, 12) uint64."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return np.stack([p.cpu().numpy().view(np.uint64) for p in parts])


def sharded_msm(scalars: np.ndarray, shard_msm: Callable[[int, np.ndarray], np.ndarray], group=None, device=None) -> np.ndarray:
    """Full MSM of `scalars` (n x 4, every rank holds the same array) against an SRS sharded by point range.

    `shard_msm(offset, scalars_slice)` computes this rank's partial sum — on a GPU box
    `lambda off, s: params.commit_range(off, s)`; the partials are gathered and folded on the host.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    off, ln = point_range(s.shape[0], rank, world)
    partial = shard_msm(off, s[off:off + ln])
    return halo2.g1_sum(all_gather_g1(partial, group, device))


def sharded_columns(cols: list, op: Callable[[list], list], group=None) -> dict[int, object]:
    """Apply a batched column op (e.g. domain.coeff_to_extended_batch) to this rank's round-robin share.
    Returns {column index: result} for the owned columns; no communication."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = columns_for_rank(len(cols), rank, world)
    out = op([cols[i] for i in mine]) if mine else []
    return dict(zip(mine, out))
