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


# ---- one NTT sharded over the ranks --------------------------------------------------------------
def ntt_slice(log_n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous natural-order slice [offset, offset+len) of a 2^log_n transform owned by `rank` (input and output)."""
    assert world & (world - 1) == 0 and (1 << log_n) >= world, "world must be a power of two not larger than the transform"
    ln = (1 << log_n) // world
    return rank * ln, ln


class ShardedNtt:
    """best_fft of one vector sharded over the process group: rank r passes a[r N/G : (r+1) N/G] and receives the same
    slice of the result.  The exchange runs inside the NTT kernels over NVLink peer mappings (csrc/dist.cu); the process
    group is used once, to all-gather the CUDA IPC handles."""

    def __init__(self, max_log_n: int, group=None, device=None):
        import ctypes

        import torch
        import torch.distributed as dist

        self.lib = halo2.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = 256  # ZKB_DIST_HANDLE_BYTES
        blob = (ctypes.c_uint8 * nbytes)()
        halo2.check(self.lib.zkb_dist_create(self.rank, self.world, max_log_n, blob))
        t = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).clone()
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=group)
        allb = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        buf = (ctypes.c_uint8 * len(allb)).from_buffer_copy(allb)
        halo2.check(self.lib.zkb_dist_connect(buf))
        self.max_log_n = max_log_n

    def best_fft_slice(self, a_slice: np.ndarray, omega_: np.ndarray, log_n: int) -> np.ndarray:
        a = np.ascontiguousarray(a_slice, dtype=np.uint64).reshape(-1, 4)
        _, ln = ntt_slice(log_n, self.rank, self.world)
        assert a.shape[0] == ln, "assertion failed: slice.len() == (1 << log_n) / world"
        out = np.empty_like(a)
        w = np.ascontiguousarray(omega_, dtype=np.uint64).reshape(4)
        halo2.check(self.lib.zkb_dist_ntt_fr(halo2._p(a), halo2._p(out), halo2._p(w), log_n))
        return out

    def close(self) -> None:
        self.lib.zkb_dist_destroy()


# ---- quotient evaluation sharded by rows -------------------------------------------------------------------------------------
def row_range(isize: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous rows [offset, offset + len) of the extended domain owned by `rank` (isize is a power of two, world divides it)."""
    assert isize % world == 0, "the extended domain must split evenly over the ranks"
    return rank * (isize // world), isize // world


def exchange_halos(slices: list, halo_lo: int, halo_hi: int, group=None) -> list:
    """Ring halo exchange — the one collective step of the row-sharded quotient evaluation.

    slices: this rank's row slice of every column, torch tensors of shape (rows, 4) int64 (limbs) on the device the process
    group's backend moves (CUDA for nccl, CPU for gloo).  Returns, per column, a (halo_lo + rows + halo_hi, 4) tensor: the last
    halo_lo rows of the previous rank's slice, the slice, the first halo_hi rows of the next rank's (the domain is cyclic: rank 0's
    predecessor is the last rank).  world = 1 wraps onto the slice itself.  All columns travel in two messages per neighbour."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if not slices:
        return []
    rows = slices[0].shape[0]
    assert halo_lo <= rows and halo_hi <= rows, "a halo longer than a shard would need more than the neighbouring rank"
    head = torch.stack([t[:halo_hi] for t in slices]) if halo_hi else None      # becomes the previous rank's upper halo
    tail = torch.stack([t[rows - halo_lo:] for t in slices]) if halo_lo else None  # becomes the next rank's lower halo
    lo = torch.empty_like(tail) if halo_lo else None
    hi = torch.empty_like(head) if halo_hi else None
    if world == 1:
        if halo_lo:
            lo.copy_(tail)
        if halo_hi:
            hi.copy_(head)
    else:
        prev, nxt = (rank - 1) % world, (rank + 1) % world
        ops = []
        if halo_hi:
            ops += [dist.P2POp(dist.isend, head.contiguous(), prev, group), dist.P2POp(dist.irecv, hi, nxt, group)]
        if halo_lo:
            ops += [dist.P2POp(dist.isend, tail.contiguous(), nxt, group), dist.P2POp(dist.irecv, lo, prev, group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    out = []
    for c, t in enumerate(slices):
        parts = ([lo[c]] if halo_lo else []) + [t] + ([hi[c]] if halo_hi else [])
        out.append(torch.cat(parts).contiguous())
    return out


def exchange_halos_inplace(padded: list, halo_lo: int, halo_hi: int, group=None) -> None:
    """The same exchange for columns that already live in padded buffers (halo_lo + rows + halo_hi, 4): only the halo rows are
    written, the rows in the middle are never copied.  This is the layout to keep the extended cosets in when the quotient
    evaluation is row-sharded: the exchange then moves (halo_lo + halo_hi) x 32 bytes per column and nothing else."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if not padded or (halo_lo == 0 and halo_hi == 0):
        return
    rows = padded[0].shape[0] - halo_lo - halo_hi
    assert halo_lo <= rows and halo_hi <= rows
    head = torch.stack([t[halo_lo:halo_lo + halo_hi] for t in padded]) if halo_hi else None
    tail = torch.stack([t[halo_lo + rows - halo_lo:halo_lo + rows] for t in padded]) if halo_lo else None
    lo = torch.empty_like(tail) if halo_lo else None
    hi = torch.empty_like(head) if halo_hi else None
    if world == 1:
        lo, hi = tail, head
    else:
        prev, nxt = (rank - 1) % world, (rank + 1) % world
        ops = []
        if halo_hi:
            ops += [dist.P2POp(dist.isend, head, prev, group), dist.P2POp(dist.irecv, hi, nxt, group)]
        if halo_lo:
            ops += [dist.P2POp(dist.isend, tail, nxt, group), dist.P2POp(dist.irecv, lo, prev, group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for c, t in enumerate(padded):
        if halo_lo:
            t[:halo_lo].copy_(lo[c])
        if halo_hi:
            t[halo_lo + rows:].copy_(hi[c])


class ShardedQuotient:
    """evaluate_h's row loop sharded by rows over the GPUs of a box: rank r owns rows [r N/G, (r+1) N/G) of every extended coset
    (the layout the sharded NTT leaves its output in) and of h.  Rotated rows near a shard boundary come from the neighbouring rank:
    one ring halo exchange of max|rotation| x rot_scale rows per column (a few KB against the 32 N/G bytes of a slice), then
    every rank runs zkb_graph_evaluate_dev on its window.  `evaluate` is the injectable per-rank call (the CPU tests pass the
    device-code emulator)."""

    def __init__(self, graph, rot_scale: int, group=None):
        self.graph, self.rot_scale, self.group = graph, rot_scale, group
        self.halo_lo, self.halo_hi = graph.rotation_span(rot_scale)

    def run(self, values, fixed=(), advice=(), instance=(), evaluate=None, **scalars):
        """values / columns: this rank's (rows, 4) int64 limb tensors; values is overwritten with the row results."""
        import torch

        cols = list(fixed) + list(advice) + list(instance)
        padded = exchange_halos(cols, self.halo_lo, self.halo_hi, self.group)
        nf, na = len(fixed), len(advice)
        pf, pa, pi = padded[:nf], padded[nf:nf + na], padded[nf + na:]
        rows = values.shape[0]
        if evaluate is None:
            stream = torch.cuda.current_stream().cuda_stream

            def evaluate(v, f, a, i):
                self.graph.evaluate_dev(v.data_ptr(), rows, [t.data_ptr() for t in f], [t.data_ptr() for t in a],
                                        [t.data_ptr() for t in i], rot_scale=self.rot_scale, window=True, halo_lo=self.halo_lo,
                                        halo_hi=self.halo_hi, stream=stream, **scalars)
        evaluate(values, pf, pa, pi)
        return values

    def alloc_column(self, rows: int, device=None):
        """(padded buffer, view of its `rows` data rows): produce the coset slice directly into the view (e.g. as the output of
        the sharded NTT) and pass the padded buffer to run_padded — no copy of the slice is ever made."""
        import torch

        buf = torch.empty((self.halo_lo + rows + self.halo_hi, 4), dtype=torch.int64, device=device)
        return buf, buf[self.halo_lo:self.halo_lo + rows]

    def run_padded(self, values, fixed=(), advice=(), instance=(), **scalars):
        """Columns are padded buffers from alloc_column; only their halo rows are exchanged and written."""
        import torch

        cols = list(fixed) + list(advice) + list(instance)
        exchange_halos_inplace(cols, self.halo_lo, self.halo_hi, self.group)
        rows = values.shape[0]
        self.graph.evaluate_dev(values.data_ptr(), rows, [t.data_ptr() for t in fixed], [t.data_ptr() for t in advice],
                                [t.data_ptr() for t in instance], rot_scale=self.rot_scale, window=True, halo_lo=self.halo_lo,
                                halo_hi=self.halo_hi, stream=torch.cuda.current_stream().cuda_stream, **scalars)
        return values
