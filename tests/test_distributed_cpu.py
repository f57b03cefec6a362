"""World-size-2 gloo test of the multi-GPU host logic (point-range MSM shards folded on the host; column sharding).

No GPU here: the per-shard compute is the oracle (checker standing in for the device), the fold is the product's
host code (zkb_g1_sum) and the plumbing is the product's distributed.py.
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ragged, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle import coracle
    from util import random_field

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    nn = n + (3 if ragged else 0)
    s = random_field(nn, 1)
    bases = coracle.g1_fixed_base_mul(random_field(nn, 2))
    got = zd.sharded_msm(s, lambda off, sl: coracle.best_multiexp(sl, bases[off:off + sl.shape[0]]))
    want = coracle.best_multiexp(s, bases)
    ok_msm = bool((got == want).all())
    # column sharding
    cols = [random_field(1 << 6, 10 + i) for i in range(5)]
    w = coracle.fr_omega(6)
    mine = zd.sharded_columns(cols, lambda cs: [coracle.best_fft(c, w, 6) for c in cs])
    ok_cols = sorted(mine) == list(range(rank, 5, world)) and all((mine[i] == coracle.best_fft(cols[i], w, 6)).all() for i in mine)
    ret[rank] = (ok_msm, ok_cols)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_sharded_msm_and_columns_gloo(ragged):
    import torch.multiprocessing as mp

    world = 2
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 64, ragged, ret), nprocs=world, join=True)
    assert all(ret[r] == (True, True) for r in range(world)), dict(ret)


def test_point_range_partition():
    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    for n in (0, 1, 7, 8, 1 << 10, (1 << 10) + 5):
        for world in (1, 2, 3, 8):
            spans = [zd.point_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(l for _, l in spans) == n
            for (o1, l1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + l1 == o2
    assert zd.columns_for_rank(5, 1, 2) == [1, 3]


# ---- one NTT sharded over 2 ranks: the kernels' distributed addressing, run by the CPU emulator in two gloo processes whose
# slices live in POSIX shared memory (the stand-in for CUDA IPC peer mappings); gloo barriers replace the device barriers ----
def _sharded_ntt_worker(rank, world, port, k, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from multiprocessing import shared_memory

    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import emu
    from oracle import coracle
    from util import random_field

    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    coracle.build()
    a = random_field(1 << k, 900 + k)
    w = coracle.fr_omega(k)
    off, ln = zd.ntt_slice(k, rank, world)
    mine = [shared_memory.SharedMemory(create=True, size=ln * 32) for _ in range(3)]  # A, W, O slices of this rank
    names = [None] * world
    dist.all_gather_object(names, [m.name for m in mine])
    peers = [[shared_memory.SharedMemory(name=nm) for nm in names[r]] if r != rank else mine for r in range(world)]
    A, W, O = ([np.ndarray((ln, 4), dtype=np.uint64, buffer=peers[r][b].buf) for r in range(world)] for b in range(3))
    A[rank][:] = a[off:off + ln]
    log_g = world.bit_length() - 1
    dist.barrier()                                                    # every rank's input slice is in place
    rc0 = emu.ntt_dist_phase(0, rank, log_g, k, w, A, W, O)           # pass 0: peer loads + peer stores
    dist.barrier()
    rc1 = emu.ntt_dist_phase(1, rank, log_g, k, w, A, W, O)           # local passes + scattering final pass
    dist.barrier()
    ret[rank] = (rc0, rc1, bool((O[rank] == coracle.best_fft(a, w, k)[off:off + ln]).all()))
    dist.barrier()
    del A, W, O
    for r in range(world):
        if r != rank:
            for m in peers[r]:
                m.close()
    for m in mine:
        m.close()
        m.unlink()
    dist.destroy_process_group()


def test_sharded_ntt_two_ranks_shared_memory():
    import torch.multiprocessing as mp

    import emu
    from oracle import coracle
    emu.build()       # compile once in the parent so the two ranks do not race on the shared objects
    coracle.build()
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_ntt_worker, args=(world, port, 13, ret), nprocs=world, join=True)
    assert all(ret[r] == (0, 0, True) for r in range(world)), dict(ret)


def test_ntt_slice_partition():
    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    for k in (3, 11, 20):
        for world in (1, 2, 4, 8):
            spans = [zd.ntt_slice(k, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(l for _, l in spans) == 1 << k
            assert all(o1 + l1 == o2 for (o1, l1), (o2, _) in zip(spans, spans[1:]))


# ---- quotient evaluation sharded by rows: ring halo exchange over gloo + the device code's row-window mode (CPU emulator) --------
def _sharded_quotient_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import emu
    import graph_cases as GC
    from oracle import coracle

    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    isize, rs = 128, 4
    c = GC.random_case(777, isize, rs, ngates=4, depth=5)   # the same case on every rank; each keeps only its rows
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = coracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                  None, None, None, y, rs, prev)
    off, rows = zd.row_range(isize, rank, world)

    def t(a):
        return torch.from_numpy(a[off:off + rows].view(np.int64).copy())

    sq = zd.ShardedQuotient(g, rs)
    values = t(prev)

    def evaluate(v, f, a, i):   # the per-rank call, on the emulator: same lowering, same per-thread interpreter, window mode
        rc, o, _ = emu.graph_evaluate(g, [x.numpy().view(np.uint64) for x in f], [x.numpy().view(np.uint64) for x in a],
                                      [x.numpy().view(np.uint64) for x in i], ch, None, None, None, y, rs, v.numpy().view(np.uint64),
                                      halo=(sq.halo_lo, sq.halo_hi))
        assert rc == 0
        v.copy_(torch.from_numpy(o.view(np.int64)))

    sq.run(values, [t(x) for x in fx], [t(x) for x in ad], [t(x) for x in ins], evaluate=evaluate)
    ok = bool((values.numpy().view(np.uint64) == want[off:off + rows]).all())
    # columns that already live in padded buffers: the in-place exchange writes the same halos
    cols = [t(x) for x in fx + ad + ins]
    ref = zd.exchange_halos(cols, sq.halo_lo, sq.halo_hi)
    bufs = []
    for col in cols:
        buf, view = sq.alloc_column(rows)
        buf.zero_()
        view.copy_(col)
        bufs.append(buf)
    zd.exchange_halos_inplace(bufs, sq.halo_lo, sq.halo_hi)
    ok &= all(bool((a == b).all()) for a, b in zip(ref, bufs))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_quotient_two_ranks_gloo():
    import torch.multiprocessing as mp

    import emu
    from oracle import coracle
    emu.build()
    coracle.build()
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_quotient_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] is True for r in range(world)), dict(ret)
