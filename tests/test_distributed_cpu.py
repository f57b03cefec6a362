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
