#!/usr/bin/env python
"""Row-sharded quotient evaluation: parity and timing under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29513 \
        tests/tools/dist_quotient_check.py [--time-log-n 24 --cols 4]

Parity: every rank builds the same seeded case (random gate set, rotations -2..3), keeps only its rows of every column on its
GPU, and runs distributed.ShardedQuotient (ring halo exchange over NCCL, then zkb_graph_evaluate_dev on the row window); the rows
are compared with the C oracle's whole-domain evaluation.  Timing: halo2-base's gate on `cols` advice columns over 2^log_n
extended rows split across the ranks, CUDA events around exchange + kernel, max over ranks.  One JSON line per item on rank 0.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time-log-n", type=int, default=0)
    ap.add_argument("--cols", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import graph_cases as GC
    from oracle import coracle
    from util import random_field

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    zkb.init(local)
    coracle.build()
    ok_all = True
    for seed, isize, rs in [(801, 1 << 10, 4), (802, 1 << 14, 8), (803, 64 * world, 1)]:
        c = GC.random_case(seed, isize, rs, ngates=5, depth=5)
        g = c["graph"]
        fx, ad, ins, ch, y, prev = GC.case_arrays(c)
        want = coracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                      None, None, None, y, rs, prev)
        off, rows = zd.row_range(isize, rank, world)

        def t(a):
            return torch.from_numpy(a[off:off + rows].view(np.int64).copy()).to(dev)

        sq = zd.ShardedQuotient(g, rs)
        values = t(prev)
        sq.run(values, [t(x) for x in fx], [t(x) for x in ad], [t(x) for x in ins], challenges=ch, y=y)
        torch.cuda.synchronize()
        ok = bool((values.cpu().numpy().view(np.uint64) == want[off:off + rows]).all())
        # the same with the columns living in padded buffers (only halo rows move)
        def padded(a):
            buf, view = sq.alloc_column(rows, dev)
            view.copy_(t(a))
            return buf
        values2 = t(prev)
        sq.run_padded(values2, [padded(x) for x in fx], [padded(x) for x in ad], [padded(x) for x in ins], challenges=ch, y=y)
        torch.cuda.synchronize()
        ok &= bool((values2.cpu().numpy().view(np.uint64) == want[off:off + rows]).all())
        flag = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(flag)
        ok = flag.item() == 0
        ok_all &= ok
        if rank == 0:
            print(json.dumps({"op": "sharded_quotient_parity", "isize": isize, "rot_scale": rs, "world": world, "halo": [sq.halo_lo, sq.halo_hi],
                              "parity": ok}), flush=True)
    if args.time_log_n:
        ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
        isize, rs, qc = 1 << args.time_log_n, 4, args.cols
        g = ev.GraphEvaluator()
        parts = []
        for i in range(qc):
            a_, b_, c_, d_ = (("advice", i, r) for r in range(4))
            parts.append(g.add_expression(("prod", ("fixed", i, 0), ("sum", ("sum", a_, ("prod", b_, c_)), ("neg", d_)))))
        g.add_horner(ev.ValueSource(ev.PREVIOUS), parts, ev.ValueSource(ev.Y))
        off, rows = zd.row_range(isize, rank, world)
        base = torch.from_numpy(random_field(rows, 900 + rank).view(np.int64)).to(dev)
        yv = random_field(1, 77)[0]
        sq = zd.ShardedQuotient(g, rs)
        adv, sel = [], []
        for lst in (adv, sel):
            for _ in range(qc):
                buf, view = sq.alloc_column(rows, dev)
                view.copy_(base)
                lst.append(buf)
        values = torch.zeros_like(base)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for it in range(args.iters + 2):
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            sq.run_padded(values, sel, adv, [], y=yv)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        tt = torch.tensor([best], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"op": "sharded_quotient_time", "log_rows": args.time_log_n, "cols": qc, "world": world, "ms": tt.item(),
                              "rows_per_s": isize / (tt.item() * 1e-3), "halo_rows": [sq.halo_lo, sq.halo_hi],
                              "halo_bytes_per_rank": (sq.halo_lo + sq.halo_hi) * 32 * 2 * qc}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
