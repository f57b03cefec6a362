#!/usr/bin/env python
"""Sharded-NTT parity and timing under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tests/tools/dist_ntt_check.py --log-n 16 20 24 26 [--time]

Parity: every rank builds the same seeded 2^log_n vector, passes its contiguous slice to ShardedNtt.best_fft_slice and
compares the returned slice with (a) the single-GPU zkb.best_fft of the whole vector (log_n <= 24) and (b) the C oracle
(log_n <= 20).  Timing (--time): device-resident zkb_dist_ntt_fr_dev on the symmetric slices, CUDA events, max over ranks,
next to the single-GPU time of the same transform.  Prints one JSON line per size on rank 0.
"""
import argparse
import ctypes
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
    ap.add_argument("--log-n", type=int, nargs="+", default=[16, 20])
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from util import random_field

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zdist = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    zkb.init(local)
    lib = zkb.lib()
    sh = zdist.ShardedNtt(max(args.log_n), device=dev)
    ok_all = True
    for k in args.log_n:
        n = 1 << k
        a = random_field(n, 4000 + k) if k <= 26 else np.tile(random_field(1 << 26, 4000 + k), (1 << (k - 26), 1))
        w = zkb.omega(k)
        off, ln = zdist.ntt_slice(k, rank, world)
        got = sh.best_fft_slice(a[off:off + ln], w, k)
        checks = {}
        if k <= 24:
            full = a.copy()
            zkb.best_fft(full, w, k)
            checks["single_gpu"] = bool((got == full[off:off + ln]).all())
        if k <= 20:
            from oracle import coracle
            coracle.build()
            checks["oracle"] = bool((got == coracle.best_fft(a, w, k)[off:off + ln]).all())
        # size-independent property (the only check above 2^24): a 2-sparse input a*delta_j1 + b*delta_j2 must give
        # out[i] = a w^(i j1) + b w^(i j2), evaluated with Python integers at sampled positions of this rank's slice
        from oracle import pyref as R
        from util import limbs_to_int
        j1, j2 = 12345 % n, n - 7
        sp = np.zeros((ln, 4), dtype=np.uint64)
        for j, src in ((j1, a[1]), (j2, a[2])):
            if off <= j < off + ln:
                sp[j - off] = src
        got_sp = sh.best_fft_slice(sp, w, k)
        wint = R.omega_for(k)
        a1, a2 = R.from_mont(limbs_to_int(a[1]), R.FR), R.from_mont(limbs_to_int(a[2]), R.FR)
        good = True
        for i in (off, off + 1, off + ln // 2 + 3, off + ln - 1):
            want = (a1 * pow(wint, i * j1, R.FR) + a2 * pow(wint, i * j2, R.FR)) % R.FR
            good &= R.from_mont(limbs_to_int(got_sp[i - off]), R.FR) == want
        checks["two_sparse_definition"] = bool(good)
        ok = all(checks.values())
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
        ok_all &= ok
        line = {"op": "dist_best_fft", "log_n": k, "world": world, "parity": ok, "checks": sorted(checks)}
        if args.time:
            stream = torch.cuda.current_stream()
            sp = ctypes.c_void_p(stream.cuda_stream)
            wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            zkb.prof.enable(True)
            for it in range(3 + args.iters):
                if it == 3:
                    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                    zkb.prof.reset()
                    e0.record(stream)
                rc = lib.zkb_dist_ntt_fr_dev(None, None, wp, k, sp)
                assert rc == 0, lib.zkb_last_error()
            e1.record(stream)
            assert lib.zkb_dist_status(sp) == 0, lib.zkb_last_error()
            ms = e0.elapsed_time(e1) / args.iters
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            parts = {nm: zkb.prof.get(nm)[0] / args.iters for nm in ("dist_ntt_pass0", "dist_ntt_middle", "dist_ntt_final", "dist_barrier")}
            zkb.prof.enable(False)
            line.update({"ms": float(tt.item()), "elems_per_s": n / (float(tt.item()) * 1e-3), "rank0_ms": parts,
                         "nvlink_bytes_per_rank": int(3 * (world - 1) / world * ln * 32)})
            # what an NCCL-based exchange would cost on top of the local passes: the sharded transform moves each slice three
            # times (gather for pass 0, scatter of pass 0, scatter of the last pass); time ONE all_to_all_single of the slice
            buf_in = torch.empty(ln * 4, dtype=torch.int64, device=dev)
            buf_out = torch.empty_like(buf_in)
            for it in range(2 + args.iters):
                if it == 2:
                    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                    e0.record(stream)
                dist.all_to_all_single(buf_out, buf_in)
            e1.record(stream)
            torch.cuda.synchronize()
            a2a = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev, dtype=torch.float64)
            dist.all_reduce(a2a, op=dist.ReduceOp.MAX)
            line["nccl_all_to_all_single_ms"] = float(a2a.item())
            line["nccl_three_exchanges_ms"] = 3 * float(a2a.item())
            del buf_in, buf_out
            if k <= 26 and rank == 0:  # single-GPU time of the same transform for comparison
                d = torch.empty(n * 4, dtype=torch.int64, device=dev)
                s = torch.empty_like(d)
                for it in range(2 + args.iters):
                    if it == 2:
                        torch.cuda.synchronize()
                        e0.record(stream)
                    lib.zkb_ntt_fr_dev(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(s.data_ptr()), 1, wp, k, sp)
                e1.record(stream)
                torch.cuda.synchronize()
                line["single_gpu_ms"] = e0.elapsed_time(e1) / args.iters
                del d, s
            dist.barrier()
        if rank == 0:
            print(json.dumps(line), flush=True)
    sh.close()
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
