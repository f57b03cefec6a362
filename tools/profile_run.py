#!/usr/bin/env python
"""Small single-op driver for ncu / timing experiments (one GPU).

  python tools/profile_run.py msm --log-n 22 [--c 17 --chunk 128] [--reps 3]
  python tools/profile_run.py ntt --log-n 22 --cols 4 [--reps 3]
  python tools/profile_run.py c2e --log-n 20 --cols 4         (coeff_to_extended k -> k+2)
  python tools/profile_run.py graph --log-n 24 --cols 1       (quotient evaluation: halo2-base gate on `cols` advice columns,
                                                               2^log_n extended rows, rot_scale 4)
Prints one JSON line with CUDA-event timings (per-kernel timers from the library's profiler).
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["msm", "ntt", "c2e", "msmbatch", "graph"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--cols", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--dist", default="U", choices=["U", "W", "E"])
    ap.add_argument("--no-table", action="store_true")
    args = ap.parse_args()
    import torch

    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    n = 1 << args.log_n
    res = {"op": args.op, "log_n": args.log_n, "cols": args.cols}
    zkb.prof.enable(True)
    if args.op == "msm":
        lib.zkb_msm_set_params(args.c, args.chunk)
        lib.zkb_srs_set_precompute(0 if args.no_table else 1)
        s = random_field(n, 1)
        if args.dist == "E":
            s[:] = s[0]
        elif args.dist == "W":
            rng = np.random.default_rng(3)
            u = rng.random(n)
            s[u < 0.5] = 0
            small = (u >= 0.5) & (u < 0.75)
            s[small, 1:] = 0
            s[small, 0] &= np.uint64(0xFFFF)
            mid = (u >= 0.75) & (u < 0.95)
            s[mid, 2:] = 0
            s[mid, 1] &= np.uint64(0xFFFFFF)
        bases = zkb.g1_fixed_base_mul(random_field(n, 2))
        params = zkb.ParamsKZG(args.log_n, bases)
        d_s = torch.from_numpy(s.view(np.int64)).to(dev)
        out = np.zeros(12, dtype=np.uint64)
        outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        times = []
        for i in range(args.reps + 1):
            if i == 1:
                zkb.prof.reset()
            torch.cuda.synchronize()
            t = time.perf_counter()
            rc = lib.zkb_msm_g1_srs_dev(params.handle_g, 0, ctypes.c_void_p(d_s.data_ptr()), n, outp, sptr)
            assert rc == 0, lib.zkb_last_error()
            times.append(time.perf_counter() - t)
        res["ms"] = 1e3 * min(times[1:])
        cb, nw, ch = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        lib.zkb_msm_get_params(n, ctypes.byref(cb), ctypes.byref(nw), ctypes.byref(ch))
        res.update(c=cb.value, windows=nw.value, chunk=ch.value, dist=args.dist)
        tc, tb = ctypes.c_uint32(), ctypes.c_uint64()
        lib.zkb_srs_precompute(params.handle_g, ctypes.byref(tc), ctypes.byref(tb))
        res.update(table_c=tc.value, table_GiB=tb.value / 2**30)
        for name in ("msm_digits", "msm_sort", "msm_accumulate", "msm_reduce"):
            ms, k = zkb.prof.get(name)
            res[name] = ms / max(k, 1)
        res["Mpts_per_s"] = n / (res["ms"] * 1e-3) / 1e6
    elif args.op == "graph":
        # evaluate_h's custom-gate pass for halo2-base: per advice column one gate q_i (a + b c - d), a..d = rotations 0..3 of the
        # column, folded with y from the previous value — on polynomials resident in HBM
        ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
        g = ev.GraphEvaluator()
        parts = []
        for i in range(args.cols):
            a, b, c, d = (("advice", i, r) for r in range(4))
            parts.append(g.add_expression(("prod", ("fixed", i, 0), ("sum", ("sum", a, ("prod", b, c)), ("neg", d)))))
        g.add_horner(ev.ValueSource(ev.PREVIOUS), parts, ev.ValueSource(ev.Y))
        adv = [zkb.Polynomial(random_field(n, 40 + i)) for i in range(args.cols)]
        sel = [zkb.Polynomial(random_field(n, 80 + i)) for i in range(args.cols)]
        values = zkb.Polynomial(random_field(n, 7))
        y = random_field(1, 8)[0]
        for i in range(args.reps + 1):
            if i == 1:
                zkb.prof.reset()
            g.evaluate(values, fixed=sel, advice=adv, y=y, rot_scale=4)
        values.to_host()[:1]
        ms, k = zkb.prof.get("graph_evaluate")
        info = g.last_info()
        res["ms"] = ms / max(k, 1)
        res.update(info)
        res["Mrows_per_s"] = n / (res["ms"] * 1e-3) / 1e6
        res["alg_GBps"] = info["bytes_per_row"] * n / (res["ms"] * 1e-3) / 1e9
        res["hbm_frac_of_6459.6"] = res["alg_GBps"] / 6459.6
        res["modmul_per_row"] = 2 * args.cols + args.cols
    elif args.op == "msmbatch":
        lib.zkb_srs_set_precompute(0 if args.no_table else 1)
        bases = zkb.g1_fixed_base_mul(random_field(n, 2))
        params = zkb.ParamsKZG(args.log_n, bases)
        cols = [random_field(n, 10 + i) for i in range(args.cols)]
        times = []
        for i in range(args.reps + 1):
            t = time.perf_counter()
            out = params.commit_batch(cols)
            times.append(time.perf_counter() - t)
        res["ms_e2e"] = 1e3 * min(times[1:])
        res["Mpts_per_s_e2e"] = n * args.cols / min(times[1:]) / 1e6
        t = time.perf_counter()
        for c in cols[:4]:
            params.commit(c)
        res["ms_per_single_commit"] = 1e3 * (time.perf_counter() - t) / 4
    else:
        k = args.log_n
        ek = k + 2 if args.op == "c2e" else k
        a = random_field(n * args.cols, 5)
        d_a = torch.from_numpy(a.view(np.int64)).to(dev)
        N = 1 << ek
        d_o = torch.empty(N * args.cols * 4, dtype=torch.int64, device=dev)
        d_s = torch.empty_like(d_o)
        w = zkb.omega(ek)
        wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for i in range(args.reps + 1):
            e0.record(stream)
            if args.op == "ntt":
                rc = lib.zkb_ntt_fr_dev(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(d_s.data_ptr()), args.cols, wp, k, sptr)
            else:
                rc = lib.zkb_coeff_to_extended_dev(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(d_o.data_ptr()),
                                                   ctypes.c_void_p(d_s.data_ptr()), args.cols, k, ek, sptr)
            assert rc == 0, lib.zkb_last_error()
            e1.record(stream)
            torch.cuda.synchronize()
            if i > 0:
                best = min(best, e0.elapsed_time(e1))
        res["ms"] = best
        res["Melems_per_s"] = N * args.cols / (best * 1e-3) / 1e6
        alg = (64.0 * N if args.op == "ntt" else 160.0 * n) * args.cols
        res["alg_GBps"] = alg / (best * 1e-3) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
