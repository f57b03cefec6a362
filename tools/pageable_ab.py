#!/usr/bin/env python
"""Pageable-host-memory drop-in calls (what a Rust Vec<Fr> is): a 2^24-point commit and a 2^22 x 16 NTT batch, wall clock.
   ZKB_STAGE_THREADS=N python tools/pageable_ab.py   (the staging pools read the variable when they start)"""
import ctypes, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field  # noqa: E402
u64p = ctypes.POINTER(ctypes.c_uint64)

def main():
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    k = 24
    n = 1 << k
    bases = zkb.g1_fixed_base_mul(random_field(n, 2))
    params = zkb.ParamsKZG(k, bases)
    del bases
    lib.zkb_srs_precompute(params.handle_g, None, None)
    s = random_field(n, 1)
    out = np.zeros(12, dtype=np.uint64)
    run = lambda: lib.zkb_msm_g1_srs(params.handle_g, s.ctypes.data_as(u64p), n, out.ctypes.data_as(u64p))
    assert run() == 0
    best = 1e30
    for _ in range(5):
        t = time.perf_counter(); run(); best = min(best, time.perf_counter() - t)
    res = {"stage_threads": os.environ.get("ZKB_STAGE_THREADS", "default"), "cpus": os.cpu_count(), "msm_2^24_pageable_ms": round(best * 1e3, 2)}
    params.close()
    N, cols = 1 << 22, 16
    a = random_field(N * cols, 3)
    ptrs = (u64p * cols)(*[ctypes.cast(a.ctypes.data + i * N * 32, u64p) for i in range(cols)])
    w = zkb.omega(22)
    run2 = lambda: lib.zkb_ntt_fr_batch(ptrs, cols, w.ctypes.data_as(u64p), 22)
    assert run2() == 0
    best = 1e30
    for _ in range(3):
        t = time.perf_counter(); run2(); best = min(best, time.perf_counter() - t)
    res["ntt_2^22x16_pageable_ms"] = round(best * 1e3, 2)
    print(json.dumps(res), flush=True)

if __name__ == "__main__":
    main()
