#!/usr/bin/env python
"""Per-call wall times of the voter-shaped drop-in sequence (k = 13, 256 columns) from pageable and from page-locked host memory:
   python tools/voter_ops_time.py [--k 13 --cols 256]"""
import argparse, ctypes, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field  # noqa: E402
u64p = ctypes.POINTER(ctypes.c_uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--cols", type=int, default=256)
    a = ap.parse_args()
    import torch
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    kk, ncols = a.k, a.cols
    nn, NN = 1 << kk, 1 << (kk + 2)
    g = zkb.g1_fixed_base_mul(random_field(nn, 7))
    p = zkb.ParamsKZG(kk, g, g)
    lib.zkb_srs_precompute(p.handle_g, None, None)
    lib.zkb_srs_precompute(p.handle_g_lagrange, None, None)
    c_np = random_field(nn * ncols, 77)
    for pinned in (False, True):
        if pinned:
            hc = torch.from_numpy(c_np.view(np.int64).copy()).pin_memory(); cb = hc.data_ptr()
            he = torch.empty((NN * ncols, 4), dtype=torch.int64).pin_memory(); eb = he.data_ptr()
        else:
            hc = c_np.copy(); cb = hc.ctypes.data
            he = np.empty((NN * ncols, 4), dtype=np.uint64); eb = he.ctypes.data
        ptrs = (u64p * ncols)(*[ctypes.cast(cb + i * nn * 32, u64p) for i in range(ncols)])
        eptrs = (u64p * ncols)(*[ctypes.cast(eb + i * NN * 32, u64p) for i in range(ncols)])
        outs = np.zeros((ncols, 12), dtype=np.uint64)
        out1 = np.zeros(12, dtype=np.uint64)
        ops = {
            "commit_lagrange_batch": lambda: lib.zkb_msm_g1_srs_batch(p.handle_g_lagrange, ptrs, ncols, nn, outs.ctypes.data_as(u64p)),
            "lagrange_to_coeff_batch": lambda: lib.zkb_lagrange_to_coeff_batch(ptrs, ncols, kk),
            "coeff_to_extended_batch": lambda: lib.zkb_coeff_to_extended_batch(ptrs, eptrs, ncols, kk, kk + 2),
            "extended_to_coeff": lambda: lib.zkb_extended_to_coeff(ctypes.cast(eb, u64p), kk, kk + 2),
            "commit_x8": lambda: [lib.zkb_msm_g1_srs(p.handle_g, ctypes.cast(cb + i * nn * 32, u64p), nn, out1.ctypes.data_as(u64p)) for i in range(8)][-1],
        }
        res = {"memory": "page-locked" if pinned else "pageable"}
        for name, fn in ops.items():
            assert fn() == 0, lib.zkb_last_error()
            best = 1e30
            for _ in range(4):
                t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
            res[name] = round(best * 1e3, 3)
        res["sum"] = round(sum(v for k_, v in res.items() if k_ != "memory"), 3)
        print(json.dumps(res), flush=True)
    p.close()


if __name__ == "__main__":
    main()
