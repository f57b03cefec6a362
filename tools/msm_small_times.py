#!/usr/bin/env python
"""Single-commit latency by size (device-resident scalars, SRS window table on / off), with the per-phase CUDA-event times:
   python tools/msm_small_times.py [--log-n 13 14 ...] — prints one JSON line per (size, mode), parity-checked by known dlog."""
import argparse, ctypes, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field  # noqa: E402
u64p = ctypes.POINTER(ctypes.c_uint64)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, nargs="+", default=[13, 14, 15, 16, 17, 18, 19, 20, 22])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=0, help="level-0 chunk override (0 = wave-fitted)")
    ap.add_argument("--table-only", action="store_true")
    a = ap.parse_args()
    import torch
    from oracle import coracle
    coracle.build()
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    lib.zkb_msm_set_params(0, a.chunk)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(); sp = ctypes.c_void_p(st.cuda_stream)
    kmax = max(a.log_n)
    dl = random_field(1 << kmax, 0xB45E)
    bases = zkb.g1_fixed_base_mul(dl)
    out = np.zeros(12, dtype=np.uint64); outp = out.ctypes.data_as(u64p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in a.log_n:
        n = 1 << k
        s = random_field(n, 100 + k)
        d_s = torch.from_numpy(s.view(np.int64)).to(dev)
        want = coracle.g1_mul(coracle.g1_generator(), coracle.fr_inner_product(s, dl[:n]))
        for table in ((1,) if a.table_only else (1, 0)):
            lib.zkb_srs_set_precompute(table)
            h = ctypes.c_uint64(0)
            assert lib.zkb_srs_register(bases.ctypes.data_as(u64p), n, ctypes.byref(h)) == 0
            run = lambda: lib.zkb_msm_g1_srs_dev(h, 0, ctypes.c_void_p(d_s.data_ptr()), n, outp, sp)
            assert run() == 0, lib.zkb_last_error()
            ok = bool((out[:8] == want).all())
            run()
            zkb.prof.enable(True); zkb.prof.reset()
            torch.cuda.synchronize()
            best = 1e30
            import time
            wall = 1e30
            for _ in range(a.reps):
                t0 = time.perf_counter()
                e0.record(st); run(); e1.record(st); torch.cuda.synchronize()
                wall = min(wall, time.perf_counter() - t0)
                best = min(best, e0.elapsed_time(e1))
            parts = {}
            for name in ("msm_digits", "msm_sort", "msm_accumulate", "msm_reduce"):
                t_ms, cnt = zkb.prof.get(name)
                parts[name[4:]] = round(t_ms / max(cnt, 1), 4)
            zkb.prof.enable(False)
            cb, tb = ctypes.c_uint32(), ctypes.c_uint64()
            lib.zkb_srs_table_info(h, ctypes.byref(cb), ctypes.byref(tb), None)
            print(json.dumps({"log_n": k, "table": bool(cb.value), "c": cb.value, "ms": round(best, 4), "wall_ms": round(wall * 1e3, 4), "parity": ok, "phases_ms": parts}), flush=True)
            lib.zkb_srs_release(h)
    lib.zkb_srs_set_precompute(2)

if __name__ == "__main__":
    main()
