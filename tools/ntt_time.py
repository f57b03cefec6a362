#!/usr/bin/env python
"""Device-resident batched NTT timings (CUDA events): python tools/ntt_time.py [--cases 22x16 24x4 26x1 ...]
Set ZKB200_LIB to A/B another build of the same ABI."""
import argparse, ctypes, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
u64p = ctypes.POINTER(ctypes.c_uint64)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="+", default=["22x16", "24x4", "26x1", "20x64", "18x64"])
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import torch
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(); sp = ctypes.c_void_p(st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for case in a.cases:
        k, cols = (int(x) for x in case.split("x"))
        N = 1 << k
        g = torch.Generator(device=dev); g.manual_seed(k)
        d = torch.randint(0, 1 << 60, (N * cols, 4), dtype=torch.int64, device=dev, generator=g)
        s = torch.empty_like(d)
        w = zkb.omega(k); wp = w.ctypes.data_as(u64p)
        run = lambda: lib.zkb_ntt_fr_dev(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(s.data_ptr()), cols, wp, k, sp)
        assert run() == 0, lib.zkb_last_error()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0.record(st); run(); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"case": case, "ntt_ms": round(best, 4), "elems_per_s": N * cols / (best * 1e-3)}), flush=True)
        del d, s

if __name__ == "__main__":
    main()
