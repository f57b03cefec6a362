#!/usr/bin/env python
"""Time of zkb_lookup_permute_expression_pair on resident columns (range-table lookup: every input value occurs in the table):
   python tools/lookup_time.py [--log-u 18 20 22] — one JSON line per size; the result is checked against numpy's sort of the values."""
import argparse, ctypes, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import ints_to_limbs, limbs_to_int  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-u", type=int, nargs="+", default=[18, 20, 22])
    a = ap.parse_args()
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    plonk = importlib.import_module("zksnap-circuits-halo2_b200.plonk")
    zkb.init(0)
    from oracle import pyref
    for k in a.log_u:
        u = 1 << k
        rng = np.random.default_rng(k)
        table_vals = np.arange(u, dtype=np.uint64) % (1 << 16)              # a 16-bit range table, repeated
        in_vals = rng.integers(0, 1 << 16, size=u, dtype=np.uint64)
        mont = lambda v: ints_to_limbs([pyref.to_mont(int(x), pyref.FR) for x in v])
        # Montgomery form of small integers through a 65536-entry table (Python big ints only for the table)
        lut = mont(np.arange(1 << 16))
        A = zkb.Polynomial(lut[in_vals])
        S = zkb.Polynomial(lut[table_vals])
        best = 1e30
        for rep in range(3):
            t0 = time.perf_counter()
            ap_, sp_ = plonk._permute_expression_pair(A, S, u)
            best = min(best, time.perf_counter() - t0)
            if rep == 0:
                got = ap_.to_host()
                ok = bool((got == lut[np.sort(in_vals)]).all())
        print(json.dumps({"log_u": k, "ms": round(best * 1e3, 3), "sorted_input_ok": ok}), flush=True)


if __name__ == "__main__":
    main()
