import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0); lib = zkb.lib()
for k in (16, 20, 22, 24):
    n = 1 << k
    s = random_field(n, k)
    t = time.perf_counter(); b = zkb.g1_fixed_base_mul(s); t1 = time.perf_counter() - t
    t = time.perf_counter(); b = zkb.g1_fixed_base_mul(s); t2 = time.perf_counter() - t
    tn = None
    if k <= 22:
        t = time.perf_counter(); b2 = zkb.g1_fixed_base_mul_naive(s); tn = time.perf_counter() - t
        assert (b == b2).all()
    t = time.perf_counter(); p = zkb.ParamsKZG.setup(k, s[0]); ts = time.perf_counter() - t
    cb, tb = ctypes.c_uint32(), ctypes.c_uint64()
    t = time.perf_counter(); lib.zkb_srs_precompute(p.handle_g, ctypes.byref(cb), ctypes.byref(tb)); tt = time.perf_counter() - t
    p.close()
    print({"k": k, "fixed_base_first_s": round(t1, 4), "fixed_base_s": round(t2, 4), "naive_s": tn and round(tn, 4), "kzg_setup_resident_s": round(ts, 4),
           "srs_table_build_s": round(tt, 4), "table_c": cb.value, "table_GiB": tb.value / 2**30}, flush=True)
# g_to_lagrange (inverse G1 FFT) on a resident SRS
for k in (12, 16, 18):
    s = random_field(1, k)[0]
    p = zkb.ParamsKZG.setup(k, s)
    h = ctypes.c_uint64(0)
    t = time.perf_counter(); rc = lib.zkb_srs_g_to_lagrange(p.handle_g, k, ctypes.byref(h)); tt = time.perf_counter() - t
    assert rc == 0
    lib.zkb_srs_release(h)
    p.close()
    print({"k": k, "g_to_lagrange_s": round(tt, 4)}, flush=True)
