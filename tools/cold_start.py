"""Cold-start costs of a fresh process (ms): what a prover that runs ONE proof per process pays before its first result."""
import ctypes, importlib, os, sys, time
t00 = time.perf_counter()
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
def t(label, fn):
    a = time.perf_counter(); r = fn(); print(f"{label:58s} {1e3 * (time.perf_counter() - a):9.1f} ms", flush=True); return r
lib = t("dlopen libzkb200.so", zkb.lib)
t("zkb_init (CUDA context)", lambda: zkb.init(0))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
s = random_field(n, 1)
t("first tiny NTT (module load of the NTT kernels)", lambda: zkb.best_fft(s[:8].copy(), zkb.omega(3), 3))
t("first fixed-base call, 64 scalars (builds the 64 MiB table of G)", lambda: zkb.g1_fixed_base_mul(s[:64]))
p = t(f"ParamsKZG::setup resident, k = {k}", lambda: zkb.ParamsKZG.setup(k, s[0]))
t("first commit (automatic policy: no SRS window table yet)", lambda: p.commit(s))
t("second commit", lambda: p.commit(s))
t("first commit_lagrange", lambda: p.commit_lagrange(s))
cb, tb = ctypes.c_uint32(), ctypes.c_uint64()
t("zkb_srs_precompute (forced SRS window table build)", lambda: lib.zkb_srs_precompute(p.handle_g, ctypes.byref(cb), ctypes.byref(tb)))
t("commit through the table", lambda: p.commit(s))
d = zkb.EvaluationDomain(4, k)
c = t("first lagrange_to_coeff (plan, twiddle tables, staging buffers)", lambda: d.lagrange_to_coeff(s))
t("second lagrange_to_coeff", lambda: d.lagrange_to_coeff(s))
e = t("first coeff_to_extended", lambda: d.coeff_to_extended(c))
t("second coeff_to_extended", lambda: d.coeff_to_extended(c))
print(f"{'total since interpreter start':58s} {1e3 * (time.perf_counter() - t00):9.1f} ms")
