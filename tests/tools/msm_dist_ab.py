"""MSM time by scalar distribution (uniform / witness-like / all-equal / all-zero), device-resident, with the kernel split."""
import ctypes, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import random_field
from sweep import witness_like
import torch
from oracle import coracle
coracle.build()
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0); lib = zkb.lib()
lib.zkb_srs_set_precompute(1)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(); sptr = ctypes.c_void_p(stream.cuda_stream)
out = np.zeros(12, dtype=np.uint64); outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in [int(a) for a in sys.argv[1:]] or [22, 24]:
    n = 1 << k
    dlog = random_field(n, 0xB45E)
    params = zkb.ParamsKZG(k, zkb.g1_fixed_base_mul(dlog))
    u = random_field(n, 0x5EED0000 + k)
    e = u.copy(); e[:] = u[0]
    z = np.zeros_like(u)
    for name, sc in (("U", u), ("W", witness_like(u, k)), ("E", e), ("Z", z)):
        d_s = torch.from_numpy(sc.view(np.int64)).to(dev)
        run = lambda: lib.zkb_msm_g1_srs_dev(params.handle_g, 0, ctypes.c_void_p(d_s.data_ptr()), n, outp, sptr)
        assert run() == 0
        want = coracle.g1_mul(coracle.g1_generator(), coracle.fr_inner_product(sc, dlog))
        ok = bool((out[:8] == want).all()) if name != "Z" else not out[8:].any()
        zkb.prof.enable(True); zkb.prof.reset()
        best = 1e30
        for _ in range(3):
            e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        parts = {nm: round(zkb.prof.get(nm)[0] / 3, 3) for nm in ("msm_digits", "msm_sort", "msm_accumulate", "msm_reduce")}
        zkb.prof.enable(False)
        print(json.dumps({"log_n": k, "dist": name, "ms": round(best, 3), "Mpts_per_s": round(n / best / 1e3, 1), "parity": ok, "kernels_ms": parts}), flush=True)
    params.close()
