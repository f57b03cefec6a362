import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field
import torch
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0); lib = zkb.lib()
lib.zkb_srs_set_precompute(1)  # slices need the SRS window table: build it up front
k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
bases = zkb.g1_fixed_base_mul(random_field(n, 2))
params = zkb.ParamsKZG(k, bases)
s = random_field(n, 1)
h = torch.from_numpy(s.view(np.int64)).pin_memory()
out = np.zeros(12, dtype=np.uint64); outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
ref = None
for slices in (1, 2, 0, 0, 3, 4, 6, 8):
    lib.zkb_msm_set_slices(slices)
    for pin in (True, False):
        ptr = ctypes.cast(h.data_ptr(), ctypes.POINTER(ctypes.c_uint64)) if pin else s.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        lib.zkb_msm_g1_srs(params.handle_g, ptr, n, outp)
        if ref is None: ref = out.copy()
        assert (out == ref).all()
        zkb.prof.enable(True); zkb.prof.reset()
        t = time.perf_counter()
        for _ in range(3): lib.zkb_msm_g1_srs(params.handle_g, ptr, n, outp)
        dt = (time.perf_counter() - t) / 3
        parts = {nm: round(zkb.prof.get(nm)[0] / 3, 2) for nm in ("msm_digits", "msm_sort", "msm_accumulate", "msm_reduce")}
        zkb.prof.enable(False)
        t = time.perf_counter()
        for _ in range(3): lib.zkb_msm_g1_srs(params.handle_g, ptr, n, outp)
        dt2 = (time.perf_counter() - t) / 3
        print("slices", slices, "pinned" if pin else "pageable", "ms", round(dt * 1e3, 2), "noprof", round(dt2 * 1e3, 2), parts, flush=True)
