"""Single-column host-buffer domain ops at 2^22: page-locked vs pageable caller memory (ms per call, best of 5)."""
import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field
import torch
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0); lib = zkb.lib()
u64p = ctypes.POINTER(ctypes.c_uint64)
k, ek = 22, 24
a = random_field(1 << k, 5)
ext = np.zeros((1 << ek, 4), dtype=np.uint64)
ha = torch.from_numpy(a.view(np.int64).copy()).pin_memory()
hext = torch.empty((1 << ek) * 4, dtype=torch.int64).pin_memory()
def best(fn, reps=5):
    fn(); b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t)
    return round(b * 1e3, 2)
pa, pe = a.ctypes.data_as(u64p), ext.ctypes.data_as(u64p)
qa, qe = ctypes.cast(ha.data_ptr(), u64p), ctypes.cast(hext.data_ptr(), u64p)
print("lagrange_to_coeff 2^22      pageable", best(lambda: lib.zkb_lagrange_to_coeff(pa, k)), " pinned", best(lambda: lib.zkb_lagrange_to_coeff(qa, k)))
print("coeff_to_extended 2^22->24  pageable", best(lambda: lib.zkb_coeff_to_extended(pa, pe, k, ek)), " pinned", best(lambda: lib.zkb_coeff_to_extended(qa, qe, k, ek)))
print("extended_to_coeff 2^24      pageable", best(lambda: lib.zkb_extended_to_coeff(pe, k, ek)), " pinned", best(lambda: lib.zkb_extended_to_coeff(qe, k, ek)))
