"""zkb_ntt_fr_batch end to end (2^22 x 16 columns): page-locked vs pageable caller memory, staging threads."""
import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_field
import torch
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0); lib = zkb.lib()
k, cols = 22, 16
N = 1 << k
a = random_field(N * cols, 5)
h = torch.from_numpy(a.view(np.int64)).pin_memory()
w = zkb.omega(k); wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
u64p = ctypes.POINTER(ctypes.c_uint64)
pin = (u64p * cols)(*[ctypes.cast(h.data_ptr() + i * N * 32, u64p) for i in range(cols)])
pag = (u64p * cols)(*[ctypes.cast(a.ctypes.data + i * N * 32, u64p) for i in range(cols)])
for name, ptrs in (("pinned", pin), ("pageable", pag)):
    lib.zkb_ntt_fr_batch(ptrs, cols, wp, k)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); lib.zkb_ntt_fr_batch(ptrs, cols, wp, k); best = min(best, time.perf_counter() - t)
    print(name, "ms", round(best * 1e3, 1), "Gelem/s", round(N * cols / best / 1e9, 3), "GB/s each way", round(N * cols * 32 / best / 1e9, 1), flush=True)
print("threads", os.environ.get("ZKB_STAGE_THREADS", "default"), "cpus", os.cpu_count())
