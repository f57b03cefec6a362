#!/bin/bash
# compute-sanitizer over the small-size all-kernel pass and the bucket sort self-test (one GPU); each tool bounded by timeout
cat > /tmp/bsort_small.py <<'PY'
import ctypes, importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
zkb = importlib.import_module("zksnap-circuits-halo2_b200"); zkb.init(0)
lib = zkb.lib(); u32p = ctypes.POINTER(ctypes.c_uint32)
rng = np.random.default_rng(1)
for n, kb, tile in ((1, 1, 0), (5000, 12, 0), (5000, 13, 512), (40000, 21, 1024), (70000, 24, 8192), (33000, 17, 0)):
    keys = rng.integers(0, 1 << kb, size=n, dtype=np.uint32); k = keys.copy(); v = np.arange(n, dtype=np.uint32)
    assert lib.zkb_bucket_sort_pairs(k.ctypes.data_as(u32p), v.ctypes.data_as(u32p), n, kb, tile) == 0
    assert (k == np.sort(keys)).all() and (keys[v] == k).all()
k = np.full(30000, 77, dtype=np.uint32); v = np.arange(30000, dtype=np.uint32)
assert lib.zkb_bucket_sort_pairs(k.ctypes.data_as(u32p), v.ctypes.data_as(u32p), 30000, 18, 0) == 0 and (k == 77).all()
print("bsort ok")
PY
for tool in memcheck racecheck; do
  echo "=== $tool: bucket sort"; timeout 600 compute-sanitizer --tool $tool python /tmp/bsort_small.py 2>&1 | tail -6
done
echo "=== memcheck: all kernels"; timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py 2>&1 | tail -6
