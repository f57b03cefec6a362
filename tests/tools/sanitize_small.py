#!/usr/bin/env python
"""Small-size pass over every kernel family for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py
Every result is still compared with the oracle, so a sanitizer-clean run is also a parity run."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import coracle  # noqa: E402
from util import random_field  # noqa: E402

coracle.build()
zkb = importlib.import_module("zksnap-circuits-halo2_b200")
zkb.init(0)
zkb.lib().zkb_srs_set_precompute(1)  # eager SRS window table: the table path is part of the pass
for k in (3, 9, 10, 11, 13):          # 1-pass, 2-pass geometries
    a = random_field(1 << k, k)
    w = zkb.omega(k)
    got = a.copy()
    zkb.best_fft(got, w, k)
    assert (got == coracle.best_fft(a, w, k)).all(), k
d = zkb.EvaluationDomain(4, 10)
a = random_field(1 << 10, 1)
ext = d.coeff_to_extended(a)
assert (ext == coracle.coeff_to_extended(a, 10, 12)).all()
assert (d.extended_to_coeff(ext)[:1 << 10] == a).all()
cols = [random_field(1 << 10, 20 + i) for i in range(5)]
for got, c in zip(d.lagrange_to_coeff_batch(cols), cols):
    assert (got == coracle.lagrange_to_coeff(c, 10)).all()
n = 700
dl = random_field(n, 2)
bases = zkb.g1_fixed_base_mul(dl)
assert (bases[:50] == coracle.g1_fixed_base_mul(dl[:50])).all()
s = random_field(n, 3)
s[5] = 0
assert (zkb.best_multiexp(s, bases) == coracle.best_multiexp(s, bases)).all()          # no table, W bucket sets
gp = np.zeros((1 << 10, 8), dtype=np.uint64)
gp[:n] = bases
params = zkb.ParamsKZG(10, gp)
assert (params.commit(s) == coracle.best_multiexp(s, bases)).all()                      # window table
got = params.commit_batch([random_field(n, 40 + i) for i in range(9)])                  # batch + device finalisation
assert (got[0] == coracle.best_multiexp(random_field(n, 40), bases)).all()
zkb.lib().zkb_msm_set_slices(3)
assert (params.commit(s) == coracle.best_multiexp(s, bases)).all()                      # slices + merge kernel
zkb.lib().zkb_msm_set_slices(0)
params.close()
sv = random_field(1, 9)[0]
p2 = zkb.ParamsKZG.setup(5, sv)
g, gl = coracle.kzg_setup(5, sv)
assert (p2.get_g() == g).all() and (p2.get_g_lagrange() == gl).all()
p2.close()
jac = np.zeros((20, 12), dtype=np.uint64)
jac[:, :8] = bases[:20]
jac[:, 8:] = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f], dtype=np.uint64)
jac[3, 8:] = 0
want = bases[:20].copy()
want[3] = 0
assert (zkb.batch_normalize(jac) == want).all()
zkb.shutdown()
print("sanitize_small ok, launches:", zkb.launch_count())
