#!/usr/bin/env python
"""Randomised differential stress run on one GPU (not part of the test suite: minutes, not seconds).
   python tests/tools/stress.py [--seconds 120] [--seed 1]
Random sizes / ops / column counts / scalar distributions / tuning overrides, every result compared with the C oracle."""
import argparse
import ctypes
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import coracle  # noqa: E402
from util import random_field  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    coracle.build()
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(0)
    lib = zkb.lib()
    rng = np.random.default_rng(args.seed)
    nmax = 1 << 16
    dlog = random_field(nmax, 99)
    bases_all = zkb.g1_fixed_base_mul(dlog)
    t_end = time.time() + args.seconds
    counts = {}
    it = 0
    while time.time() < t_end:
        it += 1
        kind = rng.choice(["ntt", "domain", "msm", "commit", "batch", "poly", "graph"])
        counts[kind] = counts.get(kind, 0) + 1
        seed = int(rng.integers(1, 1 << 30))
        if kind == "ntt":
            k = int(rng.integers(1, 19))
            a = random_field(1 << k, seed)
            w = coracle.fr_omega(k) if rng.random() < 0.5 else coracle.fr_inv(coracle.fr_omega(k))
            got = a.copy()
            zkb.best_fft(got, w, k)
            assert (got == coracle.best_fft(a, w, k)).all(), ("ntt", k, seed)
        elif kind == "domain":
            k = int(rng.integers(1, 15))
            j = int(rng.integers(2, 6))
            d = zkb.EvaluationDomain(j, k)
            ncols = int(rng.integers(1, 7))
            lib.zkb_pipeline_set(int(rng.integers(1, 4)), int(rng.choice([0, 1 << 16, 1 << 20])))
            cols = [random_field(1 << k, seed + i) for i in range(ncols)]
            for got, c in zip(d.coeff_to_extended_batch(cols), cols):
                assert (got == coracle.coeff_to_extended(c, k, d.extended_k)).all(), ("c2e", j, k, seed)
            for got, c in zip(d.lagrange_to_coeff_batch(cols), cols):
                assert (got == coracle.lagrange_to_coeff(c, k)).all(), ("l2c", k, seed)
            ext = coracle.coeff_to_extended(cols[0], k, d.extended_k)
            assert (d.extended_to_coeff(ext) == coracle.extended_to_coeff(ext, k, d.extended_k)[: d.n * d.quotient_poly_degree]).all()
            lib.zkb_pipeline_set(0, 0)
        elif kind in ("msm", "commit", "batch"):
            n = int(rng.integers(1, nmax)) if rng.random() < 0.7 else int(rng.integers(1, 200))
            s = random_field(n, seed)
            mode = rng.choice(["U", "sparse", "small", "equal"])
            if mode == "sparse":
                s[rng.random(n) < 0.8] = 0
            elif mode == "small":
                s[:, 1:] = 0
                s = coracle.fr_to_mont(s)
            elif mode == "equal":
                s[:] = s[0]
            b = bases_all[:n]
            if kind == "msm":
                lib.zkb_msm_set_params(int(rng.choice([0, 0, 5, 9, 13, 16])), int(rng.choice([0, 0, 7, 32])))
                assert (zkb.best_multiexp(s, b) == coracle.best_multiexp(s, b)).all(), ("msm", n, mode, seed)
                lib.zkb_msm_set_params(0, 0)
            else:
                k = max(6, (n - 1).bit_length())
                gp = np.zeros((1 << k, 8), dtype=np.uint64)
                gp[:n] = b
                lib.zkb_srs_set_precompute(int(rng.random() < 0.8))
                params = zkb.ParamsKZG(k, gp)
                if kind == "commit":
                    lib.zkb_msm_set_slices(int(rng.choice([0, 1, 2, 5])))
                    assert (params.commit(s) == coracle.best_multiexp(s, b)).all(), ("commit", n, mode, seed)
                    lib.zkb_msm_set_slices(0)
                else:
                    nc = int(rng.integers(1, 12))
                    cols = [s] + [random_field(n, seed + 1 + i) for i in range(nc - 1)]
                    got = params.commit_batch(cols)
                    for i, c in enumerate(cols):
                        assert (got[i] == coracle.best_multiexp(c, b)).all(), ("batch", n, nc, i, seed)
                params.close()
                lib.zkb_srs_set_precompute(1)
        elif kind == "graph":
            # quotient evaluation: a random raw program (re-written intermediates, dead values, Horner steps) or a random gate
            # set, on a random power-of-two domain, through zkb_graph_evaluate vs the oracle's direct evaluation
            import graph_cases as GC
            isize = 1 << int(rng.integers(0, 15))
            if rng.random() < 0.5:
                g, rs = GC.random_program(seed, max_len=int(rng.integers(2, 200)))
                nfix, nadv, nins, chal = 1, 2, 0, None
            else:
                c = GC.random_case(seed, 1, 1, ngates=int(rng.integers(1, 12)), depth=int(rng.integers(2, 7)))
                g, rs, nfix, nadv, nins = c["graph"], int(rng.choice([1, 2, 4, 8])), 2, 3, 1
                chal = random_field(2, seed + 50)
            cols = [[random_field(isize, seed + 10 * q + i) for i in range(m)] for q, m in enumerate((nfix, nadv, nins))]
            sc = random_field(4, seed + 60)
            prev = random_field(isize, seed + 61)
            want = coracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, cols[0], cols[1], cols[2],
                                          chal, sc[0], sc[1], sc[2], sc[3], rs, prev)
            handles = [[zkb.Polynomial(a) for a in grp] for grp in cols]
            values = zkb.Polynomial(prev)
            g.evaluate(values, *handles, challenges=chal, beta=sc[0], gamma=sc[1], theta=sc[2], y=sc[3], rot_scale=rs)
            assert (values.to_host() == want).all(), ("graph", isize, seed)
            for p_ in [values] + sum(handles, []):
                p_.free()
        else:
            n = int(rng.integers(1, 50000))
            a = random_field(n, seed)
            x = random_field(1, seed + 1)[0]
            assert (zkb.eval_polynomial(a, x) == coracle.fr_eval_polynomial(a, x)).all(), ("eval", n, seed)
            if n >= 2:
                assert (zkb.kate_division(a, x) == coracle.fr_kate_division(a, x)).all(), ("kate", n, seed)
            a[rng.random(n) < 0.1] = 0
            assert (zkb.batch_invert(a) == coracle.fr_batch_invert(a)).all(), ("inv", n, seed)
    print("stress ok:", it, "cases", counts, "launches", zkb.launch_count())


if __name__ == "__main__":
    main()
