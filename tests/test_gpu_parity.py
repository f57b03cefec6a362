"""GPU parity: the CUDA path, called through the C ABI (include/zkb200.h via ctypes), against the CPU oracle.

Bar: bit-exact (integer/field arithmetic).  Small sizes compare every output limb with oracle/zkb_oracle.c;
BASELINE.json's full sizes use size-independent properties (inverse round trip, known-discrete-log MSM,
shard-split invariance).  Nothing here reads /root/reference.
"""
import ctypes
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyref as R
from util import FQ_LIMBS, int_to_limbs, ints_to_limbs, limbs_to_int, random_field

pytestmark = pytest.mark.gpu

zkb = importlib.import_module("zksnap-circuits-halo2_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hotpath_kats.json")))


@pytest.fixture(scope="module", autouse=True)
def _init():
    zkb.init()
    yield


def mont(xs):
    return ints_to_limbs([R.to_mont(x, R.FR) for x in xs])


def unmont(a):
    return [R.from_mont(limbs_to_int(r), R.FR) for r in np.asarray(a).reshape(-1, 4)]


def aff(P):
    return np.array(R.g1_affine_encode(P), dtype=np.uint64)


# ---- golden vectors ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLD["best_fft"], ids=lambda c: f"k{c['k']}")
def test_best_fft_golden(case):
    a = mont([int(x, 16) for x in case["in"]])
    zkb.best_fft(a, mont([int(case["omega"], 16)])[0], case["k"])
    assert unmont(a) == [int(x, 16) for x in case["out"]]


def test_best_fft_survey_kat_limbs():
    a = mont(range(1, 9))
    zkb.best_fft(a, zkb.omega(3), 3)
    assert [int(x) for x in a[1]] == [0x1069F4287460CB5F, 0xEA22DD8C9B017FC5, 0xC9CDFB2B2395711E, 0x2758DB28A5C09FDD]


@pytest.mark.parametrize("case", GOLD["lagrange_to_coeff"], ids=lambda c: f"k{c['k']}")
def test_lagrange_to_coeff_golden(case):
    d = zkb.EvaluationDomain(4, case["k"])
    got = d.lagrange_to_coeff(mont([int(x, 16) for x in case["in"]]))
    assert unmont(got) == [int(x, 16) for x in case["out"]]


@pytest.mark.parametrize("case", GOLD["coeff_to_extended"], ids=lambda c: f"j{c['j']}k{c['k']}")
def test_coeff_to_extended_golden(case):
    d = zkb.EvaluationDomain(case["j"], case["k"])
    assert d.extended_k == case["extended_k"]
    a = [int(x, 16) for x in case["in"]]
    ext = d.coeff_to_extended(mont(a))
    assert unmont(ext) == [int(x, 16) for x in case["out"]]
    back = unmont(d.extended_to_coeff(ext))
    want = (a + [0] * len(back))[: d.n * d.quotient_poly_degree]
    assert back == want


@pytest.mark.parametrize("case", GOLD["msm"], ids=lambda c: c["name"])
def test_msm_golden(case):
    s = mont([int(x, 16) for x in case["scalars"]])
    b = np.stack([aff(tuple(int(c, 16) for c in p) if p else None) for p in case["bases"]])
    got = zkb.best_multiexp(s, b)
    assert R.g1_jacobian_decode([int(x) for x in got]) == tuple(int(x, 16) for x in case["out"])
    assert limbs_to_int(got[8:]) == R.FQ_R


@pytest.mark.parametrize("case", GOLD["g1_mul"], ids=lambda c: c["s"][:12])
def test_fixed_base_mul_golden(case):
    got = zkb.g1_fixed_base_mul(mont([int(case["s"], 16)]))[0]
    want = tuple(int(x, 16) for x in case["out"]) if case["out"] else None
    assert R.g1_affine_decode([int(x) for x in got]) == want


# ---- differential vs the C oracle ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", list(range(1, 19)) + [20])
def test_ntt_vs_oracle(oracle, k):
    a = random_field(1 << k, 100 + k)
    w = oracle.fr_omega(k)
    got = a.copy()
    zkb.best_fft(got, w, k)
    assert (got == oracle.best_fft(a, w, k)).all()
    wi = oracle.fr_inv(w)
    zkb.best_fft(got, wi, k)  # omega_inv as the caller would pass for an ifft
    assert (got == oracle.best_fft(oracle.best_fft(a, w, k), wi, k)).all()


@pytest.mark.parametrize("j,k", [(4, 1), (4, 5), (4, 9), (4, 10), (4, 13), (4, 15), (3, 12), (5, 11), (2, 8)])
def test_domain_vs_oracle(oracle, j, k):
    d = zkb.EvaluationDomain(j, k)
    a = random_field(1 << k, 7 * k + j)
    assert (d.lagrange_to_coeff(a) == oracle.lagrange_to_coeff(a, k)).all()
    assert (d.coeff_to_lagrange(a) == oracle.coeff_to_lagrange(a, k)).all()
    ext = d.coeff_to_extended(a)
    assert (ext == oracle.coeff_to_extended(a, k, d.extended_k)).all()
    back = d.extended_to_coeff(ext)
    assert (back == oracle.extended_to_coeff(ext, k, d.extended_k)[: d.n * d.quotient_poly_degree]).all()


def test_ntt_batch_vs_oracle(oracle):
    k = 13
    d = zkb.EvaluationDomain(4, k)
    cols = [random_field(1 << k, 300 + i) for i in range(5)]
    for got, a in zip(d.lagrange_to_coeff_batch(cols), cols):
        assert (got == oracle.lagrange_to_coeff(a, k)).all()
    for got, a in zip(d.coeff_to_extended_batch(cols), cols):
        assert (got == oracle.coeff_to_extended(a, k, d.extended_k)).all()


def _bases_known_dlog(n, seed):
    b = random_field(n, seed)
    return b, zkb.g1_fixed_base_mul(b)


def test_fixed_base_mul_vs_oracle(oracle):
    s = random_field(200, 4)
    s[0] = 0
    assert (zkb.g1_fixed_base_mul(s) == oracle.g1_fixed_base_mul(s)).all()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 1000, 4097, 1 << 13, (1 << 15) + 5])
def test_msm_vs_oracle(oracle, n):
    s = random_field(n, n)
    _, bases = _bases_known_dlog(n, 5000 + n)
    got = zkb.best_multiexp(s, bases)
    assert (got == oracle.best_multiexp(s, bases)).all()


@pytest.mark.parametrize("c,chunk", [(2, 8), (3, 16), (4, 8), (7, 3), (11, 64), (16, 256)])  # c = 2 / 3: 128 / 85 window sums in the host fold
def test_msm_window_and_chunk_overrides(oracle, c, chunk):
    n = 3000
    s = random_field(n, 42)
    _, bases = _bases_known_dlog(n, 43)
    want = oracle.best_multiexp(s, bases)
    zkb.lib().zkb_msm_set_params(c, chunk)
    try:
        assert (zkb.best_multiexp(s, bases) == want).all()
    finally:
        zkb.lib().zkb_msm_set_params(0, 0)


def test_msm_edge_cases(oracle):
    n = 600
    _, b = _bases_known_dlog(n, 5)
    ident = zkb.best_multiexp(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64))
    assert not ident[8:].any() and limbs_to_int(ident[4:8]) == R.FQ_R
    out = zkb.best_multiexp(np.zeros((10, 4), np.uint64), b[:10])
    assert not out[8:].any()
    s = random_field(n, 6)
    s[:] = s[0]  # all-equal scalars: worst bucket collision (distribution E)
    assert (zkb.best_multiexp(s, b) == oracle.best_multiexp(s, b)).all()
    b2 = np.tile(b[0], (n, 1))  # repeated bases: doubling branch
    assert (zkb.best_multiexp(s, b2) == oracle.best_multiexp(s, b2)).all()
    s = random_field(n, 7)
    s[::2] = 0
    s[7] = mont([R.FR - 1])[0]
    s[9] = mont([1])[0]
    b3 = b.copy()
    b3[5] = 0  # identity base
    y = limbs_to_int(b3[10][4:])
    b3[11][:4] = b3[10][:4]
    b3[11][4:] = int_to_limbs(R.FQ - y)  # negated base with the same scalar
    s[11] = s[10]
    assert (zkb.best_multiexp(s, b3) == oracle.best_multiexp(s, b3)).all()
    sc = np.stack([s[10], s[10]])
    out = zkb.best_multiexp(sc, b3[10:12])  # everything cancels
    assert not out[8:].any()


def test_msm_witness_like_distribution(oracle):
    n = 20000
    rng = np.random.default_rng(9)
    u = rng.random(n)
    vals = []
    for i in range(n):
        if u[i] < 0.5: vals.append(0)
        elif u[i] < 0.75: vals.append(int(rng.integers(0, 1 << 16)))
        elif u[i] < 0.95: vals.append(int.from_bytes(rng.bytes(11), "little"))
        else: vals.append(int.from_bytes(rng.bytes(31), "little") % R.FR)
    s = mont(vals)
    _, b = _bases_known_dlog(n, 11)
    assert (zkb.best_multiexp(s, b) == oracle.best_multiexp(s, b)).all()


@pytest.mark.parametrize("precompute", [1, 0])
def test_params_kzg_commit_and_range_split(oracle, precompute):
    zkb.lib().zkb_srs_set_precompute(precompute)
    k = 12
    n = 1 << k
    _, g = _bases_known_dlog(n, 21)
    _, gl = _bases_known_dlog(n, 22)
    params = zkb.ParamsKZG(k, g, gl)
    poly = random_field(n, 23)
    c1 = params.commit(poly)
    assert (c1 == oracle.best_multiexp(poly, g)).all()
    assert (params.commit_lagrange(poly) == oracle.best_multiexp(poly, gl)).all()
    # shorter polynomial commits against a prefix of the SRS
    assert (params.commit(poly[:1000]) == oracle.best_multiexp(poly[:1000], g[:1000])).all()
    # batch
    polys = [random_field(n, 30 + i) for i in range(4)]
    got = params.commit_batch(polys)
    for i, p in enumerate(polys):
        assert (got[i] == oracle.best_multiexp(p, g)).all()
    # point-range shards folded on the host == single MSM (multi-GPU invariance)
    parts = [params.commit_range(o, poly[o:o + n // 4]) for o in range(0, n, n // 4)]
    assert (zkb.g1_sum(np.stack(parts)) == c1).all()
    import ctypes
    cb, tb = ctypes.c_uint32(), ctypes.c_uint64()
    zkb.lib().zkb_srs_precompute(params.handle_g, ctypes.byref(cb), ctypes.byref(tb))
    assert (cb.value > 0) == bool(precompute)
    params.close()
    zkb.lib().zkb_srs_set_precompute(1)


@pytest.mark.parametrize("n", [64, 100, 1 << 10, 5000, 1 << 14])
def test_srs_table_path_vs_oracle(oracle, n):
    """Commit through the SRS window table (one bucket set for all windows) vs the oracle, incl. edge inputs."""
    zkb.lib().zkb_srs_set_precompute(1)
    _, g = _bases_known_dlog(n, 70 + n)
    g[3] = 0                       # identity base
    g[5] = g[4]                    # repeated base
    s = random_field(n, 71 + n)
    s[0] = 0
    s[1] = mont([1])[0]
    s[2] = mont([R.FR - 1])[0]
    s[5] = s[4]
    k = max(6, (n - 1).bit_length())
    gp = np.zeros((1 << k, 8), dtype=np.uint64)
    gp[:n] = g
    params = zkb.ParamsKZG(k, gp)
    assert (params.commit(s) == oracle.best_multiexp(s, g)).all()
    assert (params.commit(s[: n // 2]) == oracle.best_multiexp(s[: n // 2], g[: n // 2])).all()
    off = n // 3
    assert (params.commit_range(off, s[off:]) == oracle.best_multiexp(s[off:], g[off:])).all()
    params.close()


@pytest.mark.parametrize("k,ncols,precompute", [(10, 20, 1), (10, 20, 0), (13, 9, 1), (6, 3, 1), (12, 40, 1)])
def test_commit_batch_one_pass(oracle, k, ncols, precompute):
    """zkb_msm_g1_srs_batch: all columns in one digit/sort/accumulate pass, device finalisation for >= 8 columns."""
    zkb.lib().zkb_srs_set_precompute(precompute)
    n = 1 << k
    _, g = _bases_known_dlog(n, 90 + k)
    params = zkb.ParamsKZG(k, g)
    cols = [random_field(n, 900 + i) for i in range(ncols)]
    cols[0][:] = 0                 # an all-zero column commits to the identity
    cols[1][:] = cols[1][0]        # all-equal scalars
    got = params.commit_batch(cols)
    for i, p in enumerate(cols):
        assert (got[i] == oracle.best_multiexp(p, g)).all(), i
    params.close()
    zkb.lib().zkb_srs_set_precompute(1)


def test_commit_batch_wider_than_one_key_group(oracle):
    """Short columns against a long SRS: with the window table every column owns 2^(c-1) buckets for the SRS's c, so a batch soon
    exceeds the 24 key bits of one sort pass (the library's own bucket sort) and is cut into groups — every column still gets its own
    commitment."""
    K, k = 20, 10
    n = 1 << k
    _, g = _bases_known_dlog(1 << K, 77)
    params = zkb.ParamsKZG(K, g)
    wb, tb = ctypes.c_uint32(0), ctypes.c_uint64(0)
    zkb.lib().zkb_srs_precompute(params.handle_g, ctypes.byref(wb), ctypes.byref(tb))
    assert wb.value >= 18
    ncols = ((1 << 24) >> (wb.value - 1)) + 5              # one full group and a short second one
    base_cols = [random_field(n, 1200 + i) for i in range(5)]
    cols = [base_cols[i % 5] for i in range(ncols)]
    got = params.commit_batch(cols)
    want = [oracle.best_multiexp(p, g[:n]) for p in base_cols]
    for i in range(ncols):
        assert (got[i] == want[i % 5]).all(), i
    params.close()


# ---- BASELINE.json full sizes: size-independent properties -------------------------------------------------------------------
@pytest.mark.parametrize("k", [22, 24])
def test_ntt_full_size_roundtrip_and_linearity(oracle, k):
    n = 1 << k
    a = random_field(n, k)
    w = zkb.omega(k)
    wi = oracle.fr_inv(w)
    f = a.copy()
    zkb.best_fft(f, w, k)
    # spot values against the definition on a sparse probe: NTT(delta_j)[i] = omega^(i*j); by linearity check a
    # random 2-sparse input exactly
    sp = np.zeros((n, 4), dtype=np.uint64)
    j1, j2 = 12345 % n, (n - 7)
    sp[j1] = a[1]
    sp[j2] = a[2]
    zkb.best_fft(sp, w, k)
    wint = R.omega_for(k)
    a1, a2 = R.from_mont(limbs_to_int(a[1]), R.FR), R.from_mont(limbs_to_int(a[2]), R.FR)
    for i in (0, 1, 2, n // 2 + 3, n - 1):
        want = (a1 * pow(wint, i * j1, R.FR) + a2 * pow(wint, i * j2, R.FR)) % R.FR
        assert R.from_mont(limbs_to_int(sp[i]), R.FR) == want
    # inverse round trip: iNTT(NTT(a)) * n^-1 == a, through lagrange_to_coeff (which uses omega_inv and 1/n)
    d = zkb.EvaluationDomain(4, k)
    back = d.lagrange_to_coeff(f)
    assert (back == a).all()
    del wi


def test_coeff_to_extended_full_size_roundtrip():
    k = 22
    d = zkb.EvaluationDomain(4, k)
    assert d.extended_k == 24
    a = random_field(1 << k, 99)
    ext = d.coeff_to_extended(a)
    back = d.extended_to_coeff(ext)
    assert back.shape[0] == 3 << k
    assert (back[: 1 << k] == a).all() and not back[1 << k:].any()
    # the coset evaluations at i = 0 equal p(zeta): Horner over a strided sample is too slow in Python, so check
    # the first evaluation of a low-degree truncation instead
    lo = np.zeros_like(a)
    lo[:8] = a[:8]
    e0 = d.coeff_to_extended(lo)[0]
    coeffs = [R.from_mont(limbs_to_int(x), R.FR) for x in a[:8]]
    want = sum(c * pow(R.FR_ZETA, i, R.FR) for i, c in enumerate(coeffs)) % R.FR
    assert R.from_mont(limbs_to_int(e0), R.FR) == want


def test_ntt_2_26_single_gpu():
    """best_fft at the top of BASELINE.json's sweep (2^26, 2 GiB) on one GPU: exact values of a 2-sparse input against the
    definition at sampled outputs, and NTT -> lagrange_to_coeff round trip of a random vector."""
    k = 26
    n = 1 << k
    w = zkb.omega(k)
    a = random_field(n, 2600)
    sp = np.zeros((n, 4), dtype=np.uint64)
    j1, j2 = 0x2345677 % n, n - 11
    sp[j1] = a[1]
    sp[j2] = a[2]
    zkb.best_fft(sp, w, k)
    wint = R.omega_for(k)
    a1, a2 = R.from_mont(limbs_to_int(a[1]), R.FR), R.from_mont(limbs_to_int(a[2]), R.FR)
    rng = np.random.default_rng(26)
    for i in [0, 1, 2, n // 2, n // 2 + 3, n - 1] + [int(x) for x in rng.integers(0, n, 26)]:
        want = (a1 * pow(wint, i * j1, R.FR) + a2 * pow(wint, i * j2, R.FR)) % R.FR
        assert R.from_mont(limbs_to_int(sp[i]), R.FR) == want, i
    del sp
    f = a.copy()
    zkb.best_fft(f, w, k)
    assert not (f[:64] == a[:64]).all()
    back = zkb.EvaluationDomain(2, k).lagrange_to_coeff(f)
    assert (back == a).all()


def test_coeff_to_extended_full_size_vs_horner(oracle):
    """coeff_to_extended 2^22 -> 2^24 (the wrapper circuit's coset NTT) against Horner evaluation of the 2^22 coefficients at
    zeta * omega_ext^i for indices spread over the whole extended domain (C oracle's eval_polynomial = upstream's
    arithmetic::eval_polynomial restated)."""
    k = 22
    d = zkb.EvaluationDomain(4, k)
    a = random_field(1 << k, 2224)
    ext = d.coeff_to_extended(a)
    N = 1 << d.extended_k
    wext = R.omega_for(d.extended_k)
    rng = np.random.default_rng(2224)
    idx = [0, 1, 2, 3, N // 4, N // 2 - 1, N // 2, N - 1] + [int(x) for x in rng.integers(0, N, 16)]
    for i in idx:
        x = R.FR_ZETA * pow(wext, i, R.FR) % R.FR
        want = oracle.fr_eval_polynomial(a, mont([x])[0])
        assert (ext[i] == want).all(), i


@pytest.mark.parametrize("dist", ["W", "E"])
@pytest.mark.parametrize("precompute", [1, 0])
def test_msm_2_22_witness_like_and_equal_scalars(oracle, dist, precompute):
    """BASELINE sweep distributions at the wrapper size, known-dlog check: W = witness-like (50 % zero, 25 % < 2^16, 20 % < 2^88,
    5 % uniform), E = all-equal scalar (every point in the same bucket of every window)."""
    k = 22
    n = 1 << k
    rng = np.random.default_rng(2222)
    if dist == "E":
        s = np.tile(random_field(1, 5)[0], (n, 1))
    else:
        s = random_field(n, 2223)                       # Montgomery residues of uniform values
        u = rng.random(n)
        small = mont([int(x) for x in rng.integers(0, 1 << 16, 4096)])
        mid = mont([int.from_bytes(rng.bytes(11), "little") for _ in range(4096)])
        s[u < 0.5] = 0
        m = (u >= 0.5) & (u < 0.75)
        s[m] = small[rng.integers(0, 4096, int(m.sum()))]
        m = (u >= 0.75) & (u < 0.95)
        s[m] = mid[rng.integers(0, 4096, int(m.sum()))]
    b, bases = _bases_known_dlog(n, 2225)
    zkb.lib().zkb_srs_set_precompute(precompute)
    try:
        params = zkb.ParamsKZG(k, bases)
        got = params.commit(s)
        params.close()
    finally:
        zkb.lib().zkb_srs_set_precompute(1)
    want = oracle.g1_mul(oracle.g1_generator(), oracle.fr_inner_product(s, b))
    assert (got[:8] == want).all() and limbs_to_int(got[8:]) == R.FQ_R


def test_msm_full_size_known_dlog(oracle):
    """2^24 points with bases b_i*G: the MSM must equal (sum s_i*b_i)*G."""
    k = 24
    n = 1 << k
    s = random_field(n, 2024)
    b, bases = _bases_known_dlog(n, 2025)
    params = zkb.ParamsKZG(k, bases)
    got = params.commit(s)
    ip = oracle.fr_inner_product(s, b)
    want = oracle.g1_mul(oracle.g1_generator(), ip)
    assert (got[:8] == want).all() and limbs_to_int(got[8:]) == R.FQ_R
    # shard-split invariance at full size (the multi-GPU partition, folded on the host)
    parts = [params.commit_range(o, s[o:o + n // 8]) for o in range(0, n, n // 8)]
    assert (zkb.g1_sum(np.stack(parts)) == got).all()
    params.close()


# ---- host-buffer scheduler: pinned / pageable columns, many groups in flight ---------------------------------------------------
@pytest.mark.parametrize("depth,group_bytes", [(3, 1 << 18), (1, 1 << 18), (2, 1 << 20), (0, 0)])
def test_pipeline_groups_pageable_and_pinned(oracle, depth, group_bytes):
    """zkb_*_batch through the three-stream pipeline: forced small groups, pageable (numpy) and pinned (registered)
    columns mixed in one call, in-place and out-of-place ops; every column bit-exact with the oracle."""
    import ctypes
    lib = zkb.lib()
    k, ncols = 12, 11
    n = 1 << k
    d = zkb.EvaluationDomain(4, k)
    cols = [random_field(n, 1200 + i) for i in range(ncols)]
    want_l2c = [oracle.lagrange_to_coeff(a, k) for a in cols]
    want_c2e = [oracle.coeff_to_extended(a, k, d.extended_k) for a in cols]
    assert lib.zkb_pipeline_set(depth, group_bytes) == 0
    registered = []
    try:
        work = [a.copy() for a in cols]
        for i in range(0, ncols, 2):  # every other column page-locked
            assert lib.zkb_host_register(ctypes.c_void_p(work[i].ctypes.data), work[i].nbytes) == 0
            registered.append(work[i])
        ptrs = (ctypes.POINTER(ctypes.c_uint64) * ncols)(*[w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)) for w in work])
        assert lib.zkb_lagrange_to_coeff_batch(ptrs, ncols, k) == 0
        for i in range(ncols):
            assert (work[i] == want_l2c[i]).all(), i
        got = d.coeff_to_extended_batch(cols)
        for i in range(ncols):
            assert (got[i] == want_c2e[i]).all(), i
    finally:
        for w in registered:
            lib.zkb_host_unregister(ctypes.c_void_p(w.ctypes.data))
        lib.zkb_pipeline_set(0, 0)


def test_pipeline_large_columns_roundtrip():
    """Four 2^20 columns (32 MiB each, pageable): groups of one column, staged through the pinned ring by the host
    pool (parallel memcpy path, >= 4 MiB); NTT then inverse NTT returns the input."""
    import ctypes
    lib = zkb.lib()
    k, ncols = 20, 4
    cols = [random_field(1 << k, 1300 + i) for i in range(ncols)]
    work = [a.copy() for a in cols]
    ptrs = (ctypes.POINTER(ctypes.c_uint64) * ncols)(*[w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)) for w in work])
    w = zkb.omega(k)
    assert lib.zkb_pipeline_set(3, 32 << 20) == 0
    try:
        assert lib.zkb_ntt_fr_batch(ptrs, ncols, w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k) == 0
        single = cols[2].copy()
        zkb.best_fft(single, w, k)
        assert (work[2] == single).all()
        assert lib.zkb_lagrange_to_coeff_batch(ptrs, ncols, k) == 0  # omega_inv and 1/n: inverse of the forward NTT
        for i in range(ncols):
            assert (work[i] == cols[i]).all(), i
    finally:
        lib.zkb_pipeline_set(0, 0)


# ---- one NTT sharded over 2 GPUs (peer memory over NVLink); skipped on a single-GPU box -----------------------------------------
def test_sharded_ntt_two_gpus():
    import subprocess
    import sys
    if zkb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "tools", "dist_ntt_check.py"), "--log-n", "11", "16", "20"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 3 and all(l["parity"] for l in lines)


@pytest.mark.parametrize("slices,n,register", [(3, 5000, False), (4, 1 << 14, True), (7, 100, False), (1, 5000, False)])
def test_commit_point_range_slices(oracle, slices, n, register):
    """Pipelined host-scalar commit: `slices` point-range slices uploaded while the previous one computes, all adding into
    one bucket array (pageable scalars go through the pinned staging ring, registered ones are DMA'd directly)."""
    import ctypes
    lib = zkb.lib()
    zkb.lib().zkb_srs_set_precompute(1)
    _, g = _bases_known_dlog(n, 4000 + n)
    s = random_field(n, 4001 + n)
    s[3] = 0
    s[n // 2:] = s[n // 2]          # one long bucket run spanning later slices
    g[n - 1] = g[0]; s[n - 1] = s[0]
    k = max(6, (n - 1).bit_length())
    gp = np.zeros((1 << k, 8), dtype=np.uint64)
    gp[:n] = g
    params = zkb.ParamsKZG(k, gp)
    want = oracle.best_multiexp(s, g)
    assert lib.zkb_msm_set_slices(slices) == 0
    try:
        if register:
            assert lib.zkb_host_register(ctypes.c_void_p(s.ctypes.data), s.nbytes) == 0
        assert (params.commit(s) == want).all()
        off = n // 5
        assert (params.commit_range(off, s[off:]) == oracle.best_multiexp(s[off:], g[off:])).all()
    finally:
        if register:
            lib.zkb_host_unregister(ctypes.c_void_p(s.ctypes.data))
        lib.zkb_msm_set_slices(0)
        params.close()


# ---- ParamsKZG::setup, fixed-base windows, batch_normalize -------------------------------------
def test_fixed_base_windowed_vs_naive_vs_oracle(oracle):
    s = random_field(3000, 81)
    s[0] = 0
    s[1] = mont([1])[0]
    s[2] = mont([R.FR - 1])[0]
    s[3] = mont([1 << 240])[0]
    s[4] = mont([0xFFFF])[0]
    fast = zkb.g1_fixed_base_mul(s)
    assert (fast == zkb.g1_fixed_base_mul_naive(s)).all()
    assert (fast[:200] == oracle.g1_fixed_base_mul(s[:200])).all()


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000])
def test_batch_normalize_vs_oracle(oracle, n):
    aff = zkb.g1_fixed_base_mul(random_field(n, 5 + n))
    z = random_field(n, 6 + n, FQ_LIMBS)
    z2 = oracle.vec_op("fq", "mul", z, z)
    jac = np.zeros((n, 12), dtype=np.uint64)
    jac[:, :4] = oracle.vec_op("fq", "mul", aff[:, :4], z2)
    jac[:, 4:8] = oracle.vec_op("fq", "mul", aff[:, 4:], oracle.vec_op("fq", "mul", z2, z))
    jac[:, 8:] = z
    if n > 10:
        jac[7, 8:] = 0
        aff[7] = 0
    got = zkb.batch_normalize(jac)
    assert (got == oracle.g1_batch_normalize(jac)).all() and (got == aff).all()


def test_kzg_setup_vs_oracle(oracle):
    k = 6
    s = random_field(1, 2026)[0]
    g, gl = oracle.kzg_setup(k, s)
    params = zkb.ParamsKZG.setup(k, s)
    assert (params.get_g() == g).all()
    assert (params.get_g_lagrange() == gl).all()
    poly = random_field(1 << k, 9)
    assert (params.commit(poly) == oracle.best_multiexp(poly, g)).all()
    params.close()
    # s inside the domain is refused like upstream's panic
    w = zkb.omega(k)
    with pytest.raises(zkb.ZkbError):
        zkb.ParamsKZG.setup(k, w)


@pytest.mark.parametrize("k", [16, 20])
def test_kzg_setup_commit_consistency_full_size(oracle, k):
    """Size-independent KZG invariants tying setup, MSM and NTT together: commit(p) = [p(s)]G, and the commitment of the
    evaluations in the Lagrange basis equals the commitment of the coefficients in the monomial basis."""
    n = 1 << k
    s = random_field(1, 31 + k)[0]
    params = zkb.ParamsKZG.setup(k, s)
    coeffs = random_field(n, 32 + k)
    c1 = params.commit(coeffs)
    ps = oracle.fr_eval_polynomial(coeffs, s)
    assert (c1[:8] == oracle.g1_mul(oracle.g1_generator(), ps)).all()
    d = zkb.EvaluationDomain(4, k)
    evals = d.coeff_to_lagrange(coeffs)
    assert (params.commit_lagrange(evals) == c1).all()
    params.close()


def test_entry_points_are_thread_safe(oracle):
    """SURVEY.md §8b threading: rayon closures / two provers in one process may call concurrently — results must not mix."""
    import threading
    k = 12
    n = 1 << k
    _, g = _bases_known_dlog(n, 606)
    params = zkb.ParamsKZG(k, g)
    d = zkb.EvaluationDomain(4, k)
    polys = [random_field(n, 610 + i) for i in range(6)]
    want_c = [oracle.best_multiexp(p, g) for p in polys]
    want_e = [oracle.coeff_to_extended(p, k, d.extended_k) for p in polys]
    errs = []

    def worker(i):
        try:
            for _ in range(3):
                assert (params.commit(polys[i]) == want_c[i]).all()
                assert (d.coeff_to_extended(polys[i]) == want_e[i]).all()
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(polys))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    params.close()
    assert not errs, errs


# ---- resident polynomials and the Fr vector helpers ---------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4096, 4097, 100000])
def test_eval_polynomial_and_kate_division_vs_oracle(oracle, n):
    a = random_field(n, 700 + n)
    x = random_field(1, 701 + n)[0]
    assert (zkb.eval_polynomial(a, x) == oracle.fr_eval_polynomial(a, x)).all()
    if n >= 2:
        assert (zkb.kate_division(a, x) == oracle.fr_kate_division(a, x)).all()


def test_batch_invert_vs_oracle(oracle):
    a = random_field(5000, 19)
    a[0] = 0
    a[31:34] = 0
    a[4999] = mont([1])[0]
    assert (zkb.batch_invert(a) == oracle.fr_batch_invert(a)).all()


def test_resident_polynomial_chain(oracle):
    """The prover's chain on one advice column without leaving HBM: commit_lagrange, lagrange_to_coeff, commit,
    coeff_to_extended, eval at x, kate_division by x, commit of the quotient — each step against the oracle."""
    k = 12
    n = 1 << k
    s = random_field(1, 77)[0]
    params = zkb.ParamsKZG.setup(k, s)
    g, gl = params.get_g(), params.get_g_lagrange()
    d = zkb.EvaluationDomain(4, k)
    evals = random_field(n, 78)
    p = zkb.Polynomial(evals)
    assert len(p) == n
    assert (p.commit(params, lagrange=True) == oracle.best_multiexp(evals, gl)).all()
    coeffs = oracle.lagrange_to_coeff(evals, k)
    p.lagrange_to_coeff(d)
    assert (p.to_host() == coeffs).all()
    assert (p.commit(params) == oracle.best_multiexp(coeffs, g)).all()
    ext = p.coeff_to_extended(d)
    assert (ext.to_host() == oracle.coeff_to_extended(coeffs, k, d.extended_k)).all()
    ext.extended_to_coeff(d)
    assert (ext.to_host()[:n] == coeffs).all()
    x = random_field(1, 79)[0]
    assert (p.eval(x) == oracle.fr_eval_polynomial(coeffs, x)).all()
    q = p.kate_division(x)
    assert len(q) == n - 1
    qh = oracle.fr_kate_division(coeffs, x)
    assert (q.to_host() == qh).all()
    assert (q.commit(params) == oracle.best_multiexp(qh, g[: n - 1])).all()
    # KZG opening identity in the exponent: [p(s) - p(x)]G = [(s - x) q(s)]G, checked in the scalar field with the oracle
    ps, px, qs = (oracle.fr_eval_polynomial(c, s) for c in (coeffs, None, qh)) if False else (
        oracle.fr_eval_polynomial(coeffs, s), oracle.fr_eval_polynomial(coeffs, x), oracle.fr_eval_polynomial(qh, s))
    lhs = oracle.vec_op("fr", "sub", ps.reshape(1, 4), px.reshape(1, 4))
    rhs = oracle.vec_op("fr", "mul", oracle.vec_op("fr", "sub", s.reshape(1, 4), x.reshape(1, 4)), qs.reshape(1, 4))
    assert (lhs == rhs).all()
    for h in (p, ext, q):
        h.free()
    params.close()


# ---- the field arithmetic itself, on the device's PTX carry chains (10^6 random + edge values) -------------
@pytest.mark.parametrize("field,mod", [(0, R.FR), (1, R.FQ)])
def test_device_field_ops_vs_python_integers(field, mod):
    import ctypes
    from util import FR_LIMBS
    lib = zkb.lib()
    n = 1 << 20
    limbs = FR_LIMBS if field == 0 else FQ_LIMBS
    a = random_field(n, 11 + field, limbs)
    b = random_field(n, 13 + field, limbs)
    edge = [0, 1, 2, mod - 1, mod - 2, 1 << 253, (1 << 128) - 1, 1 << 128, (1 << 64) - 1, (1 << 32) - 1, mod >> 1, (mod + 1) >> 1]
    m = len(edge)
    ea, eb = np.repeat(ints_to_limbs(edge), m, axis=0), np.tile(ints_to_limbs(edge), (m, 1))   # every ordered pair
    a[: m * m], b[: m * m] = ea, eb
    out = np.zeros_like(a)
    p = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    rinv = pow(1 << 256, mod - 2, mod)
    ops = {0: lambda x, y: x * y * rinv % mod, 1: lambda x, y: (x + y) % mod, 2: lambda x, y: (x - y) % mod,
           3: lambda x, y: x * x * rinv % mod, 4: lambda x, y: (-x) % mod, 5: lambda x, y: 2 * x % mod, 6: lambda x, y: x * rinv % mod,
           7: lambda x, y: x * y * rinv % mod, 8: lambda x, y: (x + y) % mod, 9: lambda x, y: (x - y) % mod}   # lazy forms, reduced
    check_idx = list(range(m * m)) + list(range(m * m, n, 997))      # all edge pairs + a 1000-element sample in Python
    ai = [limbs_to_int(a[i]) for i in check_idx]
    bi = [limbs_to_int(b[i]) for i in check_idx]
    for op, f in ops.items():
        assert lib.zkb_field_vec_op(field, op, p(a), p(b), p(out), n) == 0, lib.zkb_last_error()
        got = [limbs_to_int(out[i]) for i in check_idx]
        assert got == [f(x, y) for x, y in zip(ai, bi)], op
        # the whole 2^20 vector against the C oracle for the three binary ops
        if op < 3 or op in (7, 8, 9):
            from oracle import coracle
            coracle.build()
            assert (out == coracle.vec_op("fr" if field == 0 else "fq", {0: "mul", 1: "add", 2: "sub", 7: "mul", 8: "add", 9: "sub"}[op], a, b)).all()
    # lazy inputs proper (values in [M, 2M)): a + M and b + M are the same residues in lazy form; the raw lazy product stays < 2M
    def plus_mod(v):
        w = v.copy()
        carry = np.zeros(v.shape[0], dtype=np.uint64)
        for j in range(4):
            t = w[:, j] + limbs[j]
            c1 = (t < w[:, j]).astype(np.uint64)
            t2 = t + carry
            c2 = (t2 < t).astype(np.uint64)
            w[:, j] = t2
            carry = c1 + c2
        return w
    la, lb = plus_mod(a), plus_mod(b)
    for op, f in ((7, ops[7]), (8, ops[8]), (9, ops[9])):
        assert lib.zkb_field_vec_op(field, op, p(la), p(lb), p(out), n) == 0
        assert [limbs_to_int(out[i]) for i in check_idx] == [f(x, y) for x, y in zip(ai, bi)], ("lazy inputs", op)
    assert lib.zkb_field_vec_op(field, 10, p(la), p(lb), p(out), n) == 0
    raw = [limbs_to_int(out[i]) for i in check_idx]
    assert all(v < 2 * mod and v % mod == x * y * rinv % mod for v, x, y in zip(raw, ai, bi))


# ---- best_fft over G1 and g_to_lagrange ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [0, 1, 4, 5])
def test_g1_fft_vs_definition(oracle, k):
    n = 1 << k
    pts = zkb.g1_fixed_base_mul(random_field(n, 160 + k))
    if n >= 8:
        pts[3] = 0
        pts[5] = pts[4]
    w = zkb.omega(k)
    assert (zkb.best_fft_g1(pts, w, k) == oracle.g1_fft_naive(pts, w)).all()


@pytest.mark.parametrize("k", [6, 12])
def test_g_to_lagrange_equals_setup_lagrange(k):
    """Two independent device paths to the Lagrange-basis SRS: the inverse G1 FFT of g (g_to_lagrange) and the direct
    [l_i(s)]G of ParamsKZG::setup must agree point for point; commit_lagrange through either gives the same commitment."""
    s = random_field(1, 900 + k)[0]
    a = zkb.ParamsKZG.setup(k, s)
    b = zkb.ParamsKZG.from_g(k, a.get_g())
    assert (b.get_g_lagrange() == a.get_g_lagrange()).all()
    ev = random_field(1 << k, 901 + k)
    assert (a.commit_lagrange(ev) == b.commit_lagrange(ev)).all()
    a.close()
    b.close()


@pytest.mark.parametrize("n", [1, 64, 65, 4097, 200000])
def test_grand_product_pieces(oracle, n):
    """The permutation / lookup grand product z[i] = prod_{j<i} num[j] / den[j] from its device pieces (batch_invert, mul,
    prefix_product on resident handles) against the oracle's element-wise ops and a Python scan on a sample."""
    num, den = random_field(n, 810 + n), random_field(n, 811 + n)
    if n > 10:
        den[3] = mont([1])[0]
    d = zkb.Polynomial(den).batch_invert()
    z = zkb.Polynomial(num).mul(d)
    ratio = oracle.vec_op("fr", "mul", num, oracle.fr_batch_invert(den))
    assert (z.to_host() == ratio).all()
    got = z.prefix_product().to_host()
    m = min(n, 300)
    acc, want = 1, []
    rinv_vals = [R.from_mont(limbs_to_int(x), R.FR) for x in ratio[:m]]
    for v in rinv_vals:
        want.append(acc)
        acc = acc * v % R.FR
    assert [R.from_mont(limbs_to_int(x), R.FR) for x in got[:m]] == want
    # the whole vector: z[i+1] = z[i] * ratio[i]
    assert (got[1:] == oracle.vec_op("fr", "mul", got[:-1], ratio[:-1])).all() and limbs_to_int(got[0]) == R.FR_R
    d.free()
    z.free()


def test_shutdown_and_reinit_in_subprocess():
    """zkb_shutdown releases everything (SRS tables, plans, pipeline, staging pool, polynomials); a second init works."""
    import subprocess
    import sys
    code = (
        "import importlib, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from util import random_field\n"
        "zkb = importlib.import_module('zksnap-circuits-halo2_b200')\n"
        "outs = []\n"
        "for rep in range(2):\n"
        "    zkb.init(0)\n"
        "    k = 12; s = random_field(1, 5)[0]\n"
        "    p = zkb.ParamsKZG.setup(k, s)\n"
        "    poly = random_field(1 << k, 6)\n"
        "    c = p.commit(poly)\n"
        "    d = zkb.EvaluationDomain(4, k)\n"
        "    e = d.coeff_to_extended_batch([poly, poly])[1]\n"
        "    h = zkb.Polynomial(poly); ev = h.eval(s)\n"
        "    outs.append((c.tobytes(), e.tobytes(), ev.tobytes()))\n"
        "    zkb.shutdown()\n"
        "assert outs[0] == outs[1]\n"
        "print('reinit ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "reinit ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_error_paths_return_codes_and_leave_the_library_usable(oracle):
    """C functions return negative codes + a message (the shim panics on them, like the reference's unwrap/expect); a failed
    call must not poison later ones."""
    import ctypes
    lib = zkb.lib()
    u64p_ = ctypes.POINTER(ctypes.c_uint64)
    out = np.zeros(12, dtype=np.uint64)
    s = random_field(16, 1)
    ERR_ARG, ERR_HANDLE = -1, -5
    assert lib.zkb_msm_g1_srs(123456789, s.ctypes.data_as(u64p_), 16, out.ctypes.data_as(u64p_)) == ERR_HANDLE
    assert b"unknown SRS handle" in lib.zkb_last_error()
    assert lib.zkb_srs_release(987654321) == ERR_HANDLE
    assert lib.zkb_poly_free(555) == ERR_HANDLE
    assert lib.zkb_ntt_fr(s.ctypes.data_as(u64p_), zkb.omega(4).ctypes.data_as(u64p_), 0) == ERR_ARG        # log_n out of range
    assert lib.zkb_ntt_fr(s.ctypes.data_as(u64p_), zkb.omega(4).ctypes.data_as(u64p_), 29) == ERR_ARG
    assert lib.zkb_ntt_fr(None, zkb.omega(4).ctypes.data_as(u64p_), 4) == ERR_ARG                             # NULL data
    assert lib.zkb_coeff_to_extended(s.ctypes.data_as(u64p_), s.ctypes.data_as(u64p_), 4, 3) == ERR_ARG       # extended_k < k
    assert lib.zkb_msm_g1(s.ctypes.data_as(u64p_), None, 16, out.ctypes.data_as(u64p_)) == ERR_ARG
    assert lib.zkb_msm_set_params(40, 0) == ERR_ARG
    assert lib.zkb_kzg_setup(0, s.ctypes.data_as(u64p_), None, None) == ERR_ARG
    _, g = _bases_known_dlog(16, 3)
    params = zkb.ParamsKZG(4, g)
    assert lib.zkb_msm_g1_srs(params.handle_g, random_field(17, 2).ctypes.data_as(u64p_), 17, out.ctypes.data_as(u64p_)) == ERR_ARG
    assert lib.zkb_msm_g1_srs_range(params.handle_g, 10, s.ctypes.data_as(u64p_), 10, out.ctypes.data_as(u64p_)) == ERR_ARG
    p = zkb.Polynomial(s)
    h = ctypes.c_uint64(0)
    assert lib.zkb_poly_coeff_to_extended(p._h, 5, 7, ctypes.byref(h)) == ERR_ARG                              # 16 elements != 2^5
    assert lib.zkb_poly_lagrange_to_coeff(p._h, 3) == ERR_ARG
    # everything still works afterwards
    assert (params.commit(s) == oracle.best_multiexp(s, g)).all()
    a = s.copy()
    zkb.best_fft(a, zkb.omega(4), 4)
    assert (a == oracle.best_fft(s, zkb.omega(4), 4)).all()
    assert (p.lagrange_to_coeff(zkb.EvaluationDomain(4, 4)).to_host() == oracle.lagrange_to_coeff(s, 4)).all()
    p.free()
    params.close()


def test_gwc_witness_polynomial_on_resident_handles(oracle):
    """A GWC-style opening witness built without leaving HBM: fold three polynomials with powers of v (scale_add), subtract the
    folded evaluation (add_const), divide by (X - x) (kate_division), commit — against the same steps done with the oracle;
    the quotient is exact, so q(X) (X - x) reproduces the folded polynomial."""
    k = 11
    n = 1 << k
    s = random_field(1, 301)[0]
    params = zkb.ParamsKZG.setup(k, s)
    polys = [random_field(n, 310 + i) for i in range(3)]
    v, x = random_field(2, 320)
    acc = zkb.Polynomial(polys[0])
    folded = polys[0]
    for p in polys[1:]:
        h = zkb.Polynomial(p)
        acc.scale_add(v, h)
        h.free()
        folded = oracle.vec_op("fr", "add", oracle.vec_op("fr", "mul", folded, np.tile(v, (n, 1))), p)
    assert (acc.to_host() == folded).all()
    ev = acc.eval(x)
    assert (ev == oracle.fr_eval_polynomial(folded, x)).all()
    neg = oracle.vec_op("fr", "sub", np.zeros((1, 4), dtype=np.uint64), ev.reshape(1, 4))[0]
    acc.add_const(neg)
    shifted = folded.copy()
    shifted[0] = oracle.vec_op("fr", "add", folded[:1], neg.reshape(1, 4))[0]
    assert (acc.to_host() == shifted).all()
    q = acc.kate_division(x)
    qh = oracle.fr_kate_division(shifted, x)
    assert (q.to_host() == qh).all()
    assert (oracle.fr_eval_polynomial(shifted, x) == 0).all()          # x is a root, so the division is exact
    assert (q.commit(params) == oracle.best_multiexp(qh, params.get_g()[: n - 1])).all()
    acc.scale_add(v)                                                    # scale only
    assert (acc.to_host() == oracle.vec_op("fr", "mul", shifted, np.tile(v, (n, 1)))).all()
    for hnd in (acc, q):
        hnd.free()
    params.close()


def test_srs_table_automatic_policy(oracle):
    """Default policy: the SRS window table is built once a handle has served 96 commitments (ski-rental); before and after,
    through single and batched commits, every result equals the oracle's."""
    import ctypes
    lib = zkb.lib()
    lib.zkb_srs_set_precompute(2)
    try:
        k, n = 8, 256
        _, g = _bases_known_dlog(n, 4242)
        params = zkb.ParamsKZG(k, g)
        bits, nbytes, commits = ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_uint64()
        polys = [random_field(n, 4300 + i) for i in range(6)]
        want = [oracle.best_multiexp(p, g) for p in polys]
        for i in range(95):
            assert (params.commit(polys[i % 6]) == want[i % 6]).all()
        lib.zkb_srs_table_info(params.handle_g, ctypes.byref(bits), ctypes.byref(nbytes), ctypes.byref(commits))
        assert bits.value == 0 and commits.value == 95
        got = params.commit_batch(polys)          # 6 more commitments: crosses the threshold
        for i in range(6):
            assert (got[i] == want[i]).all()
        lib.zkb_srs_table_info(params.handle_g, ctypes.byref(bits), ctypes.byref(nbytes), ctypes.byref(commits))
        assert bits.value > 0 and nbytes.value > 0 and commits.value == 101
        for i in range(6):
            assert (params.commit(polys[i]) == want[i]).all()
        params.close()
    finally:
        lib.zkb_srs_set_precompute(1)


# ---- tests/golden/widened_kats.json: the widened rows against vectors made from Python integers by the definitions ------------
WGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "widened_kats.json")))


def _h(x):
    return int(x, 16)


def _aff_or_id(p):
    return aff((_h(p[0]), _h(p[1])) if p else None)


def test_widened_golden_kzg_setup():
    c = WGOLD["kzg_setup"]
    params = zkb.ParamsKZG.setup(c["k"], mont([_h(c["s"])])[0])
    assert (params.get_g() == np.array([_aff_or_id(p) for p in c["g"]])).all()
    assert (params.get_g_lagrange() == np.array([_aff_or_id(p) for p in c["g_lagrange"]])).all()
    params.close()


def test_widened_golden_poly_and_scans():
    c = WGOLD["poly"]
    a, b = mont([_h(x) for x in c["coeffs"]]), mont([_h(c["point"])])[0]
    assert unmont(zkb.eval_polynomial(a, b)) == [_h(c["eval"])]
    assert unmont(zkb.kate_division(a, b)) == [_h(x) for x in c["kate_quotient"]]
    c = WGOLD["batch_invert"]
    assert unmont(zkb.batch_invert(mont([_h(x) for x in c["in"]]))) == [_h(x) for x in c["out"]]
    c = WGOLD["prefix_product"]
    p = zkb.Polynomial(mont([_h(x) for x in c["in"]]))
    assert unmont(p.prefix_product().to_host()) == [_h(x) for x in c["out"]]
    p.free()


def test_widened_golden_batch_normalize_and_g1_fft():
    c = WGOLD["batch_normalize"]
    jac = np.array([sum((R.fq_encode(_h(x)) for x in row), []) for row in c["jacobian"]], dtype=np.uint64)
    assert (zkb.batch_normalize(jac) == np.array([_aff_or_id(p) for p in c["affine"]])).all()
    c = WGOLD["g1_fft"]
    got = zkb.best_fft_g1(np.array([_aff_or_id(p) for p in c["in"]]), mont([_h(c["omega"])])[0], c["k"])
    assert (got == np.array([_aff_or_id(p) for p in c["out"]])).all()


# ---- quotient evaluation on resident cosets: zkb_graph_evaluate vs the oracle and the definition -----------
import graph_cases as GC  # noqa: E402


def _gpu_graph(c):
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    cols = [[zkb.Polynomial(a) for a in group] for group in (fx, ad, ins)]
    values = zkb.Polynomial(prev)
    c["graph"].evaluate(values, *cols, challenges=ch, y=y, rot_scale=c["rot_scale"])
    out = values.to_host()
    for p in [values] + sum(cols, []):
        p.free()
    return out


@pytest.mark.parametrize("i", range(3))
def test_graph_evaluate_golden(i):
    c = GC.golden_cases()[i]
    assert GC.unmont(_gpu_graph(c)) == c["expected"]


@pytest.mark.parametrize("seed,isize,rot_scale,ngates,depth", [(31, 1, 1, 3, 4), (32, 2, 1, 3, 4), (33, 64, 4, 4, 5), (34, 1 << 12, 4, 6, 6),
                                                               (35, 1 << 15, 8, 12, 6), (36, 1 << 10, 2, 40, 5)])
def test_graph_evaluate_vs_oracle(oracle, seed, isize, rot_scale, ngates, depth):
    c = GC.random_case(seed, isize, rot_scale, ngates=ngates, depth=depth)
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, rot_scale, prev)
    assert (_gpu_graph(c) == want).all()
    info = g.last_info()
    assert info["instructions"] < len(g.calculations) and info["slots"] < max(g.num_intermediates, 2)


def test_graph_permutation_term_vs_oracle(oracle):
    ncols, isize, rs = 4, 1 << 12, 4
    g, _ = GC.permutation_term_graph(ncols)
    fixed = [random_field(isize, 600 + i) for i in range(2 + ncols)]
    advice = [random_field(isize, 620 + i) for i in range(1 + ncols)]
    beta, gamma = random_field(2, 640)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fixed, advice, [], None,
                                 beta, gamma, None, None, rs, np.zeros((isize, 4), dtype=np.uint64))
    values = zkb.Polynomial(np.zeros((isize, 4), dtype=np.uint64))
    g.evaluate(values, [zkb.Polynomial(a) for a in fixed], [zkb.Polynomial(a) for a in advice], beta=beta, gamma=gamma, rot_scale=rs)
    assert (values.to_host() == want).all()


def test_graph_evaluate_errors_leave_library_usable():
    ev = GC.ev
    V = ev.ValueSource
    a, b = zkb.Polynomial(random_field(64, 1)), zkb.Polynomial(random_field(32, 2))
    g = ev.GraphEvaluator()
    g.add_calculation(ev.MUL, V(ev.ADVICE, 0, g.add_rotation(0)), V(ev.ADVICE, 1, 0))
    out = zkb.Polynomial(np.zeros((64, 4), dtype=np.uint64))
    with pytest.raises(zkb.ZkbError):      # column of another size
        g.evaluate(out, advice=[a, b])
    with pytest.raises(zkb.ZkbError):      # values is also a column
        g.evaluate(a, advice=[a, a])
    with pytest.raises(zkb.ZkbError):      # column index out of range
        g.evaluate(out, advice=[a])
    bad = ev.GraphEvaluator()
    bad.num_intermediates = 2
    bad.calculations.append((ev.ADD, 0, V(ev.INTERMEDIATE, 1), V(ev.CONSTANT, 1), V(ev.CONSTANT, 0)))
    with pytest.raises(zkb.ZkbError):      # intermediate read before it is written
        bad.evaluate(out)
    g.evaluate(out, advice=[a, a])         # and a good call still works: out = a * a
    sq = ev.GraphEvaluator()
    sq.add_calculation(ev.SQUARE, V(ev.ADVICE, 0, sq.add_rotation(0)))
    out2 = zkb.Polynomial(np.zeros((64, 4), dtype=np.uint64))
    sq.evaluate(out2, advice=[a])
    assert (out.to_host() == out2.to_host()).all()


def test_graph_evaluate_full_size_gate(oracle):
    """The wrapper's extended domain (k = 22 -> 2^24 rows, rot_scale 4) with halo2-base's gate q (a + b c - d): 64 sampled rows
    against Python integers, and linearity in the selector column over the whole vector (oracle element-wise add)."""
    isize, rs = 1 << 24, 4
    adv = random_field(isize, 7001)
    q1, q2 = random_field(isize, 7002), random_field(isize, 7003)
    q12 = oracle.vec_op("fr", "add", q1, q2)
    g = GC.build_custom_gates([GC.halo2_base_gate()])
    y = random_field(1, 7004)[0]
    A = zkb.Polynomial(adv)
    outs = []
    for q in (q1, q2, q12):
        Q = zkb.Polynomial(q)
        v = zkb.Polynomial(np.zeros((isize, 4), dtype=np.uint64))
        g.evaluate(v, fixed=[Q], advice=[A], y=y, rot_scale=rs)
        outs.append(v.to_host())
        Q.free()
        v.free()
    assert (oracle.vec_op("fr", "add", outs[0], outs[1]) == outs[2]).all()
    rng = np.random.default_rng(5)
    rows = [0, 1, isize - 1, isize - 4, isize - 13] + [int(x) for x in rng.integers(0, isize, 59)]
    ai = lambda r: R.from_mont(limbs_to_int(adv[r % isize]), R.FR)  # noqa: E731
    for r in rows:
        want = R.from_mont(limbs_to_int(q1[r]), R.FR) * (ai(r) + ai(r + rs) * ai(r + 2 * rs) - ai(r + 3 * rs)) % R.FR
        assert R.from_mont(limbs_to_int(outs[0][r]), R.FR) == want
    # previous value: a second pass over the result adds prev * y
    v = zkb.Polynomial(outs[0])
    Q = zkb.Polynomial(q2)
    g.evaluate(v, fixed=[Q], advice=[A], y=y, rot_scale=rs)
    got = v.to_host()
    assert (got == oracle.vec_op("fr", "add", oracle.vec_op("fr", "mul", outs[0], np.tile(y, (isize, 1))), outs[1])).all()


def test_graph_lookup_terms_vs_definition():
    g, cols, sc, prev, want = GC.lookup_case(93, 1 << 10, 4)
    values = zkb.Polynomial(GC.mont(prev))
    g.evaluate(values, [zkb.Polynomial(GC.mont(c)) for c in cols["fixed"]], [zkb.Polynomial(GC.mont(c)) for c in cols["advice"]],
               beta=GC.mont([sc["beta"]])[0], gamma=GC.mont([sc["gamma"]])[0], theta=GC.mont([sc["theta"]])[0], y=GC.mont([sc["y"]])[0],
               rot_scale=4)
    assert GC.unmont(values.to_host()) == want


def test_params_read_from_file_into_hbm(oracle, tmp_path):
    """ParamsKZG::read (RawBytes layout): the SRS goes from the file to HBM through pinned staging; commitments equal those of
    the params the file was written from; a corrupted point is refused with checks on and accepted unchecked."""
    k = 10
    s = random_field(1, 4242)[0]
    a = zkb.ParamsKZG.setup(k, s)
    path = str(tmp_path / "kzg_bn254_10.srs")
    a.write(path)
    assert os.path.getsize(path) == 4 + 2 * (64 << k) + 256
    b = zkb.ParamsKZG.read(path)
    assert b.k == k and (b.get_g() == a.get_g()).all() and (b.get_g_lagrange() == a.get_g_lagrange()).all()
    poly = random_field(1 << k, 4243)
    assert (b.commit(poly) == oracle.best_multiexp(poly, a.get_g())).all()
    assert (b.commit_lagrange(poly) == a.commit_lagrange(poly)).all()
    b.close()
    raw = bytearray(open(path, "rb").read())
    raw[4 + 64 * 17 + 3] ^= 0x40            # one bit of g[17].x
    open(path, "wb").write(raw)
    with pytest.raises(zkb.ZkbError):
        zkb.ParamsKZG.read(path)
    c = zkb.ParamsKZG.read(path, check_points=False)
    c.close()
    raw[4 + 64 * 17: 4 + 64 * 18] = bytes([0xFF]) * 64     # non-canonical limbs
    open(path, "wb").write(raw)
    with pytest.raises(zkb.ZkbError):
        zkb.ParamsKZG.read(path)
    open(path, "wb").write(raw[:1000])                       # truncated
    with pytest.raises(zkb.ZkbError):
        zkb.ParamsKZG.read(path)
    a.close()


def test_quotient_pipeline_polynomial_identity():
    """What evaluate_h relies on, end to end on resident polynomials: with A = coeff_to_extended(a), Q = coeff_to_extended(q) and
    rot_scale = 2^(extended_k - k), the row-wise gate value q (a + a(wX) a(w^2 X) - a(w^3 X)) on the extended coset IS the coset
    evaluation of that polynomial: extended_to_coeff of the values gives a polynomial G of degree < 3n with
    G(x) = q(x) (a(x) + a(w x) a(w^2 x) - a(w^3 x)) at a random x (Python integers combine the device evaluations)."""
    k = 16
    n = 1 << k
    d = zkb.EvaluationDomain(4, k)
    a, q = random_field(n, 9101), random_field(n, 9102)
    pa, pq = zkb.Polynomial(a), zkb.Polynomial(q)
    A, Q = pa.coeff_to_extended(d), pq.coeff_to_extended(d)
    g = GC.build_custom_gates([GC.halo2_base_gate()])
    values = zkb.Polynomial(np.zeros((d.extended_len(), 4), dtype=np.uint64))
    g.evaluate(values, fixed=[Q], advice=[A], y=random_field(1, 9103)[0], rot_scale=1 << (d.extended_k - k))
    values.extended_to_coeff(d)
    full = values.to_host()
    G = full[: 3 * n]
    assert not full[3 * n:].any()        # degree < 3n: the values were the coset evaluations of a low-degree polynomial
    x = random_field(1, 9104)[0]
    xi = unmont(x)[0]
    w = R.omega_for(k)
    at = [unmont(pa.eval(mont([xi * pow(w, r, R.FR) % R.FR])[0]))[0] for r in range(4)]
    want = unmont(pq.eval(x))[0] * (at[0] + at[1] * at[2] - at[3]) % R.FR
    assert unmont(zkb.eval_polynomial(G, x))[0] == want
    for p in (pa, pq, A, Q, values):
        p.free()


def test_h_pieces_sliced_committed_and_opened_on_the_device(oracle):
    """After the quotient evaluation and extended_to_coeff, h(X) is cut into pieces of n coefficients that are committed and
    evaluated one by one: Polynomial.slice keeps that in HBM."""
    k = 10
    n = 1 << k
    params = zkb.ParamsKZG.setup(k, random_field(1, 9301)[0])
    g = params.get_g()
    h = random_field(4 * n, 9302)
    H = zkb.Polynomial(h)
    x = random_field(1, 9303)[0]
    for i in range(3):
        piece = H.slice(i * n, n)
        assert len(piece) == n and (piece.to_host() == h[i * n:(i + 1) * n]).all()
        assert (piece.commit(params) == oracle.best_multiexp(h[i * n:(i + 1) * n], g)).all()
        assert (piece.eval(x) == oracle.fr_eval_polynomial(h[i * n:(i + 1) * n], x)).all()
        piece.free()
    assert len(H.slice(4 * n, 0)) == 0
    with pytest.raises(zkb.ZkbError):
        H.slice(3 * n, n + 1)
    params.close()


def test_graph_evaluate_dev_whole_domain_and_row_window(oracle):
    """zkb_graph_evaluate_dev on device pointers: whole-domain mode equals the handle path; row-window mode on windows cut by
    hand (two shards, cyclic halos) equals the whole-domain result; a rotation outside the halo is refused."""
    import torch
    dev = torch.device("cuda", 0)
    isize, rs = 1 << 12, 4
    c = GC.random_case(4711, isize, rs, ngates=5, depth=5)
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, rs, prev)
    stream = torch.cuda.current_stream().cuda_stream

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)

    cols = [[dv(a) for a in grp] for grp in (fx, ad, ins)]
    values = dv(prev)
    g.evaluate_dev(values.data_ptr(), isize, *[[t.data_ptr() for t in grp] for grp in cols], challenges=ch, y=y, rot_scale=rs, stream=stream)
    torch.cuda.synchronize()
    assert (values.cpu().numpy().view(np.uint64) == want).all()
    lo, hi = g.rotation_span(rs)
    rows = isize // 2
    for r in range(2):
        idx = np.arange(r * rows - lo, (r + 1) * rows + hi) % isize
        win = [[dv(a[idx]) for a in grp] for grp in (fx, ad, ins)]
        v = dv(prev[r * rows:(r + 1) * rows])
        g.evaluate_dev(v.data_ptr(), rows, *[[t.data_ptr() for t in grp] for grp in win], challenges=ch, y=y, rot_scale=rs,
                       window=True, halo_lo=lo, halo_hi=hi, stream=stream)
        torch.cuda.synchronize()
        assert (v.cpu().numpy().view(np.uint64) == want[r * rows:(r + 1) * rows]).all()
        with pytest.raises(zkb.ZkbError):
            g.evaluate_dev(v.data_ptr(), rows, *[[t.data_ptr() for t in grp] for grp in win], challenges=ch, y=y, rot_scale=rs,
                           window=True, halo_lo=lo, halo_hi=hi - 1, stream=stream)


def test_sharded_quotient_single_rank_wraps_onto_itself(oracle):
    """distributed.ShardedQuotient with one rank: the halo exchange is a cyclic copy of the rank's own rows."""
    import torch
    zd = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
    dev = torch.device("cuda", 0)
    isize, rs = 1 << 11, 2
    c = GC.random_case(4712, isize, rs, ngates=4, depth=5)
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, rs, prev)

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)

    values = dv(prev)
    zd.ShardedQuotient(g, rs).run(values, [dv(a) for a in fx], [dv(a) for a in ad], [dv(a) for a in ins], challenges=ch, y=y)
    torch.cuda.synchronize()
    assert (values.cpu().numpy().view(np.uint64) == want).all()


def test_sharded_quotient_two_gpus():
    import subprocess
    import sys
    if zkb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(root, "tests", "tools", "dist_quotient_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 3 and all(l["parity"] for l in lines)


def test_eip196_public_vectors():
    """External known answers for the G1 group law (EIP-196 precompile tests, tests/golden/eip196_kats.json) through
    best_multiexp on the GPU: a 2-point MSM with unit scalars is the addition, a 1-point MSM the scalar multiplication; also
    through a registered SRS with the window table."""
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eip196_kats.json")))
    one = mont([1])[0]
    for c in kat["add"]:
        out = zkb.best_multiexp(np.array([one, one]), np.array([aff((int(c["a"][0], 16), int(c["a"][1], 16))), aff((int(c["b"][0], 16), int(c["b"][1], 16)))]))
        assert R.g1_jacobian_decode([int(x) for x in out]) == (int(c["sum"][0], 16), int(c["sum"][1], 16))
    for c in kat["mul"]:
        p = aff((int(c["p"][0], 16), int(c["p"][1], 16)))
        s = mont([int(c["s"], 16) % R.FR])
        want = (int(c["out"][0], 16), int(c["out"][1], 16))
        assert R.g1_jacobian_decode([int(x) for x in zkb.best_multiexp(s, np.array([p]))]) == want
        g = np.zeros((64, 8), dtype=np.uint64)
        g[0] = p
        zkb.lib().zkb_srs_set_precompute(1)
        params = zkb.ParamsKZG(6, g)
        sc = np.zeros((64, 4), dtype=np.uint64)
        sc[0] = s[0]
        assert R.g1_jacobian_decode([int(x) for x in params.commit(sc)]) == want
        params.close()


@pytest.mark.parametrize("ncols,chunk_len,isize", [(5, 2, 1 << 10), (7, 3, 1 << 12)])
def test_permutation_argument_graph(ncols, chunk_len, isize):
    """evaluation.permutation_graph (every h(X) term of the permutation argument) on the GPU against the formulas."""
    g, cols, sc, prev, want = GC.permutation_case(70 + ncols, isize, 4, ncols=ncols, chunk_len=chunk_len)
    values = zkb.Polynomial(GC.mont(prev))
    g.evaluate(values, [zkb.Polynomial(GC.mont(c)) for c in cols["fixed"]], [zkb.Polynomial(GC.mont(c)) for c in cols["advice"]],
               beta=GC.mont([sc["beta"]])[0], gamma=GC.mont([sc["gamma"]])[0], y=GC.mont([sc["y"]])[0], rot_scale=4)
    assert GC.unmont(values.to_host()) == want


def test_quotient_of_a_satisfied_circuit_is_a_polynomial():
    """evaluate_h's purpose, end to end on resident polynomials: witness and selector (Lagrange) -> lagrange_to_coeff ->
    coeff_to_extended -> the gate on the coset (rot_scale = 4) -> times 1 / (X^n - 1) -> extended_to_coeff: a satisfying witness
    gives a quotient of degree < 2n (zeros from coefficient 2n on), one wrong cell does not."""
    k = 12
    n = 1 << k
    d = zkb.EvaluationDomain(4, k)
    g = GC.build_custom_gates([GC.halo2_base_gate()])
    T = zkb.Polynomial(GC.mont(GC.vanishing_inverse_on_coset(k, d.extended_k)))
    one = mont([1])[0]
    for broken in (None, 1234):
        w, q = GC.satisfied_gate_witness(k, 654, break_cell=broken)
        W = zkb.Polynomial(GC.mont(w)).lagrange_to_coeff(d).coeff_to_extended(d)
        Q = zkb.Polynomial(GC.mont(q)).lagrange_to_coeff(d).coeff_to_extended(d)
        h = zkb.Polynomial(np.zeros((d.extended_len(), 4), dtype=np.uint64))
        g.evaluate(h, fixed=[Q], advice=[W], y=one, rot_scale=1 << (d.extended_k - k))
        h.mul(T).extended_to_coeff(d)
        coeffs = h.to_host()
        if broken is None:
            assert coeffs[:2 * n].any() and not coeffs[2 * n:].any()
        else:
            assert coeffs[2 * n:3 * n].any()
        for p in (W, Q, h):
            p.free()


def test_quotient_evaluator_evaluate_h_on_resident_cosets(oracle):
    """QuotientEvaluator.evaluate_h on the device (default per-graph call: zkb_graph_evaluate on handles) against the same chain
    of graphs evaluated by the oracle."""
    import random
    ev = GC.ev
    rnd = random.Random(2025)
    isize, rs = 1 << 10, 4
    col = lambda s: random_field(isize, s)  # noqa: E731
    fixed, advice, instance = [col(1), col(2)], [col(3), col(4), col(5)], [col(6)]
    l0, l_last, l_active, xc = col(7), col(8), col(9), col(10)
    perm_cols = [("advice", 0), ("advice", 2), ("instance", 0)]
    sigmas, zs = [col(11 + i) for i in range(3)], [col(20), col(21)]
    lz, la, ls = col(30), col(31), col(32)
    gates = [GC.halo2_base_gate(1, 0), ("sum", ("prod", ("advice", 0, -1), ("fixed", 1, 0)), ("neg", ("instance", 0, 2)))]
    l_in, l_tab = [("advice", 1, 0), ("advice", 2, 1)], [("fixed", 0, 0), ("fixed", 1, -1)]
    y, beta, gamma, theta = random_field(4, 40)
    q = ev.QuotientEvaluator(gates, dict(columns=perm_cols, chunk_len=2, last_rotation=-3), [(l_in, l_tab)])
    state = {"v": np.zeros((isize, 4), dtype=np.uint64)}

    def oracle_eval(graph, values, f, a, i):
        state["v"] = oracle.graph_evaluate(graph.calc_array(), graph.num_intermediates, GC.mont(graph.constants), graph.rotations, f, a, i,
                                           None, beta, gamma, theta, y, rs, state["v"])

    q.evaluate_h(None, fixed, advice, instance, None, y, beta, gamma, theta, rs, l0, l_last, l_active, xc, sigmas, zs, [(lz, la, ls)],
                 evaluate=oracle_eval)
    P = zkb.Polynomial
    values = P(np.zeros((isize, 4), dtype=np.uint64))
    q.evaluate_h(values, [P(c) for c in fixed], [P(c) for c in advice], [P(c) for c in instance], None, y, beta, gamma, theta, rs,
                 P(l0), P(l_last), P(l_active), P(xc), [P(c) for c in sigmas], [P(c) for c in zs], [(P(lz), P(la), P(ls))])
    assert (values.to_host() == state["v"]).all()


# ---- the MSM's own bucket sort (csrc/bucket_sort.cuh) on the device ----------------------------------------------------------------
def _gpu_bucket_sort(keys, key_bits, tile=0):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    k = keys.copy()
    v = np.arange(keys.size, dtype=np.uint32)          # value = original position
    u32p = ctypes.POINTER(ctypes.c_uint32)
    rc = zkb.lib().zkb_bucket_sort_pairs(k.ctypes.data_as(u32p), v.ctypes.data_as(u32p), keys.size, key_bits, tile)
    assert rc == 0, zkb.lib().zkb_last_error()
    assert (k == np.sort(keys)).all(), "keys not sorted"
    assert (keys[v] == k).all(), "a value left its key"
    assert (np.sort(v) == np.arange(keys.size, dtype=np.uint32)).all(), "values lost or duplicated"


@pytest.mark.parametrize("n,key_bits,tile", [(1, 1, 0), (7, 3, 0), (5000, 11, 0), (5000, 12, 1024), (100_000, 16, 1024), (1 << 20, 21, 0),
                                            (1 << 20, 22, 8192), (1 << 21, 23, 0), (1 << 20, 24, 2048), (3_000_001, 17, 4096), (1 << 22, 13, 0)])
def test_bucket_sort_uniform_keys(n, key_bits, tile):
    rng = np.random.default_rng(n + key_bits)
    _gpu_bucket_sort(rng.integers(0, 1 << key_bits, size=n, dtype=np.uint32), key_bits, tile)


def test_bucket_sort_skewed_and_degenerate_keys():
    rng = np.random.default_rng(11)
    n = 1 << 21
    # witness-like: most entries in a handful of buckets, the rest spread out
    keys = np.where(rng.random(n) < 0.8, rng.integers(0, 8, size=n), rng.integers(0, 1 << 21, size=n)).astype(np.uint32)
    _gpu_bucket_sort(keys, 21)
    _gpu_bucket_sort(np.full(n, 0x12345, dtype=np.uint32), 21)                 # the all-equal column: ONE bucket gets everything
    _gpu_bucket_sort(np.full(n, (1 << 21) - 1, dtype=np.uint32), 21, 1024)      # the last bucket of the last segment
    _gpu_bucket_sort((np.arange(n, dtype=np.uint32)[::-1] % (1 << 18)).astype(np.uint32), 18)
    # segments whose sizes are exact multiples of the tile, and one-entry segments
    keys = np.concatenate([np.full(8192, 5 << 10, dtype=np.uint32), np.full(16384, 6 << 10, dtype=np.uint32), np.array([7 << 10, (9 << 10) | 3], dtype=np.uint32)])
    _gpu_bucket_sort(keys, 21, 8192)


def test_bucket_sort_full_size_and_refusals():
    # the headline's entry count: 2^24 points x 12 windows of c = 22 -> 2^21 buckets
    rng = np.random.default_rng(5)
    n = 12 << 24
    keys = rng.integers(0, 1 << 21, size=n, dtype=np.uint32)
    _gpu_bucket_sort(keys, 21)
    lib = zkb.lib()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    k = np.zeros(4, dtype=np.uint32)
    assert lib.zkb_bucket_sort_pairs(k.ctypes.data_as(u32p), k.ctypes.data_as(u32p), 4, 25, 0) != 0      # too wide for two levels
    assert lib.zkb_bucket_sort_pairs(k.ctypes.data_as(u32p), k.ctypes.data_as(u32p), 4, 10, 100) != 0    # tile not a multiple of the CTA
    assert lib.zkb_bucket_sort_pairs(None, k.ctypes.data_as(u32p), 4, 10, 0) != 0
