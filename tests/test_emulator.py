"""CPU emulation of the device code (csrc/hostemu.cu runs the kernels' per-thread functions) vs the oracle.

This is how limb arithmetic, NTT index maths and MSM chunk/partial logic are checked where no GPU exists; the
`-m gpu` tests repeat the comparisons on the real kernels through the C ABI.
"""
import ctypes

import numpy as np
import pytest

import emu
from oracle import pyref as R
from util import FQ_LIMBS, FR_LIMBS, int_to_limbs, ints_to_limbs, limbs_to_int, random_field


def mont(xs):
    return ints_to_limbs([R.to_mont(x, R.FR) for x in xs])


@pytest.mark.parametrize("field,mod", [("fr", FR_LIMBS), ("fq", FQ_LIMBS)])
def test_field_ops(oracle, field, mod):
    a = random_field(5000, 1, mod)
    b = random_field(5000, 2, mod)
    m = limbs_to_int(mod)
    a[0] = 0; b[1] = 0
    a[2] = int_to_limbs(m - 1); b[2] = a[2]
    a[3] = int_to_limbs(1); b[3] = int_to_limbs(m - 1)
    a[4] = int_to_limbs(m - 1); b[4] = int_to_limbs(1)
    for op in ("mul", "add", "sub"):
        assert (emu.vec_op(field, op, a, b) == oracle.vec_op(field, op, a, b)).all(), (field, op)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15])
def test_ntt_matches_best_fft(oracle, k):
    a = random_field(1 << k, 10 + k)
    w = oracle.fr_omega(k)
    assert (emu.ntt(a, k, w) == oracle.best_fft(a, w, k)).all()


def test_ntt_three_passes(oracle):
    k = 19
    a = random_field(1 << k, 19)
    w = oracle.fr_omega(k)
    assert (emu.ntt(a, k, w) == oracle.best_fft(a, w, k)).all()


def test_ntt_inverse_omega_and_delta(oracle):
    k = 11
    w = oracle.fr_omega(k)
    wi = oracle.fr_inv(w)
    a = random_field(1 << k, 3)
    assert (emu.ntt(a, k, wi) == oracle.best_fft(a, wi, k)).all()
    delta = np.zeros((1 << k, 4), dtype=np.uint64)
    delta[0] = mont([1])[0]
    assert (emu.ntt(delta, k, w) == np.tile(mont([1])[0], (1 << k, 1))).all()


@pytest.mark.parametrize("k,ek", [(3, 5), (5, 7), (9, 11), (10, 12), (12, 13), (12, 12)])
def test_coset_fusions(oracle, k, ek):
    z = R.FR_ZETA
    zeta3 = mont([1, z, z * z % R.FR])
    a = random_field(1 << k, 50 + k)
    ext = emu.ntt(a, ek, oracle.fr_omega(ek), in_scale3=zeta3)
    assert (ext == oracle.coeff_to_extended(a, k, ek)).all()
    ninv = pow(1 << ek, -1, R.FR)
    wi = mont([pow(R.omega_for(ek), -1, R.FR)])[0]
    outs = mont([ninv, ninv * z * z % R.FR, ninv * z % R.FR])
    back = emu.ntt(ext, ek, wi, out_scale3=outs)
    assert (back == oracle.extended_to_coeff(ext, k, ek)).all()
    assert (back[: 1 << k] == a).all() and not back[1 << k:].any()


def test_ntt_batched_columns(oracle):
    k, cols = 12, 3
    a = random_field(cols << k, 77).reshape(cols, -1, 4)
    w = oracle.fr_omega(k)
    got = emu.ntt(a, k, w, cols=cols)
    for i in range(cols):
        assert (got[i] == oracle.best_fft(a[i], w, k)).all()


def _bases(oracle, n, seed):
    return oracle.g1_fixed_base_mul(random_field(n, seed))


def test_fixed_base_mul(oracle):
    s = random_field(6, 3)
    s[0] = 0
    s[1] = mont([1])[0]
    s[2] = mont([R.FR - 1])[0]
    assert (emu.fixed_base_mul(s) == oracle.g1_fixed_base_mul(s)).all()


@pytest.mark.parametrize("n,c,chunk", [(1, 0, 0), (2, 0, 0), (3, 0, 0), (17, 0, 0), (64, 0, 0), (300, 0, 0),
                                       (1000, 4, 8), (1000, 7, 3), (1000, 13, 128), (2000, 0, 0)])
def test_msm_matches_best_multiexp(oracle, n, c, chunk):
    s = random_field(n, n + c)
    b = _bases(oracle, n, 1000 + n)
    assert (emu.msm(s, b, c, chunk) == oracle.best_multiexp(s, b)).all()


@pytest.mark.parametrize("n,c,chunk", [(700, 6, 3), (3000, 9, 16), (257, 4, 1)])
def test_msm_level0_throughput_form(oracle, n, c, chunk):
    """Level 0 without the in-CTA tree (what large MSMs run): thread summaries go straight to the partial list and the next
    levels combine them — same result as the tree form and as the oracle."""
    s = random_field(n, 31 + n)
    b = oracle.g1_fixed_base_mul(random_field(n, 32 + n))
    s[: n // 3] = s[0]
    want = oracle.best_multiexp(s, b)
    emu.lib().zkb_emu_msm_set_direct0(1)
    try:
        assert (emu.msm(s, b, c=c, chunk=chunk) == want).all()
        assert (emu.msm(s, b, c=c, chunk=chunk, table=True) == want).all()
    finally:
        emu.lib().zkb_emu_msm_set_direct0(0)
    assert (emu.msm(s, b, c=c, chunk=chunk) == want).all()


@pytest.mark.parametrize("n,c", [(700, 6), (3000, 9), (64, 4), (300, 12)])
def test_msm_segment_sum_both_forms(oracle, n, c):
    """The bucket reduction's segment sum as one thread per segment (throughput form) and as four lanes per segment in lock-step
    phases (lane-cooperative additions and doublings, latency form): both against the oracle, repeated bases included (doubling)."""
    s = random_field(n, 41 + n)
    b = oracle.g1_fixed_base_mul(random_field(n, 42 + n))
    b[1] = b[0]
    s[1] = s[0]
    want = oracle.best_multiexp(s, b)
    for coop in (0, 1):
        emu.lib().zkb_emu_msm_set_reduce_coop(coop)
        try:
            assert (emu.msm(s, b, c=c) == want).all()
            assert (emu.msm(s, b, c=c, table=True) == want).all()
        finally:
            emu.lib().zkb_emu_msm_set_reduce_coop(1)


def test_msm_edge_cases(oracle):
    n = 300
    b = _bases(oracle, n, 5)
    # all-equal scalars: every level-0 chunk is a single run of one bucket per window
    s = random_field(n, 6)
    s[:] = s[0]
    assert (emu.msm(s, b, 5, 4) == oracle.best_multiexp(s, b)).all()
    # all-equal scalars and bases: exercises the doubling branch of the mixed add
    b2 = np.tile(b[0], (n, 1))
    assert (emu.msm(s, b2, 6, 5) == oracle.best_multiexp(s, b2)).all()
    # zeros, identity base, r-1
    s = random_field(n, 7)
    s[::2] = 0
    s[7] = mont([R.FR - 1])[0]
    b3 = b.copy()
    b3[5] = 0
    assert (emu.msm(s, b3, 8, 16) == oracle.best_multiexp(s, b3)).all()
    # cancelling pair -> identity partial sums
    s = random_field(64, 8)
    b4 = b[:64].copy()
    y = limbs_to_int(b4[0][4:])
    b4[1][:4] = b4[0][:4]
    b4[1][4:] = int_to_limbs(R.FQ - y)
    s[1] = s[0]
    assert (emu.msm(s, b4, 4, 2) == oracle.best_multiexp(s, b4)).all()
    # everything cancels: result is the identity, returned as (0, R, 0)
    s2 = np.stack([s[0], s[0]])
    out = emu.msm(s2, b4[:2], 4, 2)
    assert not out[:4].any() and not out[8:].any() and limbs_to_int(out[4:8]) == R.FQ_R
    # all-zero scalars
    out = emu.msm(np.zeros((10, 4), np.uint64), b[:10])
    assert not out[8:].any()


def test_msm_witness_like_distribution(oracle):
    """50% zero, 25% < 2^16, 20% < 2^88, 5% uniform (SURVEY.md §8d distribution W)."""
    n = 1500
    rng = np.random.default_rng(9)
    vals = []
    for i in range(n):
        u = rng.random()
        if u < 0.5: vals.append(0)
        elif u < 0.75: vals.append(int(rng.integers(0, 1 << 16)))
        elif u < 0.95: vals.append(int.from_bytes(rng.bytes(11), "little"))
        else: vals.append(int.from_bytes(rng.bytes(31), "little") % R.FR)
    s = mont(vals)
    b = _bases(oracle, n, 11)
    assert (emu.msm(s, b) == oracle.best_multiexp(s, b)).all()


@pytest.mark.parametrize("n,ncols,c,chunk,table,dev_final", [(50, 3, 0, 0, False, False), (50, 3, 4, 4, True, False),
                                                             (128, 9, 0, 0, True, True), (128, 9, 5, 7, False, True),
                                                             (1, 2, 0, 0, True, False)])
def test_msm_batched_columns_and_table(oracle, n, ncols, c, chunk, table, dev_final):
    """Several scalar columns against the same bases in one pass (column folded into the bucket key), with and
    without the SRS window table, host and device finalisation."""
    b = _bases(oracle, n, n + 5)
    cols = np.stack([random_field(n, 100 + i) for i in range(ncols)])
    cols[0, :] = 0
    got = emu.msm_batch(cols, b, c, chunk, table, dev_final)
    for i in range(ncols):
        assert (got[i] == oracle.best_multiexp(cols[i], b)).all()


@pytest.mark.parametrize("n,c,chunk", [(5, 3, 2), (300, 5, 8), (700, 0, 0)])
def test_msm_srs_window_table(oracle, n, c, chunk):
    s = random_field(n, n)
    b = _bases(oracle, n, n + 5)
    if n >= 64:
        s[3] = 0; b[7] = 0; s[9] = s[8]; b[9] = b[8]
    assert (emu.msm(s, b, c, chunk, table=True) == oracle.best_multiexp(s, b)).all()


@pytest.mark.parametrize("n,slices,c,chunk", [(300, 4, 5, 8), (301, 3, 6, 4), (700, 4, 0, 0), (64, 7, 4, 2)])
def test_msm_point_range_slices_share_buckets(oracle, n, slices, c, chunk):
    """The pipelined host commit: slices of the point range are digit-decomposed / sorted / accumulated one after the
    other, later slices ADD into the bucket array (rmw), one reduction at the end."""
    s = random_field(n, 3 * n)
    b = _bases(oracle, n, n + 9)
    s[5] = 0; b[11] = 0; s[n - 1] = s[0]; b[n - 1] = b[0]   # same point and scalar in the first and last slice
    s[n // 2:] = s[n // 2]                                   # a long single-bucket run inside later slices
    assert (emu.msm_sliced(s, b, slices, c, chunk) == oracle.best_multiexp(s, b)).all()


# ---- distributed NTT: one transform sharded over G ranks (peer memory emulated by per-rank arrays) ---------------------
@pytest.mark.parametrize("k,log_g", [(11, 1), (12, 2), (13, 3), (16, 1), (17, 3), (19, 2), (20, 3)])
def test_ntt_dist_matches_best_fft(oracle, k, log_g):
    a = random_field(1 << k, 700 + k)
    w = oracle.fr_omega(k)
    rc, got = emu.ntt_dist(a, k, w, log_g)
    assert rc == 0
    assert (got == oracle.best_fft(a, w, k)).all()


def test_ntt_dist_rejects_single_pass():
    a = random_field(1 << 8, 1)
    rc, _ = emu.ntt_dist(a, 8, np.zeros(4, dtype=np.uint64), 1)
    assert rc == -2


# ---- setup.cuh: windowed fixed-base multiplication, batched normalisation, SRS scalars -------------------------------------------
def test_fixed_base_window_matches_oracle(oracle):
    s = random_field(40, 77)
    s[0] = 0
    s[1] = mont([1])[0]
    s[2] = mont([R.FR - 1])[0]
    s[3] = mont([1 << 240])[0]          # only the top window
    s[4] = mont([0xFFFF])[0]            # the last entry of the first window
    assert (emu.fixed_base_window(s) == oracle.g1_fixed_base_mul(s)).all()


def test_batch_normalize_matches_oracle(oracle):
    n = 37  # not a multiple of the per-thread batch
    aff = oracle.g1_fixed_base_mul(random_field(n, 5))
    z = random_field(n, 6, FQ_LIMBS)
    jac = np.zeros((n, 12), dtype=np.uint64)
    z2 = oracle.vec_op("fq", "mul", z, z)
    jac[:, :4] = oracle.vec_op("fq", "mul", aff[:, :4], z2)
    jac[:, 4:8] = oracle.vec_op("fq", "mul", aff[:, 4:], oracle.vec_op("fq", "mul", z2, z))
    jac[:, 8:] = z
    jac[7, 8:] = 0       # identities inside a batch
    jac[16, 8:] = 0
    want = oracle.g1_batch_normalize(jac)
    assert not want[7].any() and (want[0] == aff[0]).all()
    assert (emu.batch_normalize(jac) == want).all()


@pytest.mark.parametrize("k", [1, 5, 7])
def test_setup_scalars_by_definition(k):
    n = 1 << k
    s = 0x0123456789ABCDEF0F1E2D3C4B5A6978 + k
    sm = mont([s])[0]
    rc, pw = emu.setup_scalars(0, n, sm)
    assert rc == 0 and [R.from_mont(limbs_to_int(x), R.FR) for x in pw] == [pow(s, i, R.FR) for i in range(n)]
    w = R.omega_for(k)
    c = (pow(s, n, R.FR) - 1) * pow(n, R.FR - 2, R.FR) % R.FR
    rc, lg = emu.setup_scalars(1, n, sm, mont([w])[0], mont([c])[0])
    want = [pow(w, i, R.FR) * c % R.FR * pow((s - pow(w, i, R.FR)) % R.FR, R.FR - 2, R.FR) % R.FR for i in range(n)]
    assert rc == 0 and [R.from_mont(limbs_to_int(x), R.FR) for x in lg] == want
    assert sum(want) % R.FR == 1
    # s inside the domain: reported (upstream panics on the failed inversion)
    rc, _ = emu.setup_scalars(1, n, mont([pow(w, n - 1, R.FR)])[0], mont([w])[0], mont([c])[0])
    assert rc == 1


# ---- poly.cuh: eval_polynomial, kate_division, batch inversion (multi-level chunking) ----------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4096, 4097, 70000])
def test_poly_eval_and_kate_division(oracle, n):
    a = random_field(n, 500 + n)
    x = random_field(1, 501 + n)[0]
    assert (emu.poly_eval(a, x) == oracle.fr_eval_polynomial(a, x)).all()
    if n >= 2:
        assert (emu.kate_division(a, x) == oracle.fr_kate_division(a, x)).all()


def test_batch_invert(oracle):
    a = random_field(100, 9)
    a[0] = 0
    a[31] = 0
    a[32] = 0
    a[99] = mont([1])[0]
    assert (emu.batch_invert(a) == oracle.fr_batch_invert(a)).all()


# ---- property-based differential tests (hypothesis): device limb arithmetic vs Python integers, edge-biased inputs -----------------
try:
    from hypothesis import given, settings, strategies as st

    def _elems(mod):
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 128) - 1, (1 << 128), (1 << 64) - 1, (1 << 32) - 1, mod >> 1]
        return st.one_of(st.sampled_from(edge), st.integers(min_value=0, max_value=mod - 1))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.tuples(_elems(R.FR), _elems(R.FR)), min_size=1, max_size=8))
    def test_fr_ops_property(pairs):
        a = ints_to_limbs([x for x, _ in pairs])
        b = ints_to_limbs([y for _, y in pairs])
        rinv = pow(1 << 256, R.FR - 2, R.FR)
        assert limbs_to_ints_(emu.vec_op("fr", "mul", a, b)) == [x * y * rinv % R.FR for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op("fr", "add", a, b)) == [(x + y) % R.FR for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op("fr", "sub", a, b)) == [(x - y) % R.FR for x, y in pairs]

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.tuples(_elems(R.FQ), _elems(R.FQ)), min_size=1, max_size=8))
    def test_fq_ops_property(pairs):
        a = ints_to_limbs([x for x, _ in pairs])
        b = ints_to_limbs([y for _, y in pairs])
        rinv = pow(1 << 256, R.FQ - 2, R.FQ)
        assert limbs_to_ints_(emu.vec_op("fq", "mul", a, b)) == [x * y * rinv % R.FQ for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op("fq", "add", a, b)) == [(x + y) % R.FQ for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op("fq", "sub", a, b)) == [(x - y) % R.FQ for x, y in pairs]

    @settings(max_examples=40, deadline=None)
    @given(st.integers(min_value=1, max_value=9), st.integers(min_value=0, max_value=2**32 - 1))
    def test_ntt_property_inverse_roundtrip(k, seed):
        """NTT(omega) then NTT(omega^-1) returns n * a for random sizes / seeds (checked with Python integers)."""
        n = 1 << k
        a = random_field(n, seed)
        w = R.omega_for(k)
        f = emu.ntt(a, k, mont([w])[0])
        b = emu.ntt(f, k, mont([pow(w, R.FR - 2, R.FR)])[0])
        want = [R.to_mont(n * R.from_mont(limbs_to_int(x), R.FR) % R.FR, R.FR) for x in a]
        assert limbs_to_ints_(b) == want

    def limbs_to_ints_(a):
        return [limbs_to_int(r) for r in np.asarray(a).reshape(-1, 4)]
except ImportError:  # hypothesis is part of the image; keep the module importable without it
    pass


@pytest.mark.parametrize("k", [0, 1, 3, 4])
def test_g1_fft_matches_definition(oracle, k):
    n = 1 << k
    pts = oracle.g1_fixed_base_mul(random_field(n, 60 + k))
    if n >= 8:
        pts[3] = 0           # identity input
        pts[5] = pts[4]      # repeated point
    w = oracle.fr_omega(k)
    assert (emu.g1_fft(pts, k, w) == oracle.g1_fft_naive(pts, w)).all()


def test_g_to_lagrange_equals_direct_lagrange_srs(oracle):
    """g_to_lagrange(g) = (1/n) iFFT_G1(g) must reproduce the g_lagrange that setup computes from s directly."""
    k = 3
    s = random_field(1, 88)[0]
    g, gl = oracle.kzg_setup(k, s)
    w_inv = oracle.fr_inv(oracle.fr_omega(k))
    n_inv = oracle.fr_inv(mont([1 << k])[0])
    assert (emu.g1_fft(g, k, w_inv, n_inv) == gl).all()


# ---- lazy field arithmetic (values only < 2M): the forms used inside the bucket accumulation and the NTT butterflies -------------
try:
    from hypothesis import given, settings, strategies as st

    def _lazy_elems(mod):
        edge = [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod - 2, (1 << 254), (1 << 254) - 1, mod >> 1, mod + (mod >> 1)]
        return st.one_of(st.sampled_from([e for e in edge if e < 2 * mod]), st.integers(min_value=0, max_value=2 * mod - 1))

    @settings(max_examples=400, deadline=None)
    @given(st.sampled_from(["fr", "fq"]), st.lists(st.tuples(st.integers(0, 10**9), st.integers(0, 10**9)), min_size=1, max_size=1),
           st.data())
    def test_lazy_field_ops_property(field, _unused, data):
        mod = R.FR if field == "fr" else R.FQ
        pairs = data.draw(st.lists(st.tuples(_lazy_elems(mod), _lazy_elems(mod)), min_size=1, max_size=8))
        a = ints_to_limbs([x for x, _ in pairs])
        b = ints_to_limbs([y for _, y in pairs])
        rinv = pow(1 << 256, mod - 2, mod)
        assert limbs_to_ints_(emu.vec_op(field, "mul_lazy", a, b)) == [x * y * rinv % mod for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op(field, "add_lazy", a, b)) == [(x + y) % mod for x, y in pairs]
        assert limbs_to_ints_(emu.vec_op(field, "sub_lazy", a, b)) == [(x - y) % mod for x, y in pairs]
        raw = limbs_to_ints_(emu.vec_op(field, "mul_lazy_raw", a, b))
        assert all(v < 2 * mod and v % mod == x * y * rinv % mod for v, (x, y) in zip(raw, pairs))   # the lazy invariant
        assert [int(v) for v in emu.vec_op(field, "is_zero_lazy", a, b)[:, 0]] == [1 if x % mod == 0 else 0 for x, _ in pairs]
except ImportError:
    pass


@pytest.mark.parametrize("n", [1, 2, 64, 65, 4097])
def test_prefix_product_by_definition(n):
    a = random_field(n, 800 + n)
    vals = [R.from_mont(limbs_to_int(x), R.FR) for x in a]
    want, acc = [], 1
    for v in vals:
        want.append(acc)
        acc = acc * v % R.FR
    got = [R.from_mont(limbs_to_int(x), R.FR) for x in emu.prefix_product(a)]
    assert got == want


# ---- tests/golden/widened_kats.json through the device code's per-thread functions --------------------------------------------
import json  # noqa: E402
import os  # noqa: E402

WGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "widened_kats.json")))


def _h(x):
    return int(x, 16)


def _unmont(a):
    return [R.from_mont(limbs_to_int(r), R.FR) for r in np.asarray(a).reshape(-1, 4)]


def _aff(p):
    return np.array(R.g1_affine_encode((_h(p[0]), _h(p[1])) if p else None), dtype=np.uint64)


def test_widened_golden_vectors_through_device_code():
    c = WGOLD["poly"]
    a, b = mont([_h(x) for x in c["coeffs"]]), mont([_h(c["point"])])[0]
    assert _unmont(emu.poly_eval(a, b)) == [_h(c["eval"])]
    assert _unmont(emu.kate_division(a, b)) == [_h(x) for x in c["kate_quotient"]]
    c = WGOLD["batch_invert"]
    assert _unmont(emu.batch_invert(mont([_h(x) for x in c["in"]]))) == [_h(x) for x in c["out"]]
    c = WGOLD["prefix_product"]
    assert _unmont(emu.prefix_product(mont([_h(x) for x in c["in"]]))) == [_h(x) for x in c["out"]]
    c = WGOLD["batch_normalize"]
    jac = np.array([sum((R.fq_encode(_h(x)) for x in row), []) for row in c["jacobian"]], dtype=np.uint64)
    assert (emu.batch_normalize(jac) == np.array([_aff(p) for p in c["affine"]])).all()
    c = WGOLD["g1_fft"]
    got = emu.g1_fft(np.array([_aff(p) for p in c["in"]]), c["k"], mont([_h(c["omega"])])[0])
    assert (got == np.array([_aff(p) for p in c["out"]])).all()
    c = WGOLD["kzg_setup"]
    s = _h(c["s"])
    n = 1 << c["k"]
    rc, pw = emu.setup_scalars(0, n, mont([s])[0])
    assert rc == 0 and (emu.fixed_base_window(pw) == np.array([_aff(p) for p in c["g"]])).all()


# ---- quotient evaluation: the lowered program (graph_plan.hpp) run by the device's per-thread interpreter (graph.cuh) -----------
import graph_cases as GC  # noqa: E402


def _emu_graph(c, cta_threads=0):
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    return emu.graph_evaluate(c["graph"], fx, ad, ins, ch, None, None, None, y, c["rot_scale"], prev, cta_threads)


@pytest.mark.parametrize("i", range(3))
def test_graph_evaluate_golden(i):
    c = GC.golden_cases()[i]
    rc, out, _ = _emu_graph(c)
    assert rc == 0 and GC.unmont(out) == c["expected"]


@pytest.mark.parametrize("seed,isize,rot_scale,threads", [(21, 1, 1, 0), (22, 4, 1, 32), (23, 256, 4, 64), (24, 512, 8, 0), (25, 128, 2, 128)])
def test_graph_evaluate_matches_oracle(oracle, seed, isize, rot_scale, threads):
    c = GC.random_case(seed, isize, rot_scale, ngates=5, depth=5)
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, rot_scale, prev)
    rc, out, info = _emu_graph(c, threads)
    assert rc == 0 and (out == want).all()
    assert info[0] < len(g.calculations) and info[1] < g.num_intermediates   # Stores propagated; liveness: fewer slots than intermediates


def test_graph_slot_allocation_and_errors(oracle):
    ev = GC.ev
    V = ev.ValueSource
    col = random_field(8, 5)
    # a long chain t_{i+1} = t_i^2 + col needs two slots however long it is
    g = ev.GraphEvaluator()
    t = g.add_calculation(ev.STORE, V(ev.ADVICE, 0, g.add_rotation(0)))
    for _ in range(40):
        t = g.add_calculation(ev.ADD, g.add_calculation(ev.SQUARE, t), V(ev.ADVICE, 0, 0))
    rc, out, info = emu.graph_evaluate(g, [], [col], [], None, None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))
    assert rc == 0 and info[1] <= 2 and info[2] == 1
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, [], [col], [], None,
                                 None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))
    assert (out == want).all()
    # the same column passed under two kinds is read through one pointer
    g2 = ev.GraphEvaluator()
    g2.add_calculation(ev.MUL, V(ev.ADVICE, 0, g2.add_rotation(0)), V(ev.FIXED, 0, g2.add_rotation(-1)))
    rc, out, info = emu.graph_evaluate(g2, [col], [col], [], None, None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))
    assert rc == 0 and info[2] == 1 and (out == oracle.vec_op("fr", "mul", col, np.roll(col, 1, axis=0))).all()
    # an empty graph writes zeros
    rc, out, _ = emu.graph_evaluate(ev.GraphEvaluator(), [], [], [], None, None, None, None, None, 1, random_field(8, 6))
    assert rc == 0 and not out.any()
    # malformed: intermediate read before it is written, column out of range, scalar missing, non-power-of-two domain
    bad = ev.GraphEvaluator()
    bad.num_intermediates = 2
    bad.calculations.append((ev.ADD, 0, V(ev.INTERMEDIATE, 1), V(ev.CONSTANT, 1), V(ev.CONSTANT, 0)))
    assert emu.graph_evaluate(bad, [], [], [], None, None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))[0] != 0
    assert emu.graph_evaluate(g2, [], [col], [], None, None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))[0] != 0
    g3 = ev.GraphEvaluator()
    g3.add_calculation(ev.STORE, V(ev.Y))
    assert emu.graph_evaluate(g3, [], [], [], None, None, None, None, None, 1, np.zeros((8, 4), dtype=np.uint64))[0] != 0
    assert emu.graph_evaluate(g, [], [col[:6]], [], None, None, None, None, None, 1, np.zeros((6, 4), dtype=np.uint64))[0] != 0


def test_graph_lookup_terms_through_device_code():
    g, cols, sc, prev, want = GC.lookup_case(92, 128, 4)
    rc, out, info = emu.graph_evaluate(g, [GC.mont(c) for c in cols["fixed"]], [GC.mont(c) for c in cols["advice"]], [], None,
                                       GC.mont([sc["beta"]])[0], GC.mont([sc["gamma"]])[0], GC.mont([sc["theta"]])[0],
                                       GC.mont([sc["y"]])[0], 4, GC.mont(prev))
    assert rc == 0 and GC.unmont(out) == want and info[1] <= 6


@pytest.mark.parametrize("seed", range(24))
def test_graph_lowering_random_programs(oracle, seed):
    """Raw calculation lists that upstream would never emit — intermediates re-written, Stores of intermediates, dead values,
    Horner steps in odd places — through copy propagation, SSA renaming, dead-value removal, scheduling and slot allocation,
    against the oracle's direct evaluation (one array cell per intermediate, no lowering)."""
    import random
    g, rs = GC.random_program(1000 + seed)
    isize = 32
    fixed, advice = [random_field(isize, 3000 + seed)], [random_field(isize, 3100 + seed + i) for i in range(2)]
    sc = random_field(4, 3200 + seed)
    prev = random_field(isize, 3300 + seed)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fixed, advice, [], None,
                                 sc[0], sc[1], sc[2], sc[3], rs, prev)
    rc, out, info = emu.graph_evaluate(g, fixed, advice, [], None, sc[0], sc[1], sc[2], sc[3], rs, prev, random.Random(seed).choice([0, 32, 64]))
    assert rc == 0 and (out == want).all()
    assert info[0] <= len(g.calculations)


@pytest.mark.parametrize("world,seed", [(1, 51), (2, 52), (4, 53), (8, 54)])
def test_graph_row_windows_match_whole_domain(oracle, world, seed):
    """Row-window mode (the per-rank call of the row-sharded quotient evaluation): the domain cut into `world` shards, every
    column padded with the neighbouring shards' rows (cyclically), each window evaluated on its own — the concatenation equals
    the whole-domain evaluation; a rotation that reaches outside the halo is refused."""
    isize, rs = 256, 4
    c = GC.random_case(seed, isize, rs, ngates=4, depth=5)
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, rs, prev)
    lo, hi = g.rotation_span(rs)
    assert lo == 2 * rs and hi == 3 * rs          # rotations -2 .. 3
    rows = isize // world
    out = np.zeros_like(prev)
    for r in range(world):
        idx = np.arange(r * rows - lo, (r + 1) * rows + hi) % isize
        win = [[col[idx] for col in grp] for grp in (fx, ad, ins)]
        rc, o, _ = emu.graph_evaluate(g, win[0], win[1], win[2], ch, None, None, None, y, rs, prev[r * rows:(r + 1) * rows], halo=(lo, hi))
        assert rc == 0
        out[r * rows:(r + 1) * rows] = o
    assert (out == want).all()
    idx = np.arange(-lo, rows + hi - 1) % isize   # one row short at the top
    win = [[col[idx] for col in grp] for grp in (fx, ad, ins)]
    assert emu.graph_evaluate(g, win[0], win[1], win[2], ch, None, None, None, y, rs, prev[:rows], halo=(lo, hi - 1))[0] != 0


def test_eip196_public_vectors_through_device_code():
    """EIP-196 precompile vectors (tests/golden/eip196_kats.json) through the device MSM code path on the emulator."""
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eip196_kats.json")))
    pt = lambda p: np.array(R.g1_affine_encode((int(p[0], 16), int(p[1], 16))), dtype=np.uint64)  # noqa: E731
    one = mont([1])[0]
    for c in kat["add"]:
        out = emu.msm(np.array([one, one]), np.array([pt(c["a"]), pt(c["b"])]))
        assert R.g1_jacobian_decode([int(x) for x in out]) == (int(c["sum"][0], 16), int(c["sum"][1], 16))
    for c in kat["mul"]:
        out = emu.msm(mont([int(c["s"], 16) % R.FR]), np.array([pt(c["p"])]))
        assert R.g1_jacobian_decode([int(x) for x in out]) == (int(c["out"][0], 16), int(c["out"][1], 16))


@pytest.mark.parametrize("ncols,chunk_len,isize", [(5, 2, 64), (1, 2, 16), (4, 2, 32), (7, 3, 32)])
def test_permutation_argument_graph(oracle, ncols, chunk_len, isize):
    """evaluation.permutation_graph — every h(X) term of the permutation argument (first / last set, set linking at the last
    usable row, the two products with the running power of delta) — through the oracle and the device code, against the
    formulas with Python integers."""
    g, cols, sc, prev, want = GC.permutation_case(60 + ncols, isize, 4, ncols=ncols, chunk_len=chunk_len)
    fx, ad = [GC.mont(c) for c in cols["fixed"]], [GC.mont(c) for c in cols["advice"]]
    b, gm, y = (GC.mont([sc[k]])[0] for k in ("beta", "gamma", "y"))
    got = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, [], None, b, gm, None, y, 4, GC.mont(prev))
    assert GC.unmont(got) == want
    rc, out, info = emu.graph_evaluate(g, fx, ad, [], None, b, gm, None, y, 4, GC.mont(prev))
    assert rc == 0 and (out == got).all()


def test_lookup_and_custom_gate_builders_match_the_worked_examples(oracle):
    """evaluation.lookup_graph / custom_gates_graph build the graphs the worked examples in graph_cases build by hand."""
    ev = GC.ev
    g1, cols, sc, prev, want = GC.lookup_case(95, 64, 4)
    inputs = [("advice", 0, 0), ("prod", ("advice", 1, 0), ("advice", 0, 1))]
    table = [("fixed", 3, 0), ("scaled", ("fixed", 3, -1), 3)]
    g2 = ev.lookup_graph(inputs, table, ("fixed", 0), ("fixed", 1), ("fixed", 2), ("advice", 2), ("advice", 3), ("advice", 4))
    rc, out, _ = emu.graph_evaluate(g2, [GC.mont(c) for c in cols["fixed"]], [GC.mont(c) for c in cols["advice"]], [], None,
                                    GC.mont([sc["beta"]])[0], GC.mont([sc["gamma"]])[0], GC.mont([sc["theta"]])[0], GC.mont([sc["y"]])[0],
                                    4, GC.mont(prev))
    assert rc == 0 and GC.unmont(out) == want
    gates = [GC.halo2_base_gate(0, 0), GC.halo2_base_gate(1, 1)]
    assert ev.custom_gates_graph(gates).calculations == GC.build_custom_gates(gates).calculations


def test_quotient_of_a_satisfied_circuit_is_a_polynomial(oracle):
    """The prover's reason for evaluate_h, end to end on the CPU: a witness that satisfies halo2-base's gate on every row of the
    domain makes the numerator q (a + b c - d) vanish there, so its coset evaluations — advice and selector through
    lagrange_to_coeff and coeff_to_extended (oracle), the gate through the device interpreter with rot_scale = 2^(ek - k)
    (emulator) — divided by X^n - 1 are the coset evaluations of a polynomial of degree < 2n: extended_to_coeff returns zeros
    from coefficient 2n on.  One wrong witness cell and the high coefficients are no longer zero."""
    k, ek = 5, 7
    n = 1 << k
    g = GC.build_custom_gates([GC.halo2_base_gate()])
    tinv = GC.vanishing_inverse_on_coset(k, ek)
    for broken in (None, 9):
        w, q = GC.satisfied_gate_witness(k, 321, break_cell=broken)
        W = oracle.coeff_to_extended(oracle.lagrange_to_coeff(GC.mont(w), k), k, ek)
        Q = oracle.coeff_to_extended(oracle.lagrange_to_coeff(GC.mont(q), k), k, ek)
        rc, num, _ = emu.graph_evaluate(g, [Q], [W], [], None, None, None, None, GC.mont([1])[0], 1 << (ek - k), np.zeros((1 << ek, 4), dtype=np.uint64))
        assert rc == 0
        h_evals = GC.mont([a * b % R.FR for a, b in zip(GC.unmont(num), tinv)])
        h = GC.unmont(oracle.extended_to_coeff(h_evals, k, ek))
        assert len(h) >= 3 * n
        if broken is None:
            assert any(h[:2 * n]) and not any(h[2 * n:])
        else:
            assert any(h[2 * n:])


def test_quotient_evaluator_evaluate_h_against_the_formulas(oracle):
    """evaluation.QuotientEvaluator.evaluate_h — custom gates, then the permutation argument, then a lookup, each folded with y —
    with the per-graph call run by the emulator, against h computed row by row from the formulas with Python integers and
    explicit column data (no graphs, no column-layout code)."""
    import random
    ev = GC.ev
    P = R.FR
    rnd = random.Random(2024)
    isize, rs, chunk_len, last_rot = 64, 4, 2, -3
    col = lambda: [rnd.randrange(P) for _ in range(isize)]  # noqa: E731
    fixed, advice, instance = [col(), col()], [col(), col(), col()], [col()]
    l0, l_last, l_active, xc = col(), col(), col(), col()
    perm_cols = [("advice", 0), ("advice", 2), ("instance", 0)]
    sigmas, zs = [col() for _ in perm_cols], [col(), col()]
    lz, la, ls = col(), col(), col()
    gates = [GC.halo2_base_gate(1, 0), ("sum", ("prod", ("advice", 0, -1), ("fixed", 1, 0)), ("neg", ("instance", 0, 2)))]
    l_in, l_tab = [("advice", 1, 0), ("advice", 2, 1)], [("fixed", 0, 0), ("fixed", 1, -1)]
    y, beta, gamma, theta = (rnd.randrange(P) for _ in range(4))
    data = {"fixed": fixed, "advice": advice, "instance": instance}
    want = []
    for i in range(isize):
        at = lambda c, r=0: c[(i + r * rs) % isize]  # noqa: E731
        v = GC.custom_gates_value(gates, data, [], y, 0, i, rs, isize)
        pv = [data[k][j] for k, j in perm_cols]
        terms = [(1 - at(zs[0])) * at(l0), (at(zs[1]) ** 2 - at(zs[1])) * at(l_last), (at(zs[1]) - at(zs[0], last_rot)) * at(l0)]
        j = 0
        for s in range(2):
            left, right = at(zs[s], 1), at(zs[s])
            for c in pv[s * chunk_len:(s + 1) * chunk_len]:
                left = left * (at(c) + beta * at(sigmas[j]) + gamma) % P
                right = right * (at(c) + pow(ev.DELTA, j, P) * beta * at(xc) + gamma) % P
                j += 1
            terms.append((left - right) * at(l_active))
        A = (at(advice[1]) * theta + at(advice[2], 1)) % P
        S = (at(fixed[0]) * theta + at(fixed[1], -1)) % P
        d = at(la) - at(ls)
        terms += [(1 - at(lz)) * at(l0), (at(lz) ** 2 - at(lz)) * at(l_last),
                  (at(lz, 1) * (at(la) + beta) * (at(ls) + gamma) - at(lz) * (A + beta) * (S + gamma)) * at(l_active),
                  d * at(l0), d * (at(la) - at(la, -1)) * at(l_active)]
        for t in terms:
            v = (v * y + t) % P
        want.append(v)

    m = GC.mont
    sc = {k: m([val])[0] for k, val in dict(y=y, beta=beta, gamma=gamma, theta=theta).items()}
    state = {"values": np.zeros((isize, 4), dtype=np.uint64)}

    def evaluate(graph, values, f, a, i):   # the per-graph call on the emulator (values is the running array)
        rc, out, _ = emu.graph_evaluate(graph, f, a, i, None, sc["beta"], sc["gamma"], sc["theta"], sc["y"], rs, state["values"])
        assert rc == 0
        state["values"] = out

    q = ev.QuotientEvaluator(gates, dict(columns=perm_cols, chunk_len=chunk_len, last_rotation=last_rot), [(l_in, l_tab)])
    q.evaluate_h(None, [m(c) for c in fixed], [m(c) for c in advice], [m(c) for c in instance], None, sc["y"], sc["beta"], sc["gamma"],
                 sc["theta"], rs, m(l0), m(l_last), m(l_active), m(xc), [m(c) for c in sigmas], [m(c) for c in zs], [(m(lz), m(la), m(ls))],
                 evaluate=evaluate)
    assert GC.unmont(state["values"]) == want


@pytest.mark.parametrize("seed", range(12))
def test_graph_row_windows_random_programs(oracle, seed):
    """Row-window mode on the raw random programs (re-written intermediates, dead values, every scalar source, the previous
    value): 2 or 4 shards with cyclic halos reproduce the oracle's whole-domain evaluation."""
    g, rs = GC.random_program(5000 + seed)
    isize, world = 64, (2, 4)[seed % 2]
    fixed, advice = [random_field(isize, 6000 + seed)], [random_field(isize, 6100 + seed + i) for i in range(2)]
    sc = random_field(4, 6200 + seed)
    prev = random_field(isize, 6300 + seed)
    want = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fixed, advice, [], None,
                                 sc[0], sc[1], sc[2], sc[3], rs, prev)
    lo, hi = g.rotation_span(rs)
    rows = isize // world
    for r in range(world):
        idx = np.arange(r * rows - lo, (r + 1) * rows + hi) % isize
        rc, o, _ = emu.graph_evaluate(g, [fixed[0][idx]], [a[idx] for a in advice], [], None, sc[0], sc[1], sc[2], sc[3], rs,
                                      prev[r * rows:(r + 1) * rows], halo=(lo, hi))
        assert rc == 0 and (o == want[r * rows:(r + 1) * rows]).all()


@pytest.mark.parametrize("u,distinct,big", [(1, 1, False), (9, 4, False), (300, 300, True), (1500, 33, True), (4000, 256, False), (9000, 700, True), (40000, 5000, False)])
def test_permute_expression_pair_device_phases_vs_oracle(oracle, u, distinct, big):
    """lookup.cuh's per-thread phases (canonical copies, the block-wise bitonic + merge-path sort of (value, row) records,
    first-occurrence flags, table matching by binary search, rank scans, leftover assignment) against the oracle's restatement of
    upstream's algorithm.  Sizes above 1024 rows exercise the merge passes (9000 rows: four of them)."""
    from test_oracle import lookup_case
    inp, tab = lookup_case(7 * u + distinct, u, distinct, big)
    a = ints_to_limbs([R.to_mont(x, R.FR) for x in inp])
    t = ints_to_limbs([R.to_mont(x, R.FR) for x in tab])
    want = oracle.permute_expression_pair(a, t, u)
    got = emu.permute_expression_pair(a, t, u)
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()


def test_permute_expression_pair_random_shapes(oracle):
    """Random row counts (around the 1024-record block and tile boundaries too), numbers of distinct values and value widths."""
    from hypothesis import given, settings, strategies as st
    from test_oracle import lookup_case

    @settings(max_examples=25, deadline=None)
    @given(u=st.one_of(st.integers(1, 5000), st.sampled_from([1023, 1024, 1025, 2047, 2048, 2049, 4096, 4097])), frac=st.floats(0.001, 1.0),
           big=st.booleans(), seed=st.integers(0, 10**6))
    def run(u, frac, big, seed):
        distinct = max(1, min(u, int(u * frac)))
        inp, tab = lookup_case(seed, u, distinct, big)
        a = ints_to_limbs([R.to_mont(x, R.FR) for x in inp])
        t = ints_to_limbs([R.to_mont(x, R.FR) for x in tab])
        want = oracle.permute_expression_pair(a, t, u)
        got = emu.permute_expression_pair(a, t, u)
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all()

    run()


def test_permute_expression_pair_missing_value_emulated(oracle):
    a = ints_to_limbs([R.to_mont(x, R.FR) for x in (1, 2, 9)])
    t = ints_to_limbs([R.to_mont(x, R.FR) for x in (1, 2, 3)])
    with pytest.raises(ValueError):
        emu.permute_expression_pair(a, t, 3)


# ---- csrc/bucket_sort.cuh: the MSM's own most-significant-digit-first partition -------------------------------------------------------
def _check_bucket_sort(keys, key_bits, tile=0):
    keys = np.asarray(keys, dtype=np.uint32)
    vals = np.arange(keys.size, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(12345)   # distinct payloads
    got = emu.bucket_sort(keys, vals, key_bits, tile)
    assert got is not None
    gk, gv = got
    assert (gk == np.sort(keys)).all(), "keys not sorted"
    # every payload still sits next to its own key, nothing lost or duplicated (order inside a bucket is free)
    back = dict(zip(vals.tolist(), keys.tolist()))
    assert len(set(gv.tolist())) == keys.size
    assert all(back[v] == k for k, v in zip(gk.tolist(), gv.tolist()))


@pytest.mark.parametrize("n,key_bits,tile", [(1, 1, 0), (5, 3, 0), (1000, 7, 0), (5000, 12, 0), (5000, 13, 0), (40000, 16, 1024),
                                            (40000, 21, 1024), (70000, 24, 8192), (60000, 23, 0), (9000, 10, 8192), (20000, 15, 3072)])
def test_bucket_sort_uniform_keys(n, key_bits, tile):
    rng = np.random.default_rng(n + key_bits)
    _check_bucket_sort(rng.integers(0, 1 << key_bits, size=n, dtype=np.uint32), key_bits, tile)


def test_bucket_sort_skewed_and_degenerate_keys():
    rng = np.random.default_rng(7)
    n = 30000
    # witness-like: most entries in a few buckets
    keys = np.where(rng.random(n) < 0.8, rng.integers(0, 8, size=n), rng.integers(0, 1 << 18, size=n)).astype(np.uint32)
    _check_bucket_sort(keys, 18, 1024)
    _check_bucket_sort(np.full(n, 0x2ABCD, dtype=np.uint32), 18, 2048)            # the all-equal column: one bucket
    _check_bucket_sort(np.full(n, (1 << 18) - 1, dtype=np.uint32), 18, 1024)       # last bucket of the last segment
    _check_bucket_sort(np.zeros(3, dtype=np.uint32), 22, 0)
    _check_bucket_sort(np.arange(n, dtype=np.uint32)[::-1] % (1 << 14), 14, 2048)  # descending, every bucket hit
    # segment sizes that are exact multiples of the tile, and segments of one entry
    keys = np.concatenate([np.full(1024, 5 << 9, dtype=np.uint32), np.full(2048, 6 << 9, dtype=np.uint32), np.array([7 << 9, 9 << 9 | 3], dtype=np.uint32)])
    _check_bucket_sort(keys, 18, 1024)


def test_bucket_sort_key_too_wide_is_refused():
    assert emu.bucket_sort(np.arange(10, dtype=np.uint32), np.arange(10, dtype=np.uint32), 25) is None


def test_host_fold_64bit_arithmetic_matches_the_portable_chains():
    """csrc/host_fq64.hpp (the product's per-commit host fold: Horner combination, sum of partials, normalisation — 4 x 64-bit limbs)
    against the portable twins of the device code on random operands and the edge cases (0, 1, p - 1, identity / equal / opposite points)."""
    fn = emu.lib().zkb_emu_host64_check
    fn.restype = ctypes.c_uint64
    assert fn(ctypes.c_uint64(1), ctypes.c_uint64(300)) == 0
    assert fn(ctypes.c_uint64(0xDEADBEEF), ctypes.c_uint64(50)) == 0


def test_bucket_sort_random_geometries():
    """Random sizes, key widths, tile sizes and key distributions (hypothesis): sorted keys, every payload still next to its key."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(1, 30000), key_bits=st.integers(1, 24), tile_k=st.integers(0, 8), skew=st.integers(0, 3), seed=st.integers(0, 2**31))
    def run(n, key_bits, tile_k, skew, seed):
        rng = np.random.default_rng(seed)
        hi = 1 << key_bits
        if skew == 0:
            keys = rng.integers(0, hi, size=n)
        elif skew == 1:     # a few heavy buckets
            keys = np.where(rng.random(n) < 0.9, rng.integers(0, min(hi, 4), size=n), rng.integers(0, hi, size=n))
        elif skew == 2:     # one bucket
            keys = np.full(n, int(rng.integers(0, hi)))
        else:               # only the top of the range
            keys = hi - 1 - rng.integers(0, min(hi, 7), size=n)
        _check_bucket_sort(keys.astype(np.uint32), key_bits, 1024 * tile_k)

    run()
