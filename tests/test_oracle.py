"""Pins the CPU oracle (oracle/zkb_oracle.c + oracle/pyref.py) before anything trusts it.

The reference holds no golden vectors for the MSM/NTT path: these are first-principles
known-answer values plus cross-checks between two independent restatements (C limbs vs Python big-ints).
"""
import json
import os

import numpy as np
import pytest

from oracle import pyref as R
from util import FQ_LIMBS, int_to_limbs, ints_to_limbs, limbs_to_int, limbs_to_ints, random_field

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hotpath_kats.json")))


def fr_mont(xs):
    return ints_to_limbs([R.to_mont(x, R.FR) for x in xs])


def fr_unmont(a):
    return [R.from_mont(x, R.FR) for x in limbs_to_ints(a)]


def aff_mont(P):
    return np.array(R.g1_affine_encode(P), dtype=np.uint64)


def test_constants_match_survey():
    # SURVEY.md §8 [COMPUTED] constants
    assert R.FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert R.FR_ROOT_OF_UNITY_INV == 0x048127174DAABC261BBE587180F34361B22625F59115ABA70ED3E50A414E6DBA
    assert R.FR_ZETA == 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
    assert R.FR_R == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert R.FR_R2 == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert R.FQ_R == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert R.FQ_R2 == 0x06D89F71CAB8351F47AB1EFF0A417FF6B5E71911D44501FBF32CFC5B538AFA89
    assert R.FR_INV64 == 0xC2E1F593EFFFFFFF and R.FQ_INV64 == 0x87D20782E4866389
    assert R.omega_for(13) == 0x10E3D295C1599FF535A1BB49F23D81AA03BD0ED25881F9ED12B179AF67F67AE1
    assert R.omega_for(22) == 0x18C95F1AE6514E11A1B30FD7923947C5FFCEC5347F16E91B4DD654168326BEDE
    assert R.omega_for(24) == 0x1951441010B2B95A6E47A6075066A50A036F5BA978C050F2821DF86636C0FACB
    assert pow(R.FR_ZETA, 3, R.FR) == 1 and R.FR_ZETA != 1
    assert pow(R.FR_ROOT_OF_UNITY, 1 << 28, R.FR) == 1 and pow(R.FR_ROOT_OF_UNITY, 1 << 27, R.FR) != 1


def test_c_field_ops_match_bigint(oracle):
    for field, m in (("fr", R.FR), ("fq", R.FQ)):
        mod = FQ_LIMBS if field == "fq" else None
        a = random_field(500, 1, *([mod] if mod is not None else []))
        b = random_field(500, 2, *([mod] if mod is not None else []))
        # edge rows
        a[0] = 0; b[0] = 0
        a[1] = int_to_limbs(m - 1); b[1] = int_to_limbs(m - 1)
        a[2] = int_to_limbs(1); b[2] = int_to_limbs(m - 1)
        ai, bi = limbs_to_ints(a), limbs_to_ints(b)
        rinv = pow(R.R256, -1, m)
        assert limbs_to_ints(oracle.vec_op(field, "mul", a, b)) == [x * y * rinv % m for x, y in zip(ai, bi)]
        assert limbs_to_ints(oracle.vec_op(field, "add", a, b)) == [(x + y) % m for x, y in zip(ai, bi)]
        assert limbs_to_ints(oracle.vec_op(field, "sub", a, b)) == [(x - y) % m for x, y in zip(ai, bi)]


def test_c_omega_and_montgomery_layout(oracle):
    # in-memory Montgomery limbs of the SURVEY KAT out[1]
    for k in (3, 13, 22, 24):
        assert limbs_to_int(oracle.fr_omega(k)) == R.to_mont(R.omega_for(k), R.FR)
    g = oracle.g1_generator()
    assert limbs_to_int(g[:4]) == R.FQ_R and limbs_to_int(g[4:]) == 2 * R.FQ_R % R.FQ


@pytest.mark.parametrize("case", GOLD["best_fft"], ids=lambda c: f"k{c['k']}")
def test_best_fft_golden(oracle, case):
    a = [int(x, 16) for x in case["in"]]
    want = [int(x, 16) for x in case["out"]]
    w = int(case["omega"], 16)
    # python twin
    b = list(a)
    R.best_fft(b, w, case["k"])
    assert b == want
    # C oracle
    got = oracle.best_fft(fr_mont(a), fr_mont([w])[0], case["k"])
    assert fr_unmont(got) == want


def test_best_fft_survey_kat_memory_limbs(oracle):
    a = fr_mont(range(1, 9))
    out = oracle.best_fft(a, oracle.fr_omega(3), 3)
    assert [int(x) for x in out[1]] == [0x1069F4287460CB5F, 0xEA22DD8C9B017FC5, 0xC9CDFB2B2395711E, 0x2758DB28A5C09FDD]
    assert fr_unmont(out)[0] == 0x24 and fr_unmont(out)[4] == R.FR - 4
    assert fr_unmont(out)[7] == 0x303D4CCDE3F282EB3E1FA8DA0EB98F60726A9D586FEE27AA5ECD945D75D0E863


@pytest.mark.parametrize("case", GOLD["lagrange_to_coeff"], ids=lambda c: f"k{c['k']}")
def test_lagrange_to_coeff_golden(oracle, case):
    a = [int(x, 16) for x in case["in"]]
    want = [int(x, 16) for x in case["out"]]
    assert R.EvaluationDomain(4, case["k"]).lagrange_to_coeff(a) == want
    assert fr_unmont(oracle.lagrange_to_coeff(fr_mont(a), case["k"])) == want
    assert fr_unmont(oracle.coeff_to_lagrange(fr_mont(want), case["k"])) == a


@pytest.mark.parametrize("case", GOLD["coeff_to_extended"], ids=lambda c: f"j{c['j']}k{c['k']}")
def test_coeff_to_extended_golden(oracle, case):
    a = [int(x, 16) for x in case["in"]]
    want = [int(x, 16) for x in case["out"]]
    d = R.EvaluationDomain(case["j"], case["k"])
    assert d.extended_k == case["extended_k"]
    assert d.coeff_to_extended(a) == want
    got = oracle.coeff_to_extended(fr_mont(a), case["k"], d.extended_k)
    assert fr_unmont(got) == want
    back = fr_unmont(oracle.extended_to_coeff(got, case["k"], d.extended_k))
    assert back[: len(a)] == a and all(x == 0 for x in back[len(a):])
    assert d.extended_to_coeff(want) == (a + [0] * len(want))[: d.n * d.quotient_poly_degree]


def test_coeff_to_extended_survey_kat():
    e = R.EvaluationDomain(4, 2).coeff_to_extended([1, 2, 3, 4])
    assert e[0] == 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90E0
    assert e[1] == 0x02784A0A8A0E97B1E4BDF0FF35FF9CC7E84CE82C0568A0F5022C8379EABAD181
    assert e[2] == 0x0F474B2DC63AC314F5DDFFAA966E250CBC434908CCF08F98880600F34B714E16


def test_fft_c_vs_python_random_and_threads(oracle):
    for k in (0, 1, 7, 10):
        a = random_field(1 << k, 100 + k)
        w = oracle.fr_omega(k)
        ai = fr_unmont(a)
        R.best_fft(ai, R.omega_for(k), k)
        for threads in (1, 3):
            assert fr_unmont(oracle.best_fft(a, w, k, threads)) == ai


@pytest.mark.parametrize("case", GOLD["g1_mul"], ids=lambda c: c["s"][:12])
def test_g1_mul_golden(oracle, case):
    s = int(case["s"], 16)
    want = tuple(int(x, 16) for x in case["out"]) if case["out"] else None
    assert R.g1_mul(R.G1_GENERATOR, s) == want
    got = oracle.g1_mul(oracle.g1_generator(), fr_mont([s])[0])
    assert R.g1_affine_decode([int(x) for x in got]) == want
    assert oracle.g1_is_on_curve(got)


@pytest.mark.parametrize("case", GOLD["msm"], ids=lambda c: c["name"])
def test_msm_golden(oracle, case):
    scal = [int(x, 16) for x in case["scalars"]]
    bases = [tuple(int(c, 16) for c in b) if b else None for b in case["bases"]]
    want = tuple(int(x, 16) for x in case["out"])
    assert R.best_multiexp(scal, bases, 3) == want
    s = fr_mont(scal)
    b = np.stack([aff_mont(P) for P in bases])
    for threads in (1, 2, 8):
        got = oracle.best_multiexp(s, b, threads)
        assert R.g1_jacobian_decode([int(x) for x in got]) == want
        assert limbs_to_int(got[8:12]) == R.FQ_R  # z normalised to Montgomery one
    assert R.g1_jacobian_decode([int(x) for x in oracle.msm_naive(s, b)]) == want


def test_msm_known_dlog_property(oracle):
    """bases b_i*G  =>  MSM == (sum s_i b_i) * G  (the check used at full size on the GPU)."""
    n = 300
    s = random_field(n, 7)
    b = random_field(n, 8)
    bases = oracle.g1_fixed_base_mul(b)
    assert all(oracle.g1_is_on_curve(p) for p in bases[:10])
    got = oracle.g1_to_affine(oracle.best_multiexp(s, bases))
    ip = oracle.fr_inner_product(s, b)  # Montgomery(sum s_i*b_i / R) -> need true product: use to_mont of one side
    # inner product of Montgomery forms: mont_mul(sR, bR) = s*b*R  => sum is Montgomery form of sum s_i b_i
    want = oracle.g1_mul(oracle.g1_generator(), ip)
    assert (got == want).all()
    # and against python
    si, bi = fr_unmont(s), fr_unmont(b)
    tot = sum(x * y for x, y in zip(si, bi)) % R.FR
    assert R.g1_affine_decode([int(x) for x in got]) == R.g1_mul(R.G1_GENERATOR, tot)


def test_msm_edge_cases(oracle):
    g = oracle.g1_generator()
    # empty
    out = oracle.best_multiexp(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64))
    assert limbs_to_int(out[8:12]) == 0
    # all zero scalars
    out = oracle.best_multiexp(np.zeros((5, 4), np.uint64), np.tile(g, (5, 1)))
    assert limbs_to_int(out[8:12]) == 0
    # [r-1]G + [1]G = identity
    s = fr_mont([R.FR - 1, 1])
    out = oracle.best_multiexp(s, np.tile(g, (2, 1)))
    assert limbs_to_int(out[8:12]) == 0
    # all-equal scalars and bases
    s = fr_mont([12345] * 64)
    out = oracle.best_multiexp(s, np.tile(g, (64, 1)))
    assert R.g1_jacobian_decode([int(x) for x in out]) == R.g1_mul(R.G1_GENERATOR, 12345 * 64)


# ---- ParamsKZG::setup (G1 side), batch_normalize, eval_polynomial: oracle vs first principles ---------------------------------
def test_kzg_setup_oracle_by_definition(oracle):
    """g[i] = [s^i]G and g_lagrange[i] = [l_i(s)]G with l_i(s) = omega^i (s^n - 1) / (n (s - omega^i)), recomputed with
    Python integers; the Lagrange basis sums to one, so sum_i g_lagrange[i] = G; a commitment equals [p(s)]G in both bases."""
    k, n = 3, 8
    s = 0x1234567890ABCDEF1122334455667788
    g, gl = oracle.kzg_setup(k, ints_to_limbs([R.to_mont(s, R.FR)])[0])
    G = (1, 2)
    w = R.omega_for(k)
    c = (pow(s, n, R.FR) - 1) * pow(n, R.FR - 2, R.FR) % R.FR
    for i in range(n):
        assert R.g1_affine_decode([int(x) for x in g[i]]) == R.g1_mul(G, pow(s, i, R.FR))
        li = pow(w, i, R.FR) * c % R.FR * pow((s - pow(w, i, R.FR)) % R.FR, R.FR - 2, R.FR) % R.FR
        assert R.g1_affine_decode([int(x) for x in gl[i]]) == R.g1_mul(G, li)
    acc = None
    for i in range(n):
        acc = R.g1_add(acc, R.g1_affine_decode([int(x) for x in gl[i]]))
    assert acc == G
    # KZG consistency: commit(coeffs) = [p(s)]G = commit_lagrange(evaluations over the domain)
    coeffs = [3, 1, 4, 1, 5, 9, 2, 6]
    ps = sum(cf * pow(s, i, R.FR) for i, cf in enumerate(coeffs)) % R.FR
    cm = ints_to_limbs([R.to_mont(x, R.FR) for x in coeffs])
    assert R.g1_jacobian_decode([int(x) for x in oracle.best_multiexp(cm, g)]) == R.g1_mul(G, ps)
    evals = oracle.coeff_to_lagrange(cm, k)
    assert (oracle.best_multiexp(evals, gl) == oracle.best_multiexp(cm, g)).all()
    assert R.from_mont(limbs_to_int(oracle.fr_eval_polynomial(cm, ints_to_limbs([R.to_mont(s, R.FR)])[0])), R.FR) == ps


def test_batch_normalize_oracle(oracle):
    pts = []
    want = []
    for i, z in enumerate([1, 5, 0, 0xABCDEF, R.FQ - 1]):
        P = R.g1_mul((1, 2), 7 + 13 * i)
        if z == 0:
            pts.append([0] * 4 + R.fq_encode(1) + [0] * 4)
            want.append(None)
            continue
        x, y = P
        pts.append(R.fq_encode(x * z * z % R.FQ) + R.fq_encode(y * z * z * z % R.FQ) + R.fq_encode(z))
        want.append(P)
    got = oracle.g1_batch_normalize(np.array(pts, dtype=np.uint64))
    assert [R.g1_affine_decode([int(v) for v in row]) for row in got] == want


def test_kate_division_and_batch_invert_oracle(oracle):
    """kate_division by the definition a(X) = q(X) (X - b) + a(b) with Python integers; BatchInvert leaves zeros."""
    a = [5, 0, 7, 11, 13, 2, 9]
    b = 0xDEADBEEF12345
    q = [R.from_mont(limbs_to_int(x), R.FR) for x in oracle.fr_kate_division(ints_to_limbs([R.to_mont(x, R.FR) for x in a]),
                                                                              ints_to_limbs([R.to_mont(b, R.FR)])[0])]
    assert len(q) == len(a) - 1
    rem = sum(c * pow(b, i, R.FR) for i, c in enumerate(a)) % R.FR
    prod = [0] * len(a)
    for i, c in enumerate(q):            # q(X) * (X - b)
        prod[i + 1] = (prod[i + 1] + c) % R.FR
        prod[i] = (prod[i] - c * b) % R.FR
    prod[0] = (prod[0] + rem) % R.FR
    assert prod == [x % R.FR for x in a]
    v = [3, 0, R.FR - 1, 12345]
    inv = [R.from_mont(limbs_to_int(x), R.FR) for x in oracle.fr_batch_invert(ints_to_limbs([R.to_mont(x, R.FR) for x in v]))]
    assert inv == [pow(3, R.FR - 2, R.FR), 0, R.FR - 1, pow(12345, R.FR - 2, R.FR)]


# ---- tests/golden/widened_kats.json: the widened rows against vectors made from Python integers by the definitions ------------
WGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "widened_kats.json")))


def _h(x):
    return int(x, 16)


def _aff_or_id(p):
    return aff_mont((_h(p[0]), _h(p[1])) if p else None)


def test_widened_golden_kzg_setup(oracle):
    c = WGOLD["kzg_setup"]
    g, gl = oracle.kzg_setup(c["k"], fr_mont([_h(c["s"])])[0])
    assert (g == np.array([_aff_or_id(p) for p in c["g"]])).all()
    assert (gl == np.array([_aff_or_id(p) for p in c["g_lagrange"]])).all()


def test_widened_golden_poly(oracle):
    c = WGOLD["poly"]
    a, b = fr_mont([_h(x) for x in c["coeffs"]]), fr_mont([_h(c["point"])])[0]
    assert fr_unmont(oracle.fr_eval_polynomial(a, b)) == [_h(c["eval"])]
    assert fr_unmont(oracle.fr_kate_division(a, b)) == [_h(x) for x in c["kate_quotient"]]


def test_widened_golden_batch_invert_and_scan(oracle):
    c = WGOLD["batch_invert"]
    assert fr_unmont(oracle.fr_batch_invert(fr_mont([_h(x) for x in c["in"]]))) == [_h(x) for x in c["out"]]
    # the scan has no oracle entry point of its own: z[i+1] = z[i] * v[i] with the oracle's product
    c = WGOLD["prefix_product"]
    v, z = fr_mont([_h(x) for x in c["in"]]), fr_mont([_h(x) for x in c["out"]])
    assert _h(c["out"][0]) == 1 and (oracle.vec_op("fr", "mul", z[:-1], v[:-1]) == z[1:]).all()


def test_widened_golden_batch_normalize(oracle):
    c = WGOLD["batch_normalize"]
    jac = np.array([sum((R.fq_encode(_h(x)) for x in row), []) for row in c["jacobian"]], dtype=np.uint64)
    assert (oracle.g1_batch_normalize(jac) == np.array([_aff_or_id(p) for p in c["affine"]])).all()


def test_widened_golden_g1_fft(oracle):
    c = WGOLD["g1_fft"]
    pts = np.array([_aff_or_id(p) for p in c["in"]])
    got = oracle.g1_fft_naive(pts, fr_mont([_h(c["omega"])])[0])
    assert (got == np.array([_aff_or_id(p) for p in c["out"]])).all()


# ---- quotient evaluation (GraphEvaluator::evaluate in evaluate_h's row loop): oracle vs the expressions' definition ------------
import graph_cases as GC  # noqa: E402


def _oracle_graph(oracle, c, threads=0):
    g = c["graph"]
    fx, ad, ins, ch, y, prev = GC.case_arrays(c)
    return oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, fx, ad, ins, ch,
                                 None, None, None, y, c["rot_scale"], prev, threads)


@pytest.mark.parametrize("i", range(3))
def test_graph_evaluate_golden(oracle, i):
    c = GC.golden_cases()[i]
    assert GC.unmont(_oracle_graph(oracle, c)) == c["expected"]


@pytest.mark.parametrize("seed,isize,rot_scale", [(11, 1, 1), (12, 2, 1), (13, 64, 4), (14, 128, 2), (15, 32, 8)])
def test_graph_evaluate_oracle_vs_definition(oracle, seed, isize, rot_scale):
    c = GC.random_case(seed, isize, rot_scale, ngates=4, depth=5)
    assert GC.unmont(_oracle_graph(oracle, c, threads=3)) == GC.case_expected(c)


def test_graph_compile_rules():
    """add_expression's special cases (upstream evaluation.rs): zero / one / two, a + (-b) -> Sub, x * x -> Square,
    identical calculations are shared."""
    ev = GC.ev
    g = ev.GraphEvaluator()
    a, b = ("advice", 0, 0), ("advice", 0, 1)
    assert g.add_expression(("prod", ("const", 0), a)) == ev.ValueSource(ev.CONSTANT, 0)
    ra = g.add_expression(a)
    assert g.add_expression(("prod", ("const", 1), a)) == ra and g.add_expression(("sum", a, ("const", 0))) == ra
    n = len(g.calculations)
    assert g.add_expression(a) == ra and len(g.calculations) == n            # de-duplicated
    g.add_expression(("prod", ("const", 2), a))
    assert g.calculations[-1][0] == ev.DOUBLE
    g.add_expression(("prod", a, a))
    assert g.calculations[-1][0] == ev.SQUARE
    g.add_expression(("sum", a, ("neg", b)))
    assert g.calculations[-1][0] == ev.SUB and g.calculations[-1][2] == ra
    g.add_expression(("sum", ("const", 0), ("neg", b)))
    assert g.calculations[-1][0] == ev.NEGATE
    assert g.add_expression(("scaled", a, 1)) == ra and g.add_expression(("scaled", a, 0)) == ev.ValueSource(ev.CONSTANT, 0)
    assert g.add_expression(("neg", ("const", 5))) == ev.ValueSource(ev.CONSTANT, g.constants.index(R.FR - 5))
    assert g.rotations == [0, 1] and g.constants[:3] == [0, 1, 2]


def test_graph_permutation_term_oracle_vs_definition(oracle):
    """The permutation argument's product term as a graph over beta / gamma / fixed / advice sources, against the formula."""
    import random
    rnd = random.Random(77)
    ncols, isize, rs = 3, 32, 4
    g, delta = GC.permutation_term_graph(ncols)
    fixed = [[rnd.randrange(R.FR) for _ in range(isize)] for _ in range(2 + ncols)]
    advice = [[rnd.randrange(R.FR) for _ in range(isize)] for _ in range(1 + ncols)]
    beta, gamma = rnd.randrange(R.FR), rnd.randrange(R.FR)
    got = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations, [GC.mont(c) for c in fixed],
                                [GC.mont(c) for c in advice], [], None, GC.mont([beta])[0], GC.mont([gamma])[0], None, None, rs,
                                np.zeros((isize, 4), dtype=np.uint64))
    want = []
    for i in range(isize):
        left, right = advice[0][(i + rs) % isize], advice[0][i]
        for j in range(ncols):
            left = left * (advice[1 + j][i] + beta * fixed[2 + j][i] + gamma) % R.FR
            right = right * (advice[1 + j][i] + pow(delta, j, R.FR) * beta * fixed[1][i] + gamma) % R.FR
        want.append((left - right) * fixed[0][i] % R.FR)
    assert GC.unmont(got) == want


def test_graph_lookup_terms_oracle_vs_definition(oracle):
    """All five h(X) terms of a lookup argument (theta-compressed input / table, beta, gamma, rotations +1 and -1) as one graph."""
    g, cols, sc, prev, want = GC.lookup_case(91, 64, 4)
    got = oracle.graph_evaluate(g.calc_array(), g.num_intermediates, GC.mont(g.constants), g.rotations,
                                [GC.mont(c) for c in cols["fixed"]], [GC.mont(c) for c in cols["advice"]], [], None,
                                GC.mont([sc["beta"]])[0], GC.mont([sc["gamma"]])[0], GC.mont([sc["theta"]])[0], GC.mont([sc["y"]])[0],
                                4, GC.mont(prev))
    assert GC.unmont(got) == want


# ---- external public vectors: the EIP-196 (alt_bn128) precompile tests, tests/golden/eip196_kats.json ---------------------------
EIP196 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eip196_kats.json")))


def _pt(p):
    return (int(p[0], 16), int(p[1], 16))


def test_eip196_vectors_pin_both_oracles(oracle):
    """bn256Add / bn256ScalarMul 'chfast' vectors: the big-int twin, the C oracle's group law, its scalar multiplication and
    best_multiexp (as a 2-point and a 1-point MSM) all reproduce the published outputs."""
    one = fr_mont([1])[0]
    for c in EIP196["add"]:
        a, b, s = _pt(c["a"]), _pt(c["b"]), _pt(c["sum"])
        assert R.g1_add(a, b) == s
        assert R.g1_affine_decode([int(x) for x in oracle.g1_add_affine(aff_mont(a), aff_mont(b))]) == s
        msm = oracle.best_multiexp(np.array([one, one]), np.array([aff_mont(a), aff_mont(b)]))
        assert R.g1_jacobian_decode([int(x) for x in msm]) == s
    for c in EIP196["mul"]:
        p, k, out = _pt(c["p"]), int(c["s"], 16) % R.FR, _pt(c["out"])
        assert R.g1_mul(p, k) == out
        km = fr_mont([k])[0]
        assert R.g1_affine_decode([int(x) for x in oracle.g1_mul(aff_mont(p), km)]) == out
        assert R.g1_jacobian_decode([int(x) for x in oracle.best_multiexp(np.array([km]), np.array([aff_mont(p)]))]) == out


def test_published_halo2curves_montgomery_constants(oracle):
    """The Montgomery constants published in halo2curves' bn256 fr.rs / fq.rs (INV, R, R2, R3 — written down from memory, the
    crate is not vendored): they are 2^256, 2^512, 2^768 mod the modulus and -m^-1 mod 2^64, i.e. the in-memory layout this
    repository assumes (a * 2^256 mod m in four little-endian u64 limbs) is the crate's; the C oracle's to-Montgomery conversion
    of 1 gives the same R."""
    fr = dict(inv=0xc2e1f593efffffff,
              R=[0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f],
              R2=[0x1bb8e645ae216da7, 0x53fe3ab1e35c59e3, 0x8c49833d53bb8085, 0x0216d0b17f4e44a5],
              R3=[0x5e94d8e1b4bf0040, 0x2a489cbe1cfbb6b8, 0x893cc664a19fcfed, 0x0cf8594b7fcc657c])
    fq = dict(inv=0x87d20782e4866389,
              R=[0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f],
              R2=[0xf32cfc5b538afa89, 0xb5e71911d44501fb, 0x47ab1eff0a417ff6, 0x06d89f71cab8351f],
              R3=[0xb1cd6dafda1530df, 0x62f210e6a7283db6, 0xef7f0b0c0ada0afb, 0x20fd6e902d592544])
    for consts, m in ((fr, R.FR), (fq, R.FQ)):
        assert consts["inv"] == (-pow(m, -1, 1 << 64)) % (1 << 64)
        for name, e in (("R", 256), ("R2", 512), ("R3", 768)):
            assert limbs_to_int(consts[name]) == pow(2, e, m), name
    one = np.array([[1, 0, 0, 0]], dtype=np.uint64)
    assert [int(x) for x in oracle.fr_to_mont(one)[0]] == fr["R"]
    assert [int(x) for x in oracle.vec_op("fq", "mul", np.array([fq["R2"]], dtype=np.uint64), one)[0]] == fq["R"]   # mont(R2, 1) = R


# ---- lookup argument: permute_expression_pair ----------------------------------------------------------------------------------
def py_permute_expression_pair(inp, tab, u):
    """upstream's algorithm with Python containers (sorted list, Counter as the BTreeMap), on canonical integers"""
    from collections import Counter
    a = sorted(inp[:u])
    left = Counter(tab[:u])
    out_t, rep = [0] * u, []
    for row, v in enumerate(a):
        if row == 0 or v != a[row - 1]:
            out_t[row] = v
            if left.get(v, 0) == 0:
                raise ValueError("ConstraintSystemFailure")
            left[v] -= 1
        else:
            rep.append(row)
    for v in sorted(left):
        for _ in range(left[v]):
            out_t[rep.pop()] = v
    assert not rep
    return a, out_t


def lookup_case(seed, u, distinct, big=False):
    import random
    rnd = random.Random(seed)
    pool = [rnd.randrange(R.FR) if big else rnd.randrange(1 << 20) for _ in range(distinct)]
    tab = pool + [rnd.choice(pool) for _ in range(u - distinct)]
    rnd.shuffle(tab)
    inp = [rnd.choice(pool[: max(1, distinct // 2)]) for _ in range(u)]
    return inp, tab


@pytest.mark.parametrize("u,distinct,big", [(1, 1, False), (7, 3, False), (64, 64, True), (500, 17, True), (2000, 256, False)])
def test_permute_expression_pair_vs_python(oracle, u, distinct, big):
    inp, tab = lookup_case(u + distinct, u, distinct, big)
    want_a, want_t = py_permute_expression_pair(inp, tab, u)
    m = lambda xs: ints_to_limbs([R.to_mont(x, R.FR) for x in xs])  # noqa: E731
    got_a, got_t = oracle.permute_expression_pair(m(inp + [5, 6]), m(tab + [7, 8]), u)   # rows past usable_rows are ignored
    un = lambda a: [R.from_mont(x, R.FR) for x in limbs_to_ints(a)]  # noqa: E731
    assert un(got_a) == want_a and un(got_t) == want_t
    # the lookup argument's conditions: A'[i] == S'[i] or A'[i] == A'[i-1]; S' is a permutation of the table
    assert all(want_a[i] == want_t[i] or want_a[i] == want_a[i - 1] for i in range(u)) and want_a[0] == want_t[0]
    assert sorted(want_t) == sorted(tab[:u])


def test_permute_expression_pair_missing_value(oracle):
    m = lambda xs: ints_to_limbs([R.to_mont(x, R.FR) for x in xs])  # noqa: E731
    with pytest.raises(ValueError):
        oracle.permute_expression_pair(m([1, 2, 9]), m([1, 2, 3]), 3)
