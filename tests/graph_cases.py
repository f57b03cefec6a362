"""Quotient-evaluation test cases: halo2 `Expression`s evaluated BY THE DEFINITION with Python integers, and builders for the
graphs the three implementations (oracle, emulator, CUDA) are checked with.

`eval_expr` is the meaning of an expression at a row of the extended domain (a column query reads the row
(idx + rotation * rot_scale) mod isize); `custom_gates_value` is what halo2-axiom's evaluate_h computes for the custom gates,
value = previous; for each gate polynomial: value = value * y + gate(row).  Neither knows anything about GraphEvaluator."""
import importlib
import random

import numpy as np

from oracle import pyref as R
from util import ints_to_limbs, limbs_to_int

ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
P = R.FR


def eval_expr(e, cols, challenges, idx, rot_scale, isize):
    t = e[0]
    if t == "const":
        return e[1] % P
    if t in ("fixed", "advice", "instance"):
        return cols[t][e[1]][(idx + e[2] * rot_scale) % isize]
    if t == "challenge":
        return challenges[e[1]]
    if t == "neg":
        return -eval_expr(e[1], cols, challenges, idx, rot_scale, isize) % P
    if t == "sum":
        return (eval_expr(e[1], cols, challenges, idx, rot_scale, isize) + eval_expr(e[2], cols, challenges, idx, rot_scale, isize)) % P
    if t == "prod":
        return eval_expr(e[1], cols, challenges, idx, rot_scale, isize) * eval_expr(e[2], cols, challenges, idx, rot_scale, isize) % P
    if t == "scaled":
        return eval_expr(e[1], cols, challenges, idx, rot_scale, isize) * e[2] % P
    raise ValueError(t)


def custom_gates_value(gates, cols, challenges, y, prev, idx, rot_scale, isize):
    v = prev
    for g in gates:
        v = (v * y + eval_expr(g, cols, challenges, idx, rot_scale, isize)) % P
    return v


# halo2-base's only custom gate (FlexGateConfig): q * (a + b * c - d) with a, b, c, d = advice at rotations 0, 1, 2, 3
def halo2_base_gate(advice_col=0, selector_col=0):
    a, b, c, d = (("advice", advice_col, r) for r in range(4))
    return ("prod", ("fixed", selector_col, 0), ("sum", ("sum", a, ("prod", b, c)), ("neg", d)))


def random_expr(rnd, depth, ncols, nchal):
    if depth == 0 or rnd.random() < 0.15:
        k = rnd.randrange(6)
        if k == 0:
            return ("const", rnd.choice([0, 1, 2, P - 1, rnd.randrange(P)]))
        if k == 1 and nchal:
            return ("challenge", rnd.randrange(nchal))
        kind = rnd.choice(["fixed", "advice", "advice", "instance"])
        return (kind, rnd.randrange(ncols[kind]), rnd.choice([0, 0, 1, -1, 2, 3, -2]))
    k = rnd.randrange(8)
    if k == 0:
        return ("neg", random_expr(rnd, depth - 1, ncols, nchal))
    if k == 1:
        return ("scaled", random_expr(rnd, depth - 1, ncols, nchal), rnd.choice([0, 1, 2, rnd.randrange(P)]))
    if k in (2, 3, 4):
        return ("sum", random_expr(rnd, depth - 1, ncols, nchal), random_expr(rnd, depth - 1, ncols, nchal))
    if k == 5:
        x = random_expr(rnd, depth - 1, ncols, nchal)
        return ("prod", x, x)  # Square
    return ("prod", random_expr(rnd, depth - 1, ncols, nchal), random_expr(rnd, depth - 1, ncols, nchal))


def build_custom_gates(gates):
    """evaluate_h's `custom_gates` evaluator: every gate polynomial compiled, then Horner(PreviousValue, parts, Y)."""
    g = ev.GraphEvaluator()
    parts = [g.add_expression(e) for e in gates]
    g.add_horner(ev.ValueSource(ev.PREVIOUS), parts, ev.ValueSource(ev.Y))
    return g


def mont(xs):
    return ints_to_limbs([R.to_mont(x, P) for x in xs])


def unmont(a):
    return [R.from_mont(limbs_to_int(r), P) for r in np.asarray(a).reshape(-1, 4)]


def random_case(seed, isize, rot_scale, ngates=3, depth=4, ncols=None, nchal=2):
    """-> dict with integer columns / scalars, the gate expressions and the compiled graph"""
    rnd = random.Random(seed)
    ncols = ncols or {"fixed": 2, "advice": 3, "instance": 1}
    cols = {k: [[rnd.randrange(P) for _ in range(isize)] for _ in range(n)] for k, n in ncols.items()}
    cols["fixed"][0][1 % isize] = 0
    cols["advice"][0][0] = P - 1
    gates = [halo2_base_gate()] + [random_expr(rnd, depth, ncols, nchal) for _ in range(ngates - 1)]
    return dict(isize=isize, rot_scale=rot_scale, cols=cols, challenges=[rnd.randrange(P) for _ in range(nchal)], y=rnd.randrange(P),
                prev=[rnd.randrange(P) for _ in range(isize)], gates=gates, graph=build_custom_gates(gates))


def case_expected(c, rows=None):
    rows = range(c["isize"]) if rows is None else rows
    return [custom_gates_value(c["gates"], c["cols"], c["challenges"], c["y"], c["prev"][i], i, c["rot_scale"], c["isize"]) for i in rows]


def case_arrays(c):
    """Montgomery limb arrays of a case: (fixed, advice, instance lists, challenges, y, prev)"""
    return ([mont(col) for col in c["cols"]["fixed"]], [mont(col) for col in c["cols"]["advice"]],
            [mont(col) for col in c["cols"]["instance"]], mont(c["challenges"]), mont([c["y"]])[0], mont(c["prev"]))


def golden_cases():
    """tests/golden/graph_kats.json back as cases (gate expressions as nested tuples, integers parsed)."""
    import json
    import os

    def tup(e):
        return tuple(tup(x) if isinstance(x, list) else x for x in e)

    cases = []
    for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "graph_kats.json"))):
        gates = [tup(g) for g in c["gates"]]
        cases.append(dict(isize=c["isize"], rot_scale=c["rot_scale"], gates=gates, graph=build_custom_gates(gates),
                          cols={k: [[int(x, 16) for x in col] for col in v] for k, v in c["cols"].items()},
                          challenges=[int(x, 16) for x in c["challenges"]], y=int(c["y"], 16), prev=[int(x, 16) for x in c["prev"]],
                          expected=[int(x, 16) for x in c["expected"]]))
    return cases


def permutation_term_graph(ncols, fold=False):
    """One chunk of the permutation argument's h(X) term, written directly as calculations over value sources:
        l_active * ( z(wX) * prod_i (col_i + beta * sigma_i + gamma)  -  z(X) * prod_i (col_i + delta^i * beta * X + gamma) )
    fixed columns: 0 = l_active, 1 = coset of X, 2 .. 2+ncols-1 = sigma cosets; advice: 0 = z, 1 .. ncols = the columns;
    constants carry delta^i.  fold: the term is folded into the previous value with y.  -> (graph, delta)"""
    V = ev.ValueSource
    g = ev.GraphEvaluator()
    r0, r1 = g.add_rotation(0), g.add_rotation(1)
    delta = pow(7, 1 << 28, P)  # halo2curves Fr::DELTA = GENERATOR^(2^S)
    beta, gamma = V(ev.BETA), V(ev.GAMMA)
    left = g.add_calculation(ev.STORE, V(ev.ADVICE, 0, r1))
    right = g.add_calculation(ev.STORE, V(ev.ADVICE, 0, r0))
    bx = g.add_calculation(ev.MUL, beta, V(ev.FIXED, 1, r0))
    for i in range(ncols):
        col = V(ev.ADVICE, 1 + i, r0)
        t = g.add_calculation(ev.MUL, beta, V(ev.FIXED, 2 + i, r0))
        t = g.add_calculation(ev.ADD, g.add_calculation(ev.ADD, col, t), gamma)
        left = g.add_calculation(ev.MUL, left, t)
        u = g.add_calculation(ev.MUL, bx, g.add_constant(pow(delta, i, P)))
        u = g.add_calculation(ev.ADD, g.add_calculation(ev.ADD, col, u), gamma)
        right = g.add_calculation(ev.MUL, right, u)
    term = g.add_calculation(ev.MUL, g.add_calculation(ev.SUB, left, right), V(ev.FIXED, 0, r0))
    if fold:   # h = h * y + term, as evaluate_h folds every term into the running value
        g.add_horner(V(ev.PREVIOUS), [term], V(ev.Y))
    return g, delta


def lookup_terms_graph(input_exprs, table_exprs, z_col, a_perm_col, s_perm_col):
    """The five h(X) terms of one lookup argument as ONE graph, each folded into the previous value with y:
        l_0 (1 - z)                 l_last (z^2 - z)
        l_active ( z(wX) (A' + beta) (S' + gamma) - z(X) (A + beta) (S + gamma) )
        l_0 (A' - S')               l_active (A' - S') (A' - A'(w^-1 X))
    A, S = the input / table expressions compressed with theta (Horner), as upstream's lookup GraphEvaluator computes them.
    fixed columns: 0 = l_0, 1 = l_last, 2 = l_active; advice columns z_col, a_perm_col, s_perm_col hold z, A', S'."""
    V = ev.ValueSource
    g = ev.GraphEvaluator()
    theta, beta, gamma = V(ev.THETA), V(ev.BETA), V(ev.GAMMA)
    one = g.add_constant(1)

    def compress(exprs):
        parts = [g.add_expression(e) for e in exprs]
        return g.add_horner(g.add_constant(0), parts, theta)

    A, S = compress(input_exprs), compress(table_exprs)
    r0, r1, rm1 = g.add_rotation(0), g.add_rotation(1), g.add_rotation(-1)
    l0, llast, lact = V(ev.FIXED, 0, r0), V(ev.FIXED, 1, r0), V(ev.FIXED, 2, r0)
    z, zw = V(ev.ADVICE, z_col, r0), V(ev.ADVICE, z_col, r1)
    ap, apm, sp = V(ev.ADVICE, a_perm_col, r0), V(ev.ADVICE, a_perm_col, rm1), V(ev.ADVICE, s_perm_col, r0)
    c = g.add_calculation
    t1 = c(ev.MUL, l0, c(ev.SUB, one, z))
    t2 = c(ev.MUL, llast, c(ev.SUB, c(ev.SQUARE, z), z))
    left = c(ev.MUL, c(ev.MUL, zw, c(ev.ADD, ap, beta)), c(ev.ADD, sp, gamma))
    right = c(ev.MUL, c(ev.MUL, z, c(ev.ADD, A, beta)), c(ev.ADD, S, gamma))
    t3 = c(ev.MUL, lact, c(ev.SUB, left, right))
    d = c(ev.SUB, ap, sp)
    t4 = c(ev.MUL, l0, d)
    t5 = c(ev.MUL, lact, c(ev.MUL, d, c(ev.SUB, ap, apm)))
    g.add_horner(V(ev.PREVIOUS), [t1, t2, t3, t4, t5], V(ev.Y))
    return g


def lookup_terms_expected(input_exprs, table_exprs, z_col, a_perm_col, s_perm_col, cols, theta, beta, gamma, y, prev, rot_scale, isize):
    """The same five terms from their formulas with Python integers."""
    out = []
    for i in range(isize):
        def comp(exprs):
            v = 0
            for e in exprs:
                v = (v * theta + eval_expr(e, cols, [], i, rot_scale, isize)) % P
            return v
        A, S = comp(input_exprs), comp(table_exprs)
        adv, fx = cols["advice"], cols["fixed"]
        z, zw = adv[z_col][i], adv[z_col][(i + rot_scale) % isize]
        ap, apm, sp = adv[a_perm_col][i], adv[a_perm_col][(i - rot_scale) % isize], adv[s_perm_col][i]
        l0, llast, lact = fx[0][i], fx[1][i], fx[2][i]
        terms = [l0 * (1 - z), llast * (z * z - z),
                 lact * (zw * (ap + beta) * (sp + gamma) - z * (A + beta) * (S + gamma)),
                 l0 * (ap - sp), lact * (ap - sp) * (ap - apm)]
        v = prev[i]
        for t in terms:
            v = (v * y + t) % P
        out.append(v)
    return out


def lookup_case(seed, isize, rot_scale):
    rnd = random.Random(seed)
    cols = {"fixed": [[rnd.randrange(P) for _ in range(isize)] for _ in range(4)],      # l_0, l_last, l_active, a table column
            "advice": [[rnd.randrange(P) for _ in range(isize)] for _ in range(5)],     # two inputs, z, A', S'
            "instance": []}
    inputs = [("advice", 0, 0), ("prod", ("advice", 1, 0), ("advice", 0, 1))]
    table = [("fixed", 3, 0), ("scaled", ("fixed", 3, -1), 3)]
    sc = dict(theta=rnd.randrange(P), beta=rnd.randrange(P), gamma=rnd.randrange(P), y=rnd.randrange(P))
    prev = [rnd.randrange(P) for _ in range(isize)]
    g = lookup_terms_graph(inputs, table, 2, 3, 4)
    want = lookup_terms_expected(inputs, table, 2, 3, 4, cols, sc["theta"], sc["beta"], sc["gamma"], sc["y"], prev, rot_scale, isize)
    return g, cols, sc, prev, want


def random_program(seed, max_len=60):
    """A raw calculation list over 1 fixed + 2 advice columns and beta / gamma / theta / y / previous value: intermediates are
    re-written, Stores copy intermediates, some values are dead, Horner steps appear anywhere.  -> (graph, rot_scale)"""
    V = ev.ValueSource
    rnd = random.Random(seed)
    rs = rnd.choice([1, 2, 4])
    g = ev.GraphEvaluator()
    for r in (0, 1, -1, 3):
        g.add_rotation(r)
    for _ in range(3):
        g.add_constant(rnd.randrange(P))
    nint = rnd.randrange(2, 9)
    g.num_intermediates = nint
    written = []

    def src():
        k = rnd.randrange(10)
        if k < 4 and written:
            return V(ev.INTERMEDIATE, rnd.choice(written))
        if k < 6:
            return V(ev.ADVICE, rnd.randrange(2), rnd.randrange(4))
        if k == 6:
            return V(ev.FIXED, 0, rnd.randrange(4))
        if k == 7:
            return V(ev.CONSTANT, rnd.randrange(len(g.constants)))
        if k == 8:
            return V(rnd.choice([ev.BETA, ev.GAMMA, ev.THETA, ev.Y]))
        return V(ev.PREVIOUS)

    z = V(ev.CONSTANT, 0)
    for _ in range(rnd.randrange(1, max_len)):
        op = rnd.choice([ev.ADD, ev.SUB, ev.MUL, ev.MUL, ev.SQUARE, ev.DOUBLE, ev.NEGATE, ev.STORE, ev.STORE, ev.MUL_ADD])
        t = rnd.randrange(nint)
        a, b, c = src(), src(), src()
        if op == ev.MUL_ADD and t in written and rnd.random() < 0.7:
            a = V(ev.INTERMEDIATE, t)     # a Horner step
        g.calculations.append((op, t, a, b if op in (ev.ADD, ev.SUB, ev.MUL, ev.MUL_ADD) else z, c if op == ev.MUL_ADD else z))
        if t not in written:
            written.append(t)
    return g, rs


def permutation_case(seed, isize, rot_scale, ncols=5, chunk_len=2, last_rotation=-4):
    """The complete permutation argument (evaluation.permutation_graph) on random columns and its value from the formulas with
    Python integers.  fixed columns: 0 l_0, 1 l_last, 2 l_active, 3 the coset of X, 4.. the sigma cosets; advice columns:
    0..ncols-1 the permuted columns, then one product coset z_i per set."""
    rnd = random.Random(seed)
    nsets = (ncols + chunk_len - 1) // chunk_len
    cols = {"fixed": [[rnd.randrange(P) for _ in range(isize)] for _ in range(4 + ncols)],
            "advice": [[rnd.randrange(P) for _ in range(isize)] for _ in range(ncols + nsets)], "instance": []}
    columns = [("advice", i) for i in range(ncols)]
    sigmas = [("fixed", 4 + i) for i in range(ncols)]
    zs = [("advice", ncols + i) for i in range(nsets)]
    g = ev.permutation_graph(columns, chunk_len, last_rotation, ("fixed", 0), ("fixed", 1), ("fixed", 2), ("fixed", 3), sigmas, zs)
    sc = dict(beta=rnd.randrange(P), gamma=rnd.randrange(P), y=rnd.randrange(P))
    prev = [rnd.randrange(P) for _ in range(isize)]
    fx, adv = cols["fixed"], cols["advice"]
    want = []
    for i in range(isize):
        nxt, lst = (i + rot_scale) % isize, (i + last_rotation * rot_scale) % isize
        z = [adv[ncols + s] for s in range(nsets)]
        terms = [(1 - z[0][i]) * fx[0][i], (z[-1][i] * z[-1][i] - z[-1][i]) * fx[1][i]]
        terms += [(z[s][i] - z[s - 1][lst]) * fx[0][i] for s in range(1, nsets)]
        j = 0
        for s in range(nsets):
            left, right = z[s][nxt], z[s][i]
            for col in range(s * chunk_len, min((s + 1) * chunk_len, ncols)):
                left = left * (adv[col][i] + sc["beta"] * fx[4 + j][i] + sc["gamma"]) % P
                right = right * (adv[col][i] + pow(ev.DELTA, j, P) * sc["beta"] * fx[3][i] + sc["gamma"]) % P
                j += 1
            terms.append((left - right) * fx[2][i])
        v = prev[i]
        for t in terms:
            v = (v * sc["y"] + t) % P
        want.append(v)
    return g, cols, sc, prev, want


def satisfied_gate_witness(k, seed, break_cell=None):
    """A witness column for halo2-base's gate on n = 2^k rows: the selector is on at rows 0, 4, 8, ... and every gate
    w[i] + w[i+1] w[i+2] - w[i+3] = 0 holds (break_cell: one cell is changed afterwards).  -> (w, q) as integer lists"""
    rnd = random.Random(seed)
    n = 1 << k
    w, q = [0] * n, [0] * n
    for j in range(0, n, 4):
        a, b, c = (rnd.randrange(P) for _ in range(3))
        w[j:j + 4] = [a, b, c, (a + b * c) % P]
        q[j] = 1
    if break_cell is not None:
        w[break_cell] = (w[break_cell] + 1) % P
    return w, q


def vanishing_inverse_on_coset(k, ek):
    """1 / (X^n - 1) at the points zeta w_ext^i of the extended coset, i = 0 .. 2^ek - 1 (upstream's t_evaluations, inverted)."""
    n, N = 1 << k, 1 << ek
    zeta = pow(7, 2 * (P - 1) // 3, P)   # halo2curves bn256::Fr::ZETA, the coset generator
    w_ext = R.omega_for(ek)
    period = N // n
    vals = [pow(pow(zeta, n, P) * pow(w_ext, i * n, P) % P - 1, P - 2, P) for i in range(period)]
    return [vals[i % period] for i in range(N)]
