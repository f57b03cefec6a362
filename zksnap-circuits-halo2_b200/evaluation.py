"""Host mirror of halo2-axiom `plonk/evaluation.rs`: `GraphEvaluator` (quotient evaluation).

Upstream (un-vendored; reached from create_proof, /root/reference/aggregator/src/wrapper.rs:129-137) compiles the custom
gates and the lookup expressions into constants, rotations and calculations over value sources, then `evaluate_h` runs the
graph for every row of the extended domain.  This module keeps the same names and the same compilation rules
(`add_rotation`, `add_constant`, `add_calculation` with de-duplication, `add_expression` with its zero / one / two /
negation special cases) and `evaluate` runs the row loop on the device over polynomials resident in HBM
(`zkb_graph_evaluate`, include/zkb200.h).  Building a graph needs no GPU; evaluating it has no CPU fallback.

Expressions are tuples: ("const", int) ("fixed", col, rot) ("advice", col, rot) ("instance", col, rot) ("challenge", i)
("neg", e) ("sum", a, b) ("prod", a, b) ("scaled", e, int) — upstream's `Expression` without selectors (already substituted
when `evaluate_h` runs).  Scalars are canonical integers mod r.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple

import numpy as np

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_R = (1 << 256) % FR

# ZKB_SRC_* / ZKB_CALC_* of include/zkb200.h
CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, STORE, MUL_ADD = range(8)


class ValueSource(NamedTuple):
    kind: int
    index: int = 0
    rotation: int = 0


class CValueSource(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("index", ctypes.c_uint32), ("rotation", ctypes.c_uint32)]


class CCalculation(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint32), ("target", ctypes.c_uint32), ("a", CValueSource), ("b", CValueSource), ("c", CValueSource)]


_u64p = ctypes.POINTER(ctypes.c_uint64)


class CGraph(ctypes.Structure):
    _fields_ = [("calculations", ctypes.POINTER(CCalculation)), ("num_calculations", ctypes.c_size_t),
                ("num_intermediates", ctypes.c_uint32), ("constants", _u64p), ("num_constants", ctypes.c_size_t),
                ("rotations", ctypes.POINTER(ctypes.c_int32)), ("num_rotations", ctypes.c_size_t)]


class CGraphInputs(ctypes.Structure):
    _fields_ = [("fixed", _u64p), ("num_fixed", ctypes.c_size_t), ("advice", _u64p), ("num_advice", ctypes.c_size_t),
                ("instance", _u64p), ("num_instance", ctypes.c_size_t), ("challenges", _u64p), ("num_challenges", ctypes.c_size_t),
                ("beta", _u64p), ("gamma", _u64p), ("theta", _u64p), ("y", _u64p), ("rot_scale", ctypes.c_int32)]


def _mont_limbs(xs) -> np.ndarray:
    out = np.zeros((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        m = (x % FR) * _R % FR
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def _ptr(a: np.ndarray | None):
    return a.ctypes.data_as(_u64p) if a is not None and a.size else None


class GraphEvaluator:
    """`GraphEvaluator<C>`: constants start as [0, 1, 2] like upstream's Default."""

    def __init__(self):
        self.constants: list[int] = [0, 1, 2]
        self.rotations: list[int] = []
        self.calculations: list[tuple] = []   # (op, target, a, b, c)
        self.num_intermediates = 0
        self._existing: dict[tuple, int] = {}  # (op, a, b) -> target of the calculation add_calculation made for it

    # ---- upstream's builders ---------------------------------------------------------------------------------------
    def add_rotation(self, rotation: int) -> int:
        if rotation in self.rotations:
            return self.rotations.index(rotation)
        self.rotations.append(rotation)
        return len(self.rotations) - 1

    def add_constant(self, constant: int) -> ValueSource:
        constant %= FR
        if constant not in self.constants:
            self.constants.append(constant)
        return ValueSource(CONSTANT, self.constants.index(constant))

    def add_calculation(self, op: int, a: ValueSource, b: ValueSource | None = None) -> ValueSource:
        """An identical earlier calculation is reused (upstream's `existing_calculation`)."""
        z = ValueSource(CONSTANT, 0)
        key = (op, a, b or z)
        if key in self._existing:
            return ValueSource(INTERMEDIATE, self._existing[key])
        target = self.num_intermediates
        self.num_intermediates += 1
        self.calculations.append((op, target, a, b or z, z))
        self._existing[key] = target
        return ValueSource(INTERMEDIATE, target)

    def add_horner(self, start: ValueSource, parts: list[ValueSource], factor: ValueSource) -> ValueSource:
        """Calculation::Horner(start, parts, factor): value = start; for part: value = value * factor + part."""
        target = self.num_intermediates
        self.num_intermediates += 1
        z = ValueSource(CONSTANT, 0)
        self.calculations.append((STORE, target, start, z, z))
        for p in parts:
            self.calculations.append((MUL_ADD, target, ValueSource(INTERMEDIATE, target), factor, p))
        return ValueSource(INTERMEDIATE, target)

    def _is(self, v: ValueSource, constant: int) -> bool:
        return v.kind == CONSTANT and self.constants[v.index] == constant

    def add_expression(self, e: tuple) -> ValueSource:
        tag = e[0]
        if tag == "const":
            return self.add_constant(e[1])
        if tag in ("fixed", "advice", "instance"):
            kind = {"fixed": FIXED, "advice": ADVICE, "instance": INSTANCE}[tag]
            return self.add_calculation(STORE, ValueSource(kind, e[1], self.add_rotation(e[2])))
        if tag == "challenge":
            return self.add_calculation(STORE, ValueSource(CHALLENGE, e[1]))
        if tag == "neg":
            if e[1][0] == "const":
                return self.add_constant(-e[1][1])
            r = self.add_expression(e[1])
            return r if self._is(r, 0) else self.add_calculation(NEGATE, r)
        if tag == "sum":
            a, b = e[1], e[2]
            if b[0] == "neg" or a[0] == "neg":        # a + (-b) -> a - b
                pos, neg = (a, b[1]) if b[0] == "neg" else (b, a[1])
                rp, rn = self.add_expression(pos), self.add_expression(neg)
                if self._is(rp, 0):
                    return self.add_calculation(NEGATE, rn)
                if self._is(rn, 0):
                    return rp
                return self.add_calculation(SUB, rp, rn)
            ra, rb = self.add_expression(a), self.add_expression(b)
            if self._is(ra, 0):
                return rb
            if self._is(rb, 0):
                return ra
            return self.add_calculation(ADD, *sorted((ra, rb)))
        if tag == "prod":
            ra, rb = self.add_expression(e[1]), self.add_expression(e[2])
            if self._is(ra, 0) or self._is(rb, 0):
                return ValueSource(CONSTANT, 0)
            if self._is(ra, 1):
                return rb
            if self._is(rb, 1):
                return ra
            if self._is(ra, 2):
                return self.add_calculation(DOUBLE, rb)
            if self._is(rb, 2):
                return self.add_calculation(DOUBLE, ra)
            if ra == rb:
                return self.add_calculation(SQUARE, ra)
            return self.add_calculation(MUL, *sorted((ra, rb)))
        if tag == "scaled":
            f = e[2] % FR
            if f == 0:
                return ValueSource(CONSTANT, 0)
            if f == 1:
                return self.add_expression(e[1])
            cst = self.add_constant(f)
            return self.add_calculation(MUL, self.add_expression(e[1]), cst)
        raise ValueError(f"unknown expression {tag!r}")

    # ---- C view (the layout zkb_graph_evaluate, the emulator and the oracle all read) -------------------------------------
    def calc_array(self) -> np.ndarray:
        """num_calculations x 11 u32: op, target, then (kind, index, rotation) of a, b, c — zkb_calculation's layout."""
        out = np.zeros((len(self.calculations), 11), dtype=np.uint32)
        for i, (op, target, a, b, c) in enumerate(self.calculations):
            out[i] = [op, target, *a, *b, *c]
        return out

    def c_graph(self):
        """(zkb_graph, keep-alive objects)"""
        calcs = self.calc_array()
        consts = _mont_limbs(self.constants)
        rots = np.array(self.rotations, dtype=np.int32)
        g = CGraph(calcs.ctypes.data_as(ctypes.POINTER(CCalculation)) if len(calcs) else None, len(calcs), self.num_intermediates,
                   _ptr(consts), len(self.constants), rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if len(rots) else None,
                   len(rots))
        return g, (calcs, consts, rots)

    @staticmethod
    def c_inputs(fixed, advice, instance, challenges, beta, gamma, theta, y, rot_scale):
        """Columns are integer handles (device) or host addresses (emulator); scalars are (4,) Montgomery limb arrays or None."""
        cols = [np.array(list(c), dtype=np.uint64) for c in (fixed, advice, instance)]
        ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4) if challenges is not None and len(challenges) else None
        sc = [np.ascontiguousarray(s, dtype=np.uint64).reshape(4) if s is not None else None for s in (beta, gamma, theta, y)]
        inp = CGraphInputs(_ptr(cols[0]), len(cols[0]), _ptr(cols[1]), len(cols[1]), _ptr(cols[2]), len(cols[2]),
                           _ptr(ch), 0 if ch is None else ch.shape[0], *[_ptr(s) for s in sc], rot_scale)
        return inp, (cols, ch, sc)

    def evaluate(self, values, fixed=(), advice=(), instance=(), challenges=None, beta=None, gamma=None, theta=None, y=None,
                 rot_scale: int = 1) -> None:
        """values[idx] = graph.evaluate(..., previous_value = values[idx], idx, rot_scale, isize) for every row, on the device.
        `values` and the columns are `halo2.Polynomial`s of the extended domain's size."""
        from . import halo2

        g, keep_g = self.c_graph()
        inp, keep_i = self.c_inputs([p._h.value for p in fixed], [p._h.value for p in advice], [p._h.value for p in instance],
                                    challenges, beta, gamma, theta, y, rot_scale)
        halo2.check(halo2.lib().zkb_graph_evaluate(ctypes.byref(g), ctypes.byref(inp), values._h))
        del keep_g, keep_i

    def evaluate_dev(self, d_values: int, rows: int, fixed=(), advice=(), instance=(), challenges=None, beta=None, gamma=None,
                     theta=None, y=None, rot_scale: int = 1, window: bool = False, halo_lo: int = 0, halo_hi: int = 0,
                     stream: int = 0) -> None:
        """`evaluate` on device pointers and the caller's stream (zkb_graph_evaluate_dev).  Columns are device addresses.
        window=True: a row window of the extended domain — every column holds halo_lo + rows + halo_hi elements and row i reads
        element halo_lo + i + rotation * rot_scale (the per-rank call of distributed.ShardedQuotient)."""
        from . import halo2

        g, keep_g = self.c_graph()
        inp, keep_i = self.c_inputs(list(fixed), list(advice), list(instance), challenges, beta, gamma, theta, y, rot_scale)
        halo2.check(halo2.lib().zkb_graph_evaluate_dev(ctypes.byref(g), ctypes.byref(inp), ctypes.c_void_p(d_values), rows,
                                                       1 if window else 0, halo_lo, halo_hi, ctypes.c_void_p(stream)))
        del keep_g, keep_i

    def rotation_span(self, rot_scale: int) -> tuple[int, int]:
        """(halo_lo, halo_hi): how many rows before / after a row its column queries reach."""
        used = {s.rotation for c in self.calculations for s in c[2:5] if s.kind in (FIXED, ADVICE, INSTANCE)}
        offs = [self.rotations[r] * rot_scale for r in used] or [0]
        return max(0, -min(offs)), max(0, max(offs))

    @staticmethod
    def last_info() -> dict:
        from . import halo2

        v = [ctypes.c_uint32(0) for _ in range(4)]
        halo2.check(halo2.lib().zkb_graph_last_info(*[ctypes.byref(x) for x in v]))
        return dict(zip(("instructions", "slots", "polys_read", "bytes_per_row"), [x.value for x in v]))


# ---- the hand-written terms of evaluate_h as graphs -------------------------------------------------------------------------------
# Upstream compiles only the custom gates and the lookups' compressed expressions into GraphEvaluators and writes the permutation
# and lookup product terms out by hand inside evaluate_h's row loop.  On the device they are graphs over the same value sources
# (every coset a fixed / advice / instance column), folded into the running value exactly as upstream folds them:
# value = value * y + term, in upstream's order [UPSTREAM-UNVERIFIED: halo2-axiom is not vendored; the terms are those of the
# halo2 book's permutation and lookup arguments, the order is the one of halo2_proofs 0.3's evaluate_h].
DELTA = pow(7, 1 << 28, FR)   # halo2curves bn256::Fr::DELTA = MULTIPLICATIVE_GENERATOR^(2^S)


def _col(ref, rot_idx: int) -> ValueSource:
    kind = {"fixed": FIXED, "advice": ADVICE, "instance": INSTANCE}[ref[0]]
    return ValueSource(kind, ref[1], rot_idx)


def permutation_graph(columns: list, chunk_len: int, last_rotation: int, l0, l_last, l_active, x_coset, sigmas: list, zs: list) -> GraphEvaluator:
    """All h(X) terms of the permutation argument, folded from the previous value with y:
        l_0 (1 - z_0)
        l_last (z_last^2 - z_last)
        l_0 (z_i - z_{i-1}(w^last_rotation X))                                          for every set i > 0
        l_active ( z_i(wX) prod_j (col_j + beta sigma_j + gamma) - z_i(X) prod_j (col_j + delta^j beta X + gamma) )   for every set
    columns: the permutation's columns in order, as ("advice" | "fixed" | "instance", index); they are cut into sets of
    chunk_len = cs.degree() - 2.  l0, l_last, l_active, x_coset (the coset evaluation of the polynomial X, i.e. zeta w^idx), one
    sigma coset per column and one product coset z_i per set are given the same way; last_rotation = -(blinding_factors + 1).
    The power of delta keeps running across the sets, as upstream's `current_delta` does."""
    g = GraphEvaluator()
    r0, r1, rl = g.add_rotation(0), g.add_rotation(1), g.add_rotation(last_rotation)
    sets = [columns[i:i + chunk_len] for i in range(0, len(columns), chunk_len)]
    assert len(zs) == len(sets) and len(sigmas) == len(columns)
    beta, gamma, one = ValueSource(BETA), ValueSource(GAMMA), ValueSource(CONSTANT, 1)
    c = g.add_calculation
    terms = [c(MUL, c(SUB, one, _col(zs[0], r0)), _col(l0, r0))]
    zl = _col(zs[-1], r0)
    terms.append(c(MUL, c(SUB, c(SQUARE, zl), zl), _col(l_last, r0)))
    for i in range(1, len(sets)):
        terms.append(c(MUL, c(SUB, _col(zs[i], r0), _col(zs[i - 1], rl)), _col(l0, r0)))
    bx = c(MUL, beta, _col(x_coset, r0))
    j = 0
    for i, cols in enumerate(sets):
        left, right = _col(zs[i], r1), _col(zs[i], r0)
        for col in cols:
            v = _col(col, r0)
            left = c(MUL, left, c(ADD, c(ADD, v, c(MUL, beta, _col(sigmas[j], r0))), gamma))
            right = c(MUL, right, c(ADD, c(ADD, v, c(MUL, bx, g.add_constant(pow(DELTA, j, FR)))), gamma))
            j += 1
        terms.append(c(MUL, c(SUB, left, right), _col(l_active, r0)))
    g.add_horner(ValueSource(PREVIOUS), terms, ValueSource(Y))
    return g


def lookup_graph(input_exprs: list, table_exprs: list, l0, l_last, l_active, z, permuted_input, permuted_table) -> GraphEvaluator:
    """All h(X) terms of one lookup argument, folded from the previous value with y:
        l_0 (1 - z)        l_last (z^2 - z)
        l_active ( z(wX) (a' + beta) (s' + gamma) - z(X) (A + beta) (S + gamma) )
        l_0 (a' - s')      l_active (a' - s') (a' - a'(w^-1 X))
    A, S: the input / table expressions compressed with theta (Horner from zero), as upstream's per-lookup GraphEvaluator does."""
    g = GraphEvaluator()
    theta, beta, gamma, one = ValueSource(THETA), ValueSource(BETA), ValueSource(GAMMA), ValueSource(CONSTANT, 1)
    c = g.add_calculation

    def compress(exprs):
        return g.add_horner(ValueSource(CONSTANT, 0), [g.add_expression(e) for e in exprs], theta)

    A, S = compress(input_exprs), compress(table_exprs)
    r0, r1, rm1 = g.add_rotation(0), g.add_rotation(1), g.add_rotation(-1)
    zz, zw = _col(z, r0), _col(z, r1)
    ap, apm, sp = _col(permuted_input, r0), _col(permuted_input, rm1), _col(permuted_table, r0)
    d = c(SUB, ap, sp)
    terms = [c(MUL, c(SUB, one, zz), _col(l0, r0)),
             c(MUL, c(SUB, c(SQUARE, zz), zz), _col(l_last, r0)),
             c(MUL, c(SUB, c(MUL, c(MUL, zw, c(ADD, ap, beta)), c(ADD, sp, gamma)), c(MUL, c(MUL, zz, c(ADD, A, beta)), c(ADD, S, gamma))), _col(l_active, r0)),
             c(MUL, d, _col(l0, r0)),
             c(MUL, c(MUL, d, c(SUB, ap, apm)), _col(l_active, r0))]
    g.add_horner(ValueSource(PREVIOUS), terms, ValueSource(Y))
    return g


def custom_gates_graph(gate_polys: list) -> GraphEvaluator:
    """evaluate_h's `custom_gates` evaluator: every gate polynomial compiled with add_expression, then
    Horner(PreviousValue, parts, Y)."""
    g = GraphEvaluator()
    g.add_horner(ValueSource(PREVIOUS), [g.add_expression(e) for e in gate_polys], ValueSource(Y))
    return g


class QuotientEvaluator:
    """`plonk::evaluation::Evaluator` with its `evaluate_h`: the custom gates, then the permutation argument, then every lookup,
    each folded into the running value with y, over cosets resident in HBM.  Construction compiles the graphs that do not depend
    on where the auxiliary cosets sit; `evaluate_h` appends l_0 / l_last / l_active / the coset of X / the sigma cosets to the
    fixed columns and the product / permuted cosets to the advice columns and runs one zkb_graph_evaluate per argument.

    permutation: None or dict(columns=[("advice" | "fixed" | "instance", index), ...], chunk_len=cs.degree() - 2,
    last_rotation=-(blinding_factors + 1));  lookups: [(input_expressions, table_expressions), ...]."""

    def __init__(self, gate_polys: list, permutation: dict | None = None, lookups: list | tuple = ()):
        self.custom_gates = custom_gates_graph(gate_polys) if gate_polys else None
        self.permutation = permutation
        self.lookups = list(lookups)

    def evaluate_h(self, values, fixed, advice, instance, challenges, y, beta, gamma, theta, rot_scale, l0, l_last, l_active,
                   x_coset=None, sigma_cosets=(), permutation_product_cosets=(), lookup_cosets=(), evaluate=None):
        """values: the running h evaluations (start from zeros), overwritten.  lookup_cosets: per lookup (product, permuted
        input, permuted table).  `evaluate(graph, values, fixed, advice, instance)` is the per-graph call; the default is the
        device (GraphEvaluator.evaluate on halo2.Polynomial handles)."""
        if evaluate is None:
            def evaluate(graph, v, f, a, i):
                graph.evaluate(v, f, a, i, challenges=challenges, beta=beta, gamma=gamma, theta=theta, y=y, rot_scale=rot_scale)
        fixed, advice, instance = list(fixed), list(advice), list(instance)
        if self.custom_gates is not None:
            evaluate(self.custom_gates, values, fixed, advice, instance)
        nf, na = len(fixed), len(advice)
        aux = [l0, l_last, l_active]           # fixed columns nf, nf + 1, nf + 2 for every argument below
        if self.permutation is not None:
            p = self.permutation
            ncols = len(p["columns"])
            nsets = (ncols + p["chunk_len"] - 1) // p["chunk_len"]
            assert x_coset is not None and len(sigma_cosets) == ncols and len(permutation_product_cosets) == nsets
            g = permutation_graph(p["columns"], p["chunk_len"], p["last_rotation"], ("fixed", nf), ("fixed", nf + 1), ("fixed", nf + 2),
                                  ("fixed", nf + 3), [("fixed", nf + 4 + j) for j in range(ncols)], [("advice", na + s) for s in range(nsets)])
            evaluate(g, values, fixed + aux + [x_coset] + list(sigma_cosets), advice + list(permutation_product_cosets), instance)
        assert len(lookup_cosets) == len(self.lookups)
        for (inputs, table), (z, a_perm, s_perm) in zip(self.lookups, lookup_cosets):
            g = lookup_graph(inputs, table, ("fixed", nf), ("fixed", nf + 1), ("fixed", nf + 2), ("advice", na), ("advice", na + 1), ("advice", na + 2))
            evaluate(g, values, fixed + aux, advice + [z, a_perm, s_perm], instance)
        return values
