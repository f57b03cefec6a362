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
        key = (op, a, b or z, z)
        for c in self.calculations:
            if (c[0], c[2], c[3], c[4]) == key and c[0] != MUL_ADD:
                return ValueSource(INTERMEDIATE, c[1])
        target = self.num_intermediates
        self.num_intermediates += 1
        self.calculations.append((op, target, a, b or z, z))
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
