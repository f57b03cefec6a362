"""plonk.py — the hot-path calls of halo2's keygen_pk / create_proof, composed on polynomials resident in HBM.

The reference reaches the MSM / NTT path only through `keygen_pk` and `create_proof::<_, ProverGWC<_>, ...>`
(/root/reference/aggregator/src/wrapper.rs:106-109, 129-137) and its only self-check is the verifier's accept bit
(wrapper.rs:141-155).  This module is the host-side mirror of that composition for a halo2-base-shaped circuit (one custom
gate per advice column + a permutation argument), written against the same C ABI a patched halo2 would call:

    ProvingKey.keygen / .write / .read   keygen_pk's device work (fixed / sigma polynomials and their extended cosets, l_0, l_last,
                                         l_active) kept RESIDENT under handles and reused by every proof of the IVC loop
                                         (wrapper.rs:884-900); the file form is raw Montgomery limbs (SerdeFormat::RawBytesUnchecked,
                                         wrapper.rs:970-988), read straight into HBM by zkb_poly_load_file
    permutation_commit                   plonk::permutation::prover::Argument::commit — the grand product z of every column set:
                                         numerators / denominators by the row interpreter, batch inversion, product scan,
                                         blinding rows, commitment
    create_proof                         advice commitments -> z -> h(X) on the extended cosets (QuotientEvaluator.evaluate_h,
                                         divide_by_vanishing_poly, extended_to_coeff, h pieces) -> evaluations -> GWC openings
                                         (scale_add fold, kate_division, commit)

The transcript is a hash stub (Blake2b over the bytes written so far): challenges depend on everything committed before them,
which is all the acceptance test needs; the real Poseidon / Blake2b transcripts stay on the Rust side.  Nothing here is a
CPU fallback: every polynomial operation is a libzkb200 call; Python integers only prepare O(n) keygen inputs and challenges.
"""
from __future__ import annotations

import hashlib
import struct
from typing import NamedTuple

import numpy as np

from . import evaluation as ev
from .halo2 import EvaluationDomain, ParamsKZG, Polynomial

FR = ev.FR
DELTA = ev.DELTA
_R = (1 << 256) % FR
_RINV = pow(_R, -1, FR)


def mont(x: int) -> np.ndarray:
    v = x % FR * _R % FR
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)


def mont_vec(xs) -> np.ndarray:
    return np.stack([mont(x) for x in xs]) if len(xs) else np.zeros((0, 4), dtype=np.uint64)


def unmont(limbs) -> int:
    return sum(int(limbs[j]) << (64 * j) for j in range(4)) * _RINV % FR


class Circuit(NamedTuple):
    """halo2-base-shaped constraint system: `gates` are expression tuples over ("advice" | "fixed", column, rotation) as
    evaluation.GraphEvaluator.add_expression takes them; degree 4 (extended_k = k + 2, permutation chunks of 2 columns)."""
    k: int
    num_advice: int
    num_fixed: int
    gates: list
    permutation_columns: list          # [("advice", i), ...]
    blinding_factors: int = 5
    lookups: tuple = ()                # ((input expressions, table expressions), ...)

    @property
    def n(self) -> int:
        return 1 << self.k

    @property
    def usable_rows(self) -> int:
        return self.n - (self.blinding_factors + 1)

    @property
    def chunk_len(self) -> int:
        return 2                        # cs.degree() - 2

    def queries(self):
        """(kind, column, rotation) of every column query of the gates and the permutation, first-seen order (the prover's
        evaluation order)"""
        seen = []

        def walk(e):
            if e[0] in ("advice", "fixed"):
                if e not in seen:
                    seen.append(e)
            else:
                for sub in e[1:]:
                    if isinstance(sub, tuple):
                        walk(sub)

        for g in self.gates:
            walk(g)
        for inputs, table in self.lookups:
            for e in list(inputs) + list(table):
                walk(e)
        for kind, col in self.permutation_columns:   # enable_equality queries the column at the current rotation
            if (kind, col, 0) not in seen:
                seen.append((kind, col, 0))
        return seen


class Transcript:
    """hash-chain stand-in for halo2's transcript: write_point / write_scalar absorb bytes, squeeze_challenge hashes the state"""

    def __init__(self, label: bytes = b"zkb200-plonk"):
        self.h = hashlib.blake2b(label)

    def write_point(self, jac: np.ndarray) -> None:
        self.h.update(np.ascontiguousarray(jac, dtype=np.uint64).tobytes())

    def write_scalar(self, x: int) -> None:
        self.h.update(int(x).to_bytes(32, "little"))

    def squeeze_challenge(self) -> int:
        d = self.h.digest()
        self.h.update(b"\x01" + d)
        return int.from_bytes(d, "little") % FR


# ---- proving key ---------------------------------------------------------------------------------------------------------------
_MAGIC = b"ZKBPK1\0\0"


class ProvingKey:
    """The per-circuit polynomials of halo2's ProvingKey, resident in HBM: fixed_polys / fixed_cosets, the permutation's sigma
    polys / cosets, l_0 / l_last / l_active cosets — plus two derived cosets the quotient evaluation reads (X on the extended coset,
    and the identity columns delta^j X of the permutation numerators in the Lagrange basis).  Built once (keygen) or loaded
    once (read) and shared by every create_proof: a second proof uploads witness data only."""

    def __init__(self, circuit: Circuit):
        self.circuit = circuit
        self.domain = EvaluationDomain(4, circuit.k)
        self.fixed_polys, self.fixed_cosets, self.fixed_lagrange = [], [], []
        self.sigma_polys, self.sigma_cosets, self.sigma_lagrange = [], [], []
        self.id_lagrange = []
        self.l0 = self.l_last = self.l_active = self.x_coset = None
        self.fixed_commitments, self.sigma_commitments = [], []

    # what keygen_pk computes on the device
    @classmethod
    def keygen(cls, params: ParamsKZG, circuit: Circuit, fixed_lagrange: list, sigma_mapping: list) -> "ProvingKey":
        """fixed_lagrange: one (n, 4) Montgomery array per fixed column; sigma_mapping[j][i] = (column j', row i') that cell (j, i)
        of permutation column j is mapped to by the copy constraints (identity = (j, i))."""
        pk = cls(circuit)
        d, n, u = pk.domain, circuit.n, circuit.usable_rows
        omega = unmont(d.get_omega())
        wp = [1] * n
        for i in range(1, n):
            wp[i] = wp[i - 1] * omega % FR
        for col in fixed_lagrange:
            lag = Polynomial(col)
            pk.fixed_commitments.append(lag.commit(params, lagrange=True))
            pk.fixed_lagrange.append(lag)
            p = lag.slice(0, n).lagrange_to_coeff(d)
            pk.fixed_polys.append(p)
            pk.fixed_cosets.append(p.coeff_to_extended(d))
        for j, mapping in enumerate(sigma_mapping):
            lag = Polynomial(mont_vec([pow(DELTA, cj, FR) * wp[ri] % FR for (cj, ri) in mapping]))
            pk.sigma_commitments.append(lag.commit(params, lagrange=True))
            pk.sigma_lagrange.append(lag)
            p = lag.slice(0, n).lagrange_to_coeff(d)
            pk.sigma_polys.append(p)
            pk.sigma_cosets.append(p.coeff_to_extended(d))
        pk._derive(u)
        return pk

    def _derive(self, u: int) -> None:
        """l_0, l_last, l_active, the coset of X and the identity columns: functions of (k, blinding_factors) only"""
        d, n = self.domain, self.circuit.n
        unit = lambda i: [1 if r == i else 0 for r in range(n)]  # noqa: E731

        def coset_of_lagrange(vals):
            p = Polynomial(mont_vec(vals)).lagrange_to_coeff(d)
            c = p.coeff_to_extended(d)
            p.free()
            return c

        self.l0 = coset_of_lagrange(unit(0))
        self.l_last = coset_of_lagrange(unit(u))
        self.l_active = coset_of_lagrange([1 if r < u else 0 for r in range(n)])
        x = Polynomial(mont_vec([0, 1] + [0] * (n - 2)))
        self.x_coset = x.coeff_to_extended(d)
        for j in range(len(self.circuit.permutation_columns)):     # delta^j X in the Lagrange basis = delta^j omega^i
            idp = Polynomial(mont_vec([0, pow(DELTA, j, FR)] + [0] * (n - 2)))
            self.id_lagrange.append(idp.coeff_to_lagrange(d))
        x.free()

    # ---- file form: header, then every polynomial as raw limbs (SerdeFormat::RawBytesUnchecked element encoding).  The container
    # layout of halo2's ProvingKey::write is [UPSTREAM-UNVERIFIED] (halo2-axiom is not vendored); this one is self-describing and
    # the reader takes offsets, so adopting upstream's order is a change of the index table only.
    def write(self, path: str) -> None:
        c = self.circuit
        groups = [("fixed_lagrange", self.fixed_lagrange), ("fixed_polys", self.fixed_polys), ("fixed_cosets", self.fixed_cosets), ("sigma_lagrange", self.sigma_lagrange),
                  ("sigma_polys", self.sigma_polys), ("sigma_cosets", self.sigma_cosets)]
        with open(path, "wb") as f:
            f.write(_MAGIC + struct.pack("<IIIII", c.k, c.num_advice, c.num_fixed, len(c.permutation_columns), c.blinding_factors))
            for _, polys in groups:
                for p in polys:
                    f.write(p.to_host().tobytes())
            for cm in self.fixed_commitments + self.sigma_commitments:
                f.write(np.ascontiguousarray(cm, dtype=np.uint64).tobytes())

    @classmethod
    def read(cls, path: str, circuit: Circuit) -> "ProvingKey":
        """Every polynomial goes from the file straight into HBM (zkb_poly_load_file: pinned double buffering, no host copy of
        the arrays); only the 96-byte commitments are read on the host."""
        pk = cls(circuit)
        n, N = circuit.n, pk.domain.extended_len()
        with open(path, "rb") as f:
            head = f.read(len(_MAGIC) + 20)
        if head[:len(_MAGIC)] != _MAGIC:
            raise ValueError("not a zkb200 proving-key file")
        k, na, nf, np_, bf = struct.unpack("<IIIII", head[len(_MAGIC):])
        if (k, na, nf, np_, bf) != (circuit.k, circuit.num_advice, circuit.num_fixed, len(circuit.permutation_columns), circuit.blinding_factors):
            raise ValueError("proving key was generated for a different circuit")
        off = len(head)

        def take(count, length):
            nonlocal off
            out = []
            for _ in range(count):
                out.append(Polynomial.load_file(path, off, length))
                off += length * 32
            return out

        pk.fixed_lagrange, pk.fixed_polys, pk.fixed_cosets = take(nf, n), take(nf, n), take(nf, N)
        pk.sigma_lagrange, pk.sigma_polys, pk.sigma_cosets = take(np_, n), take(np_, n), take(np_, N)
        with open(path, "rb") as f:
            f.seek(off)
            cms = np.frombuffer(f.read((nf + np_) * 96), dtype=np.uint64).reshape(-1, 12)
        pk.fixed_commitments, pk.sigma_commitments = [c.copy() for c in cms[:nf]], [c.copy() for c in cms[nf:]]
        pk._derive(circuit.usable_rows)
        return pk

    def free(self) -> None:
        for p in (self.fixed_lagrange + self.fixed_polys + self.fixed_cosets + self.sigma_polys + self.sigma_cosets + self.sigma_lagrange + self.id_lagrange
                  + [self.l0, self.l_last, self.l_active, self.x_coset]):
            if p is not None:
                p.free()


# ---- plonk::permutation::prover::Argument::commit ---------------------------------------------------------------------------------
def permutation_commit(params: ParamsKZG, pk: ProvingKey, advice_lagrange: list, beta: int, gamma: int, blind_rows) -> list:
    """One grand product z per set of chunk_len columns (Lagrange basis, resident):
        z[0] = last value of the previous set (1 for the first),  z[i+1] = z[i] * prod_j (v_j[i] + beta delta^j w^i + gamma) / (v_j[i] + beta sigma_j[i] + gamma)
    for the usable rows, random blinding values in the last blinding_factors rows.  Numerator and denominator products are one
    row-interpreter pass each (rot_scale 1 on the 2^k domain), the division is a batch inversion + element-wise product, the
    running product a device scan.  blind_rows(set_index) -> (blinding_factors, 4) Montgomery array.  Returns [(z_lagrange, commitment)]."""
    c = pk.circuit
    n, u, bf = c.n, c.usable_rows, c.blinding_factors
    cols = c.permutation_columns
    out = []
    start = 1
    for s0 in range(0, len(cols), c.chunk_len):
        chunk = list(range(s0, min(s0 + c.chunk_len, len(cols))))
        gn, gd = ev.GraphEvaluator(), ev.GraphEvaluator()
        num = den = None
        for t, j in enumerate(chunk):
            v = ("advice", t, 0)
            tn = ("sum", ("sum", v, ("prod", ("challenge", 0), ("fixed", t, 0))), ("challenge", 1))
            num = tn if num is None else ("prod", num, tn)
            den = tn if den is None else ("prod", den, tn)   # same shape: the fixed columns differ (identity vs sigma)
        gn.add_expression(num)
        gd.add_expression(den)
        adv = [advice_lagrange[cols[j][1]] for j in chunk]
        ch = np.stack([mont(beta), mont(gamma)])
        zn, zd = Polynomial.zeros(n), Polynomial.zeros(n)
        gn.evaluate(zn, fixed=[pk.id_lagrange[j] for j in chunk], advice=adv, challenges=ch, rot_scale=1)
        gd.evaluate(zd, fixed=[pk.sigma_lagrange[j] for j in chunk], advice=adv, challenges=ch, rot_scale=1)
        zd.batch_invert()
        zn.mul(zd).prefix_product()                 # z[0] = 1, z[i] = prod_{t<i} ratio[t]
        if start != 1:
            zn.scale_add(mont(start))
        last = unmont(zn.slice(u, 1).to_host()[0])   # z[u]: 1 for the last set of a satisfied permutation
        zn.write(u + 1, blind_rows(len(out)))
        zd.free()
        out.append((zn, zn.commit(params, lagrange=True)))
        start = last
        assert bf == n - u - 1
    return out


# ---- plonk::lookup::prover: commit_permuted and commit_product ------------------------------------------------------------------------
def _permute_expression_pair(input_: Polynomial, table: Polynomial, usable_rows: int):
    import ctypes

    from .halo2 import check, lib
    ha, ht = ctypes.c_uint64(0), ctypes.c_uint64(0)
    check(lib().zkb_lookup_permute_expression_pair(input_._h, table._h, usable_rows, ctypes.byref(ha), ctypes.byref(ht)))
    return Polynomial(_handle=ha.value), Polynomial(_handle=ht.value)


def lookup_commit_permuted(params: ParamsKZG, pk: ProvingKey, lookup, advice_lagrange: list, theta: int, blind_rows):
    """Argument::commit_permuted: compress the input / table expressions with theta over the 2^k domain (one row-interpreter pass
    each, rot_scale 1), permute_expression_pair on the device, blinding rows, commitments.
    Returns dict(A, S: compressed columns; Ap, Sp: permuted columns; commitments)."""
    c = pk.circuit
    n, u = c.n, c.usable_rows
    inputs, table = lookup
    cols = {}
    for name, exprs in (("A", inputs), ("S", table)):
        g = ev.GraphEvaluator()
        g.add_horner(ev.ValueSource(ev.CONSTANT, 0), [g.add_expression(e) for e in exprs], ev.ValueSource(ev.THETA))
        out = Polynomial.zeros(n)
        g.evaluate(out, fixed=pk.fixed_lagrange, advice=advice_lagrange, theta=mont(theta), rot_scale=1)
        cols[name] = out
    ap, sp = _permute_expression_pair(cols["A"], cols["S"], u)
    ap.write(u, blind_rows())
    sp.write(u, blind_rows())
    cols.update(Ap=ap, Sp=sp, commitments=[ap.commit(params, lagrange=True), sp.commit(params, lagrange=True)])
    return cols


def lookup_commit_product(params: ParamsKZG, pk: ProvingKey, cols: dict, beta: int, gamma: int, blind_rows):
    """Argument::commit_product: z[0] = 1, z[i+1] = z[i] (A_i + beta)(S_i + gamma) / ((A'_i + beta)(S'_i + gamma)) on the usable rows."""
    c = pk.circuit
    n, u = c.n, c.usable_rows
    g = ev.GraphEvaluator()
    g.add_expression(("prod", ("sum", ("advice", 0, 0), ("challenge", 0)), ("sum", ("advice", 1, 0), ("challenge", 1))))
    ch = np.stack([mont(beta), mont(gamma)])
    num, den = Polynomial.zeros(n), Polynomial.zeros(n)
    g.evaluate(num, advice=[cols["A"], cols["S"]], challenges=ch, rot_scale=1)
    g.evaluate(den, advice=[cols["Ap"], cols["Sp"]], challenges=ch, rot_scale=1)
    den.batch_invert()
    num.mul(den).prefix_product()
    den.free()
    num.write(u + 1, blind_rows())
    return num, num.commit(params, lagrange=True)


# ---- create_proof -----------------------------------------------------------------------------------------------------------------
def create_proof(params: ParamsKZG, pk: ProvingKey, advice_lagrange_host: list, rng: np.random.Generator, hooks: dict | None = None) -> dict:
    """advice_lagrange_host: one (n, 4) Montgomery array per advice column with the witness in the usable rows (the blinding rows
    are filled here).  hooks: test instrumentation — {"after_advice" | "after_z" | "after_h": fn(list of resident polynomials)}
    may perturb device data to show that the acceptance check notices."""
    from .halo2 import _fr  # noqa: F401
    hooks = hooks or {}
    c, d = pk.circuit, pk.domain
    n, u, N = c.n, c.usable_rows, d.extended_len()
    omega = unmont(d.get_omega())
    tr = Transcript()
    for cm in pk.fixed_commitments + pk.sigma_commitments:
        tr.write_point(cm)

    def rand_fr(count):
        v = rng.integers(0, 1 << 62, size=(count, 4), dtype=np.uint64)   # < 2^254 > r is possible: reduce through Python ints
        return mont_vec([sum(int(r[j]) << (62 * j) for j in range(4)) % FR for r in v])

    # 1. advice columns: blind, upload once, commit in the Lagrange basis
    advice = []
    for col in advice_lagrange_host:
        a = np.array(col, dtype=np.uint64, copy=True).reshape(n, 4)
        a[u:] = rand_fr(n - u)
        advice.append(Polynomial(a))
    if "after_advice" in hooks:
        hooks["after_advice"](advice)
    advice_commitments = [p.commit(params, lagrange=True) for p in advice]
    for cm in advice_commitments:
        tr.write_point(cm)
    # 2. lookups: permuted columns (theta), then (beta, gamma) the permutation and lookup products
    theta = tr.squeeze_challenge()
    lookups = [lookup_commit_permuted(params, pk, lk, advice, theta, lambda: rand_fr(n - u)) for lk in c.lookups]
    for lk in lookups:
        for cm in lk["commitments"]:
            tr.write_point(cm)
    beta, gamma = tr.squeeze_challenge(), tr.squeeze_challenge()
    zs = permutation_commit(params, pk, advice, beta, gamma, lambda s: rand_fr(c.blinding_factors))
    if "after_z" in hooks:
        hooks["after_z"]([z for z, _ in zs])
    z_commitments = [z.commit(params, lagrange=True) for z, _ in zs] if "after_z" in hooks else [cm for _, cm in zs]
    for cm in z_commitments:
        tr.write_point(cm)
    for lk in lookups:
        lk["z"], lk["z_commitment"] = lookup_commit_product(params, pk, lk, beta, gamma, lambda: rand_fr(c.blinding_factors))
        tr.write_point(lk["z_commitment"])
    y = tr.squeeze_challenge()
    # 3. h(X): coefficient forms, extended cosets, evaluate_h, / (X^n - 1), back to coefficients, pieces
    advice_polys = [p.slice(0, n).lagrange_to_coeff(d) for p in advice]
    z_polys = [z.slice(0, n).lagrange_to_coeff(d) for z, _ in zs]
    advice_cosets = [p.coeff_to_extended(d) for p in advice_polys]
    z_cosets = [p.coeff_to_extended(d) for p in z_polys]
    for lk in lookups:   # coefficient forms and extended cosets of z, A', S'
        lk["polys"] = [lk[name].slice(0, n).lagrange_to_coeff(d) for name in ("z", "Ap", "Sp")]
        lk["cosets"] = [p.coeff_to_extended(d) for p in lk["polys"]]
    q = ev.QuotientEvaluator(c.gates, dict(columns=c.permutation_columns, chunk_len=c.chunk_len, last_rotation=-(c.blinding_factors + 1)),
                             list(c.lookups))
    h = Polynomial.zeros(N)
    q.evaluate_h(h, pk.fixed_cosets, advice_cosets, [], None, mont(y), mont(beta), mont(gamma), mont(theta), 1 << (d.extended_k - c.k),
                 pk.l0, pk.l_last, pk.l_active, pk.x_coset, pk.sigma_cosets, z_cosets, [tuple(lk["cosets"]) for lk in lookups])
    if "after_h" in hooks:
        hooks["after_h"]([h])
    d.divide_by_vanishing_poly(h)
    h.extended_to_coeff(d)
    pieces = [h.slice(i * n, n) for i in range(d.quotient_poly_degree)]
    h_commitments = [p.commit(params) for p in pieces]
    for cm in h_commitments:
        tr.write_point(cm)
    x = tr.squeeze_challenge()
    # 4. evaluations
    queries = []   # (label, polynomial, commitment, point, evaluation)

    def query(label, poly, cm, rot):
        pt = x * pow(omega, rot % n, FR) % FR
        val = unmont(poly.eval(mont(pt)))
        queries.append((label, poly, cm, pt, val))
        tr.write_scalar(val)

    for kind, col, rot in c.queries():
        if kind == "advice":
            query(("advice", col, rot), advice_polys[col], advice_commitments[col], rot)
        else:
            query(("fixed", col, rot), pk.fixed_polys[col], pk.fixed_commitments[col], rot)
    for j, p in enumerate(pk.sigma_polys):
        query(("sigma", j, 0), p, pk.sigma_commitments[j], 0)
    for s, p in enumerate(z_polys):
        query(("z", s, 0), p, z_commitments[s], 0)
        query(("z", s, 1), p, z_commitments[s], 1)
        if s + 1 < len(z_polys):
            query(("z", s, -(c.blinding_factors + 1)), p, z_commitments[s], -(c.blinding_factors + 1))
    for li, lk in enumerate(lookups):
        zp, app, spp = lk["polys"]
        query(("lookup_z", li, 0), zp, lk["z_commitment"], 0)
        query(("lookup_z", li, 1), zp, lk["z_commitment"], 1)
        query(("lookup_a", li, 0), app, lk["commitments"][0], 0)
        query(("lookup_a", li, -1), app, lk["commitments"][0], -1)
        query(("lookup_s", li, 0), spp, lk["commitments"][1], 0)
    for i, p in enumerate(pieces):
        query(("h", i, 0), p, h_commitments[i], 0)
    # 5. GWC multiopen: per distinct point, fold polynomials and evaluations with powers of v, divide by (X - point), commit
    v = tr.squeeze_challenge()
    openings = []
    points = []
    for _, _, _, pt, _ in queries:
        if pt not in points:
            points.append(pt)
    for pt in points:
        group = [qq for qq in queries if qq[3] == pt]
        acc = group[0][1].slice(0, n)
        acc_eval = group[0][4]
        for _, poly, _, _, val in group[1:]:
            acc.scale_add(mont(v), poly)
            acc_eval = (acc_eval * v + val) % FR
        acc.add_const(mont(-acc_eval))
        wit = acc.kate_division(mont(pt))
        w_cm = wit.commit(params)
        tr.write_point(w_cm)
        openings.append({"point": pt, "labels": [g[0] for g in group], "witness": w_cm})
        acc.free()
        wit.free()
    proof = {
        "advice_commitments": advice_commitments, "z_commitments": z_commitments, "h_commitments": h_commitments,
        "evals": {qq[0]: qq[4] for qq in queries}, "eval_points": {qq[0]: qq[3] for qq in queries},
        "commitment_of": {qq[0]: qq[2] for qq in queries}, "openings": openings,
        "lookup_commitments": [lk["commitments"] + [lk["z_commitment"]] for lk in lookups],
        "challenges": {"theta": theta, "beta": beta, "gamma": gamma, "y": y, "x": x, "v": v},
    }
    for lk in lookups:
        for p in [lk["A"], lk["S"], lk["Ap"], lk["Sp"], lk["z"]] + lk["polys"] + lk["cosets"]:
            p.free()
    for p in advice + advice_polys + z_polys + advice_cosets + z_cosets + pieces + [h] + [z for z, _ in zs]:
        p.free()
    return proof
