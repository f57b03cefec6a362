"""End-to-end acceptance analogue of the reference's only self-check — `assert!(accept)` after verify_proof
(/root/reference/aggregator/src/wrapper.rs:141-155) — for the hot path composed the way create_proof composes it
(zksnap-circuits-halo2_b200/plonk.py): advice commitments, permutation grand product, h(X) on the extended cosets, h pieces,
evaluations, GWC openings — all on polynomials resident in HBM.

The verifier below is plain Python integers (oracle/pyref.py for the group): it recomputes the challenges, checks the vanishing
identity  h(x) (x^n - 1) = fold_y(gates(x), permutation terms(x))  from the opened evaluations, and checks every KZG opening
IN THE EXPONENT with the test-known trapdoor s:  C - [v]G == [s - x] W  (what the pairing check decides, without pairings).
One wrong device output anywhere (witness cell, product column, quotient coset, commitment) makes it reject.
"""
import ctypes
import importlib
import os
import tempfile

import numpy as np
import pytest

from oracle import pyref as R

zkb = importlib.import_module("zksnap-circuits-halo2_b200")
plonk = importlib.import_module("zksnap-circuits-halo2_b200.plonk")
ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")

pytestmark = pytest.mark.gpu
FR = R.FR
K = 12
S_TRAPDOOR = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % FR


def _circuit():
    a = [("advice", 0, r) for r in range(4)]
    gate = ("prod", ("fixed", 0, 0), ("sum", ("sum", a[0], ("prod", a[1], a[2])), ("neg", a[3])))   # halo2-base: q (a + b c - d)
    # lookup: advice column 2 (range-checked cells) must lie in the table held by fixed column 1
    return plonk.Circuit(k=K, num_advice=3, num_fixed=2, gates=[gate], permutation_columns=[("advice", 0), ("advice", 1)],
                         lookups=(([("advice", 2, 0)], [("fixed", 1, 0)]),))


def _witness(c, seed):
    """column 0: halo2-base gate chains on rows 0, 4, 8, ...; column 1: free cells, 200 of them copy-constrained to cells of
    column 0.  -> (advice columns as int lists, selector, sigma mapping)"""
    import random
    rnd = random.Random(seed)
    n, u = c.n, c.usable_rows
    w0, w1, q = [0] * n, [rnd.randrange(FR) for _ in range(n)], [0] * n
    for j in range(0, u - 4, 4):
        a, b, cc = (rnd.randrange(FR) for _ in range(3))
        w0[j:j + 4] = [a, b, cc, (a + b * cc) % FR]
        q[j] = 1
    mapping = [[(j, i) for i in range(n)] for j in range(2)]
    rows1 = rnd.sample(range(u), 200)
    rows0 = rnd.sample(range(u - 4), 200)
    for r1, r0 in zip(rows1, rows0):
        w1[r1] = w0[r0]
        mapping[1][r1], mapping[0][r0] = (0, r0), (1, r1)     # a 2-cycle between cell (1, r1) and cell (0, r0)
    w2 = [rnd.randrange(256) if rnd.random() < 0.7 else 7 for _ in range(n)]     # many repeated values: exercises the leftover rows
    table = [i % 256 for i in range(n)]
    return [w0, w1, w2], [q, table], mapping


@pytest.fixture(scope="module")
def setup():
    zkb.init()
    c = _circuit()
    params = zkb.ParamsKZG.setup(K, plonk.mont(S_TRAPDOOR))
    advice, fixed, mapping = _witness(c, 99)
    pk = plonk.ProvingKey.keygen(params, c, [plonk.mont_vec(f) for f in fixed], mapping)
    yield c, params, pk, [plonk.mont_vec(w) for w in advice]
    pk.free()
    params.close()


def _point(jac):
    return R.g1_jacobian_decode([int(x) for x in jac])


def verify(c, pk, proof):
    """-> (accept, reason)"""
    n, u = c.n, c.usable_rows
    omega = R.omega_for(c.k)
    G = (1, 2)
    # ---- challenges, in the prover's order
    tr = plonk.Transcript()
    for cm in pk.fixed_commitments + pk.sigma_commitments:
        tr.write_point(cm)
    for cm in proof["advice_commitments"]:
        tr.write_point(cm)
    theta = tr.squeeze_challenge()
    for lk in proof["lookup_commitments"]:
        tr.write_point(lk[0])
        tr.write_point(lk[1])
    beta, gamma = tr.squeeze_challenge(), tr.squeeze_challenge()
    for cm in proof["z_commitments"]:
        tr.write_point(cm)
    for lk in proof["lookup_commitments"]:
        tr.write_point(lk[2])
    y = tr.squeeze_challenge()
    for cm in proof["h_commitments"]:
        tr.write_point(cm)
    x = tr.squeeze_challenge()
    evals = proof["evals"]
    for label in evals:                     # dict order == the prover's query order
        tr.write_scalar(evals[label])
    v = tr.squeeze_challenge()
    if (theta, beta, gamma, y, x, v) != tuple(proof["challenges"][k_] for k_ in ("theta", "beta", "gamma", "y", "x", "v")):
        return False, "challenges"
    # ---- vanishing identity at x
    xn = pow(x, n, FR)
    lag = lambda i: pow(omega, i, FR) * (xn - 1) % FR * pow(n * (x - pow(omega, i, FR)) % FR, FR - 2, FR) % FR  # noqa: E731
    l0, l_last = lag(0), lag(u)
    l_active = (1 - sum(lag(i) for i in range(u, n))) % FR

    def expr(e):
        t = e[0]
        if t in ("advice", "fixed"):
            return evals[(t, e[1], e[2])]
        if t == "neg":
            return -expr(e[1]) % FR
        if t == "sum":
            return (expr(e[1]) + expr(e[2])) % FR
        if t == "prod":
            return expr(e[1]) * expr(e[2]) % FR
        raise ValueError(t)

    value = 0
    for g in c.gates:
        value = (value * y + expr(g)) % FR
    cols = c.permutation_columns
    nsets = (len(cols) + c.chunk_len - 1) // c.chunk_len
    last_rot = -(c.blinding_factors + 1)
    z = lambda s, r: evals[("z", s, r)]  # noqa: E731
    terms = [l0 * (1 - z(0, 0)) % FR, l_last * (z(nsets - 1, 0) ** 2 - z(nsets - 1, 0)) % FR]
    for s in range(1, nsets):
        terms.append(l0 * (z(s, 0) - z(s - 1, last_rot)) % FR)
    j = 0
    for s in range(nsets):
        left, right = z(s, 1), z(s, 0)
        for col in cols[s * c.chunk_len:(s + 1) * c.chunk_len]:
            val = evals[("advice", col[1], 0)]
            left = left * (val + beta * evals[("sigma", j, 0)] + gamma) % FR
            right = right * (val + beta * x % FR * pow(plonk.DELTA, j, FR) + gamma) % FR
            j += 1
        terms.append((left - right) * l_active % FR)
    for t in terms:
        value = (value * y + t) % FR
    for li, (inputs, table) in enumerate(c.lookups):
        compress = lambda exprs: __import__("functools").reduce(lambda acc, e: (acc * theta + expr(e)) % FR, exprs, 0)  # noqa: E731
        A, S = compress(inputs), compress(table)
        lz, lzw = evals[("lookup_z", li, 0)], evals[("lookup_z", li, 1)]
        ap, apm, sp = evals[("lookup_a", li, 0)], evals[("lookup_a", li, -1)], evals[("lookup_s", li, 0)]
        for t in (l0 * (1 - lz), l_last * (lz * lz - lz), l_active * (lzw * (ap + beta) * (sp + gamma) - lz * (A + beta) * (S + gamma)),
                  l0 * (ap - sp), l_active * (ap - sp) * (ap - apm)):
            value = (value * y + t) % FR
    hx = sum(pow(xn, i, FR) * evals[("h", i, 0)] for i in range(len(proof["h_commitments"]))) % FR
    if value != hx * (xn - 1) % FR:
        return False, "vanishing identity"
    # ---- KZG openings in the exponent (trapdoor s known to the test)
    for op in proof["openings"]:
        acc_c, acc_e = None, 0
        for i, label in enumerate(op["labels"]):
            cpt = _point(proof["commitment_of"][label])
            acc_c = cpt if i == 0 else R.g1_add(R.g1_mul(acc_c, v), cpt)
            acc_e = evals[label] if i == 0 else (acc_e * v + evals[label]) % FR
            if proof["eval_points"][label] != op["point"]:
                return False, "opening grouping"
        lhs = R.g1_add(acc_c, R.g1_neg(R.g1_mul(G, acc_e)))
        rhs = R.g1_mul(_point(op["witness"]), (S_TRAPDOOR - op["point"]) % FR)
        if lhs != rhs:
            return False, "opening at %x" % op["point"]
    opened = {lab for op in proof["openings"] for lab in op["labels"]}
    if opened != set(evals):
        return False, "an evaluation was never opened"
    return True, "ok"


def test_verifier_accepts_a_proof_made_on_resident_polynomials(setup):
    c, params, pk, advice = setup
    proof = plonk.create_proof(params, pk, advice, np.random.default_rng(1))
    ok, why = verify(c, pk, proof)
    assert ok, why


def _bump(poly, row):
    v = poly.slice(row, 1).to_host()
    v[0, 0] ^= np.uint64(1)
    poly.write(row, v)


@pytest.mark.parametrize("where", ["after_advice", "after_z", "after_h"])
def test_verifier_rejects_when_one_device_value_is_wrong(setup, where):
    """one advice cell / one row of the permutation product / one value of the quotient numerator on the coset, flipped on the device"""
    c, params, pk, advice = setup
    row = {"after_advice": 16, "after_z": 100, "after_h": 12345}[where]
    proof = plonk.create_proof(params, pk, advice, np.random.default_rng(2), hooks={where: lambda polys: _bump(polys[0], row)})
    ok, why = verify(c, pk, proof)
    assert not ok and why == "vanishing identity", why


def test_verifier_rejects_a_wrong_commitment_or_evaluation(setup):
    c, params, pk, advice = setup
    proof = plonk.create_proof(params, pk, advice, np.random.default_rng(3))
    assert verify(c, pk, proof)[0]
    bad = dict(proof)
    bad["evals"] = dict(proof["evals"])
    lab = ("advice", 1, 0)
    bad["evals"][lab] = (bad["evals"][lab] + 1) % FR
    assert not verify(c, pk, bad)[0]
    # a different polynomial's commitment in place of an h piece: the challenges (recomputed) no longer match the proof's
    bad = dict(proof)
    bad["h_commitments"] = [proof["h_commitments"][1], proof["h_commitments"][0]] + proof["h_commitments"][2:]
    assert not verify(c, pk, bad)[0]
    # an opening witness of another point
    bad = dict(proof)
    bad["openings"] = [dict(o) for o in proof["openings"]]
    bad["openings"][0]["witness"], bad["openings"][1]["witness"] = proof["openings"][1]["witness"], proof["openings"][0]["witness"]
    ok, why = verify(c, pk, bad)
    assert not ok and why.startswith("opening")


def test_proving_key_is_resident_second_proof_uploads_witness_only(setup):
    """fixed / sigma polynomials and cosets, l_0 / l_last / l_active stay in HBM across proofs (the IVC loop,
    /root/reference/aggregator/src/wrapper.rs:884-900): a second create_proof moves only the advice columns (plus the blinding rows
    of the product columns and the lowered row programs, a few KB) host -> device."""
    c, params, pk, advice = setup
    lib = zkb.lib()
    plonk.create_proof(params, pk, advice, np.random.default_rng(4))
    h2d = ctypes.c_uint64(0)
    lib.zkb_transfer_stats(ctypes.byref(h2d), 1)
    proof = plonk.create_proof(params, pk, advice, np.random.default_rng(5))
    lib.zkb_transfer_stats(ctypes.byref(h2d), 0)
    witness_bytes = c.num_advice * c.n * 32
    assert witness_bytes <= h2d.value <= witness_bytes + (64 << 10), (h2d.value, witness_bytes)
    resident = (len(pk.fixed_polys) + 3 * len(pk.sigma_polys)) * c.n * 32 + (len(pk.fixed_cosets) + len(pk.sigma_cosets) + 4) * 4 * c.n * 32
    assert h2d.value < resident / 2                 # nowhere near re-uploading the key
    assert verify(c, pk, proof)[0]


def test_proving_key_file_round_trip_straight_into_hbm(setup):
    """ProvingKey.write -> ProvingKey.read (zkb_poly_load_file: raw limbs from the file into HBM through pinned staging):
    the proof made with the loaded key is byte-identical to the one made with the generated key (same prover randomness)."""
    c, params, pk, advice = setup
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "pk.bin")
        pk.write(path)
        pk2 = plonk.ProvingKey.read(path, c)
    try:
        for a, b in zip(pk.fixed_cosets + pk.sigma_cosets + pk.sigma_polys, pk2.fixed_cosets + pk2.sigma_cosets + pk2.sigma_polys):
            assert (a.to_host() == b.to_host()).all()
        p1 = plonk.create_proof(params, pk, advice, np.random.default_rng(6))
        p2 = plonk.create_proof(params, pk2, advice, np.random.default_rng(6))
        for key in ("advice_commitments", "z_commitments", "h_commitments"):
            assert all((x == y).all() for x, y in zip(p1[key], p2[key]))
        assert p1["evals"] == p2["evals"]
        assert all((o1["witness"] == o2["witness"]).all() for o1, o2 in zip(p1["openings"], p2["openings"]))
        assert verify(c, pk2, p2)[0]
    finally:
        pk2.free()
    with pytest.raises(ValueError):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "pk.bin")
            pk.write(path)
            plonk.ProvingKey.read(path, c._replace(num_advice=4))


def test_lookup_input_outside_the_table_is_a_constraint_failure(setup):
    """upstream: permute_expression_pair returns Error::ConstraintSystemFailure; here the C ABI's ZKB_ERR_ARG surfaces as ZkbError"""
    c, params, pk, advice = setup
    bad = [a.copy() for a in advice]
    bad[2][5] = plonk.mont(300)
    with pytest.raises(zkb.ZkbError):
        plonk.create_proof(params, pk, bad, np.random.default_rng(8))


@pytest.mark.parametrize("u,distinct,big", [(1000, 1000, True), (5000, 33, True), (1 << 16, 256, False), ((1 << 18) - 6, 4096, False)])
def test_permute_expression_pair_device_vs_oracle(oracle, u, distinct, big):
    from test_oracle import lookup_case
    from util import ints_to_limbs
    inp, tab = lookup_case(3 * u + distinct, u, distinct, big)
    n = 1 << (u - 1).bit_length()
    pad = [0] * (n - u)
    a = ints_to_limbs([R.to_mont(x, R.FR) for x in inp + pad])
    t = ints_to_limbs([R.to_mont(x, R.FR) for x in tab + pad])
    want = oracle.permute_expression_pair(a, t, u)
    A, T = zkb.Polynomial(a), zkb.Polynomial(t)
    ap, sp = plonk._permute_expression_pair(A, T, u)
    got_a, got_s = ap.to_host(), sp.to_host()
    assert (got_a[:u] == want[0]).all() and (got_s[:u] == want[1]).all()
    assert not got_a[u:].any() and not got_s[u:].any()
