"""Generates tests/golden/widened_kats.json: first-principles vectors for the rows widened beyond MSM / NTT
(ParamsKZG::setup, eval_polynomial, kate_division, BatchInvert, batch_normalize, best_fft over G1, prefix products).

Run:  python tests/golden/gen_golden_widened.py
Every number comes from Python integer arithmetic by the definitions (oracle/pyref.py supplies only the field
constants and the textbook affine group law) — not from the C oracle and not from the code under test.
Values are canonical integers in hex; tests convert to the Montgomery in-memory layout.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyref as R  # noqa: E402

hx = hex
P, Q = R.FR, R.FQ


def inv(a, m):
    return pow(a, m - 2, m)


def pt(p):
    return [hx(p[0]), hx(p[1])] if p else None


def main():
    rnd = random.Random(0xC0FFEE)
    G = R.G1_GENERATOR
    out = {}

    # ParamsKZG::setup, G1 side: g[i] = [s^i]G, g_lagrange[i] = [l_i(s)]G, l_i(s) = w^i (s^n - 1) / (n (s - w^i))
    k, n = 3, 8
    s = rnd.randrange(2, P)
    w = R.omega_for(k)
    c = (pow(s, n, P) - 1) * inv(n, P) % P
    lag = [pow(w, i, P) * c % P * inv((s - pow(w, i, P)) % P, P) % P for i in range(n)]
    assert sum(lag) % P == 1
    out["kzg_setup"] = {"k": k, "s": hx(s), "g": [pt(R.g1_mul(G, pow(s, i, P))) for i in range(n)],
                        "g_lagrange": [pt(R.g1_mul(G, l)) for l in lag]}

    # eval_polynomial / kate_division: a(X) = q(X) (X - b) + a(b)
    a = [rnd.randrange(P) for _ in range(20)]
    a[3] = 0
    b = rnd.randrange(P)
    ev = sum(cf * pow(b, i, P) for i, cf in enumerate(a)) % P
    q = [0] * (len(a) - 1)
    q[-1] = a[-1]
    for i in range(len(a) - 2, 0, -1):
        q[i - 1] = (a[i] + b * q[i]) % P
    chk = [0] * len(a)
    for i, cf in enumerate(q):
        chk[i + 1] = (chk[i + 1] + cf) % P
        chk[i] = (chk[i] - cf * b) % P
    chk[0] = (chk[0] + ev) % P
    assert chk == a
    out["poly"] = {"coeffs": [hx(x) for x in a], "point": hx(b), "eval": hx(ev), "kate_quotient": [hx(x) for x in q]}

    # BatchInvert (zeros stay zero) and the grand-product scan z[0] = 1, z[i] = prod_{j<i} v[j]
    v = [rnd.randrange(1, P) for _ in range(12)]
    v[2] = 0
    v[7] = 1
    v[9] = P - 1
    out["batch_invert"] = {"in": [hx(x) for x in v], "out": [hx(inv(x, P) if x else 0) for x in v]}
    u = [rnd.randrange(1, P) for _ in range(70)]      # longer than one device chunk (64)
    z, acc = [], 1
    for x in u:
        z.append(acc)
        acc = acc * x % P
    out["prefix_product"] = {"in": [hx(x) for x in u], "out": [hx(x) for x in z]}

    # Curve::batch_normalize: Jacobian (x z^2, y z^3, z) -> affine; z = 0 is the identity
    pts = [R.g1_mul(G, rnd.randrange(1, P)) for _ in range(6)]
    zs = [1, rnd.randrange(2, Q), 0, Q - 1, rnd.randrange(2, Q), 2]
    jac = []
    for p, zz in zip(pts, zs):
        jac.append([hx(0), hx(1), hx(0)] if zz == 0 else [hx(p[0] * zz * zz % Q), hx(p[1] * pow(zz, 3, Q) % Q), hx(zz)])
    out["batch_normalize"] = {"jacobian": jac, "affine": [pt(p) if zz else None for p, zz in zip(pts, zs)]}

    # best_fft over G1 by the DFT definition, n = 4
    k4 = 2
    w4 = R.omega_for(k4)
    p4 = [R.g1_mul(G, rnd.randrange(1, P)) for _ in range(4)]
    p4[2] = None  # an identity input
    res = []
    for i in range(4):
        acc = None
        for j in range(4):
            acc = R.g1_add(acc, R.g1_mul(p4[j], pow(w4, i * j, P)) if p4[j] else None)
        res.append(acc)
    out["g1_fft"] = {"k": k4, "omega": hx(w4), "in": [pt(p) for p in p4], "out": [pt(p) for p in res]}

    with open(os.path.join(HERE, "widened_kats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "widened_kats.json"))


if __name__ == "__main__":
    main()
