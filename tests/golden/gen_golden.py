"""Generates tests/golden/*.json from the pure-Python big-int twin (oracle/pyref.py).

Run:  python tests/golden/gen_golden.py
The reference repo holds no golden vectors for the MSM/NTT path and cannot be executed here (Rust, no
cargo, un-vendored deps), so these fixtures are first-principles values: every number is produced by
Python integer arithmetic from the definitions (DFT sum, Horner evaluation on the coset, double-and-add).
Values are canonical integers in hex; tests convert to the Montgomery in-memory layout.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyref as R  # noqa: E402


def hx(x):
    return hex(x)


def main():
    rnd = random.Random(0x5EED)
    out = {}

    # --- best_fft: DFT by definition (O(n^2) sums), sizes 2^1..2^6
    ffts = []
    for k in range(1, 7):
        n = 1 << k
        a = [rnd.randrange(R.FR) for _ in range(n)]
        if k == 3:
            a = list(range(1, 9))  # the SURVEY.md §8c KAT
        w = R.omega_for(k)
        ffts.append({"k": k, "omega": hx(w), "in": [hx(x) for x in a], "out": [hx(x) for x in R.dft_naive(a, w)]})
    out["best_fft"] = ffts

    # --- lagrange_to_coeff: inverse DFT by definition
    l2c = []
    for k in (2, 5):
        n = 1 << k
        a = [rnd.randrange(R.FR) for _ in range(n)]
        wi = pow(R.omega_for(k), R.FR - 2, R.FR)
        ninv = pow(n, R.FR - 2, R.FR)
        l2c.append({"k": k, "in": [hx(x) for x in a], "out": [hx(x * ninv % R.FR) for x in R.dft_naive(a, wi)]})
    out["lagrange_to_coeff"] = l2c

    # --- coeff_to_extended: Horner evaluation of the polynomial at zeta * omega_ext^i
    c2e = []
    for (j, k) in ((4, 2), (4, 4), (3, 3), (5, 3)):
        d = R.EvaluationDomain(j, k)
        a = [1, 2, 3, 4] if (j, k) == (4, 2) else [rnd.randrange(R.FR) for _ in range(1 << k)]
        ev = []
        for i in range(d.extended_len()):
            x = R.FR_ZETA * pow(d.extended_omega, i, R.FR) % R.FR
            acc = 0
            for c in reversed(a):
                acc = (acc * x + c) % R.FR
            ev.append(acc)
        c2e.append({"j": j, "k": k, "extended_k": d.extended_k, "in": [hx(x) for x in a], "out": [hx(x) for x in ev]})
    out["coeff_to_extended"] = c2e

    # --- MSM: double-and-add sums
    G = R.G1_GENERATOR
    msms = []
    bases = [R.g1_mul(G, i + 1) for i in range(8)]
    scal = [i * 0x1234567 + 5 for i in range(8)]
    msms.append({"name": "survey_kat", "scalars": [hx(s) for s in scal], "bases": [[hx(b[0]), hx(b[1])] for b in bases],
                 "out": [hx(c) for c in R.msm_naive(scal, bases)]})
    bases = [R.g1_mul(G, rnd.randrange(1, R.FR)) for _ in range(40)]
    scal = [rnd.randrange(R.FR) for _ in range(40)]
    scal[0], scal[1], scal[2], scal[3] = 0, 1, R.FR - 1, (1 << 253)
    bases[5] = bases[4]                      # repeated base
    bases[7] = R.g1_neg(bases[6]); scal[7] = scal[6]   # cancelling pair
    bases[9] = None                          # identity base
    msms.append({"name": "random40_edges", "scalars": [hx(s) for s in scal],
                 "bases": [[hx(b[0]), hx(b[1])] if b else None for b in bases],
                 "out": [hx(c) for c in R.msm_naive(scal, bases)]})
    out["msm"] = msms

    # --- scalar multiples of the generator
    out["g1_mul"] = [{"s": hx(s), "out": ([hx(c) for c in R.g1_mul(G, s)] if R.g1_mul(G, s) else None)}
                     for s in (1, 2, 3, R.FR - 1, R.FR, 0xDEADBEEF, rnd.randrange(R.FR))]

    with open(os.path.join(HERE, "hotpath_kats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "hotpath_kats.json"))


if __name__ == "__main__":
    main()
