"""Generates tests/golden/graph_kats.json: quotient-evaluation vectors.

Run:  python tests/golden/gen_golden_graph.py
Inputs are seeded random columns over an extended domain of 32 rows (k = 3, extended_k = 5, rot_scale = 4); the expected
values come from tests/graph_cases.py's `custom_gates_value` — the expressions evaluated by their definition with Python
integers, value = previous * y + gate(row) per gate — not from a GraphEvaluator, the C oracle or the code under test.
Values are canonical integers in hex."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", ".."), os.path.join(HERE, "..")]
import graph_cases as G  # noqa: E402


def main():
    out = []
    for seed, isize, rot_scale, ngates in [(0xA1, 32, 4, 1), (0xA2, 32, 4, 4), (0xA3, 16, 1, 3)]:
        c = G.random_case(seed, isize, rot_scale, ngates=ngates)
        out.append({"isize": isize, "rot_scale": rot_scale, "gates": c["gates"],
                    "cols": {k: [[hex(x) for x in col] for col in v] for k, v in c["cols"].items()},
                    "challenges": [hex(x) for x in c["challenges"]], "y": hex(c["y"]), "prev": [hex(x) for x in c["prev"]],
                    "expected": [hex(x) for x in G.case_expected(c)]})
    with open(os.path.join(HERE, "graph_kats.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote graph_kats.json")


if __name__ == "__main__":
    main()
