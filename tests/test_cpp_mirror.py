"""The C++ host mirror of the halo2 interface (include/zkb200_halo2.hpp) compiles against the C ABI and behaves."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "zksnap-circuits-halo2_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", PKG, "-lzkb200", f"-Wl,-rpath,{PKG}"])


def test_cpp_mirror_cpu():
    _build()
    out = subprocess.check_output([EXE], text=True)
    assert "cpu ok" in out


@pytest.mark.gpu
def test_cpp_mirror_gpu():
    _build()
    out = subprocess.check_output([EXE, "gpu"], text=True)
    assert "gpu ok" in out


def _build_example():
    src = os.path.join(ROOT, "examples", "prover_ops.cpp")
    exe = os.path.join(ROOT, "examples", "prover_ops")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", PKG, "-lzkb200", f"-Wl,-rpath,{PKG}"])
    return exe


def test_cpp_example_compiles_and_fails_loudly_without_gpu():
    exe = _build_example()
    import importlib
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    if zkb.device_count() == 0:
        r = subprocess.run([exe, "10"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_example_prover_ops_gpu():
    """examples/prover_ops.cpp: setup -> commit_lagrange -> lagrange_to_coeff -> commit -> coeff_to_extended ->
    extended_to_coeff through the compiled C++ host mirror, with its own consistency checks."""
    exe = _build_example()
    out = subprocess.check_output([exe, "16"], text=True)
    assert "k=16 ok" in out


def test_cpp_graph_builders_match_the_python_mirror():
    """The C++ host mirror's add_expression / custom_gates_graph / permutation_graph / lookup_graph (and its host Fr arithmetic
    for the constants) produce exactly the graphs of zksnap-circuits-halo2_b200/evaluation.py, which the oracle / emulator / GPU
    tests pin to the formulas: constants, rotations and every calculation record are compared."""
    import importlib
    ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
    src = os.path.join(ROOT, "tests", "cpp", "graph_mirror_dump.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "graph_mirror_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lzkb200",
                           f"-Wl,-rpath,{PKG}"])
    got, cur = {}, None
    for line in subprocess.check_output([exe], text=True).splitlines():
        tag, *rest = line.split()
        if tag == "graph":
            cur = got.setdefault(rest[0], {"ni": int(rest[1]), "const": [], "rot": [], "calc": []})
        elif tag == "const":
            cur["const"].append([int(x, 16) for x in rest])
        elif tag == "rot":
            cur["rot"].append(int(rest[0]))
        else:
            cur["calc"].append([int(x) for x in rest])

    def gate(adv, sel):
        a, b, c, d = (("advice", adv, r) for r in range(4))
        return ("prod", ("fixed", sel, 0), ("sum", ("sum", a, ("prod", b, c)), ("neg", d)))

    gates = [gate(0, 0), gate(1, 1),
             ("sum", ("scaled", ("instance", 0, -1), 7), ("neg", ("const", 5))),
             ("prod", ("const", 2), ("prod", ("advice", 0, 0), ("advice", 0, 0))),
             ("sum", ("const", 0), ("prod", ("const", 1), ("challenge", 1))),
             ("sum", ("neg", ("fixed", 1, 2)), ("advice", 2, 0)),
             ("scaled", ("advice", 1, 1), 1)]
    want = {"custom_gates": ev.custom_gates_graph(gates),
            "permutation": ev.permutation_graph([("advice", i) for i in range(5)], 2, -4, ("fixed", 0), ("fixed", 1), ("fixed", 2), ("fixed", 3),
                                                [("fixed", 4 + i) for i in range(5)], [("advice", 5 + i) for i in range(3)]),
            "lookup": ev.lookup_graph([("advice", 0, 0), ("prod", ("advice", 1, 0), ("advice", 0, 1))],
                                      [("fixed", 3, 0), ("scaled", ("fixed", 3, -1), 3)], ("fixed", 0), ("fixed", 1), ("fixed", 2),
                                      ("advice", 2), ("advice", 3), ("advice", 4))}
    assert set(got) == set(want)
    for name, g in want.items():
        c = got[name]
        assert c["ni"] == g.num_intermediates and c["rot"] == g.rotations, name
        assert c["const"] == [[int(x) for x in row] for row in ev._mont_limbs(g.constants)], name
        assert c["calc"] == [[int(x) for x in row] for row in g.calc_array()], name


def test_cpp_host_fr_arithmetic_matches_python_integers():
    """halo2::fr::{mul, add, neg, from_u64} (the host arithmetic behind a graph's constants) on random and edge operands."""
    import random
    r = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    R = (1 << 256) % r
    src = os.path.join(ROOT, "tests", "cpp", "fr_host_arith.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "fr_host_arith")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lzkb200",
                           f"-Wl,-rpath,{PKG}"])
    rnd = random.Random(5)
    edge = [0, 1, 2, r - 1, r - 2, R, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 192, r >> 1, (r >> 1) + 1]
    vals = edge + [rnd.randrange(r) for _ in range(60)]
    limbs = lambda x: " ".join("%x" % ((x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF) for i in range(4))  # noqa: E731
    lines, want = [], []
    for a in vals:
        for b in rnd.sample(vals, 6) + [a]:
            lines.append(f"mul {limbs(a)} {limbs(b)}")
            want.append(a * b * pow(R, -1, r) % r)       # Montgomery product
            lines.append(f"add {limbs(a)} {limbs(b)}")
            want.append((a + b) % r)
        lines.append(f"neg {limbs(a)} {limbs(0)}")
        want.append(-a % r)
        lines.append(f"u64 {limbs(a & 0xFFFFFFFFFFFFFFFF)} {limbs(0)}")
        want.append((a & 0xFFFFFFFFFFFFFFFF) * R % r)
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    got = [sum(int(x, 16) << (64 * i) for i, x in enumerate(l.split())) for l in out if l.strip()]
    assert got == want
