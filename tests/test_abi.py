"""The C-ABI library loads without a GPU, exports every symbol include/zkb200.h declares, fails loudly on
compute calls when no CUDA device exists, and its host-only helpers agree with the oracle."""
import importlib
import os

import numpy as np
import pytest

from oracle import pyref as R
from util import limbs_to_int, random_field

zkb = importlib.import_module("zksnap-circuits-halo2_b200")


def test_library_exports_every_declared_symbol():
    lib = zkb.lib()
    syms = zkb.header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_header_cites_reference_call_sites():
    text = open(os.path.join(os.path.dirname(__file__), "..", "include", "zkb200.h")).read()
    for cite in ("aggregator/src/wrapper.rs:106-109", "aggregator/src/wrapper.rs:129-137", "voter/benches/voter_circuit.rs"):
        assert cite in text


def test_omega_matches_oracle(oracle):
    for k in (0, 1, 3, 13, 15, 22, 24, 28):
        assert limbs_to_int(zkb.omega(k)) == R.to_mont(R.omega_for(k), R.FR)
        assert (zkb.omega(k) == oracle.fr_omega(k)).all()


def test_g1_sum_host_combine(oracle):
    """zkb_g1_sum is host code (the multi-GPU fold); check it against the oracle without a GPU."""
    s = random_field(5, 1)
    pts = oracle.g1_fixed_base_mul(s)
    jac = np.zeros((5, 12), dtype=np.uint64)
    jac[:, :8] = pts
    jac[:, 8:] = np.array(R.limbs(R.FQ_R), dtype=np.uint64)
    jac[3] = 0  # an identity entry (z = 0)
    want = np.zeros(8, dtype=np.uint64)
    for i in (0, 1, 2, 4):
        want = oracle.g1_add_affine(want, pts[i])
    got = zkb.g1_sum(jac)
    assert (got[:8] == want).all() and limbs_to_int(got[8:]) == R.FQ_R
    ident = zkb.g1_sum(np.zeros((0, 12), dtype=np.uint64))
    assert not ident[8:].any() and limbs_to_int(ident[4:8]) == R.FQ_R


def test_evaluation_domain_geometry():
    for (j, k) in ((4, 13), (4, 15), (4, 22), (3, 5), (5, 3), (2, 4)):
        d = zkb.EvaluationDomain(j, k)
        ref = R.EvaluationDomain(j, k)
        assert d.extended_k == ref.extended_k and d.quotient_poly_degree == ref.quotient_poly_degree


def test_no_cpu_fallback_without_device():
    if zkb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(zkb.ZkbError) as e:
        zkb.init()
    assert e.value.code == -4
    a = random_field(8, 1)
    with pytest.raises(zkb.ZkbError):
        zkb.best_fft(a, zkb.omega(3), 3)
    with pytest.raises(zkb.ZkbError):
        zkb.best_multiexp(a, np.zeros((8, 8), dtype=np.uint64))


def test_argument_checks_mirror_rust_asserts():
    a = random_field(8, 1)
    with pytest.raises(AssertionError):
        zkb.best_fft(a, zkb.omega(4), 4)  # a.len() != 1 << log_n
    with pytest.raises(AssertionError):
        zkb.best_multiexp(a, np.zeros((7, 8), dtype=np.uint64))


def test_graph_structs_have_the_c_layout(tmp_path):
    """The ctypes mirrors of zkb_value_source / zkb_calculation / zkb_graph / zkb_graph_inputs (evaluation.py) against what the
    C compiler lays out for include/zkb200.h — the oracle and the emulator read the same 11-u32 calculation records."""
    import ctypes
    import importlib
    import subprocess
    ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "zkb200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(zkb_value_source), sizeof(zkb_calculation),'
                   ' offsetof(zkb_calculation, b), sizeof(zkb_graph), offsetof(zkb_graph, constants), offsetof(zkb_graph, rotations),'
                   ' sizeof(zkb_graph_inputs), offsetof(zkb_graph_inputs, challenges), offsetof(zkb_graph_inputs, beta),'
                   ' offsetof(zkb_graph_inputs, rot_scale)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    c = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    py = [ctypes.sizeof(ev.CValueSource), ctypes.sizeof(ev.CCalculation), ev.CCalculation.b.offset, ctypes.sizeof(ev.CGraph),
          ev.CGraph.constants.offset, ev.CGraph.rotations.offset, ctypes.sizeof(ev.CGraphInputs), ev.CGraphInputs.challenges.offset,
          ev.CGraphInputs.beta.offset, ev.CGraphInputs.rot_scale.offset]
    assert c == py and c[0] == 12 and c[1] == 44


def test_t_evaluations_inverse_is_the_inverse_of_the_vanishing_polynomial_on_the_coset():
    """EvaluationDomain.t_evaluations_inverse (host integers, tiled): every entry times (x^n - 1) at its coset point is one."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyref as R
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    for j, k in [(4, 3), (4, 5), (3, 4), (5, 2)]:
        d = zkb.EvaluationDomain(j, k)
        t = d.t_evaluations_inverse()
        assert t.shape == (d.extended_len(), 4)
        zeta = pow(7, 2 * (R.FR - 1) // 3, R.FR)
        w = R.omega_for(d.extended_k)
        for i in range(d.extended_len()):
            x = zeta * pow(w, i, R.FR) % R.FR
            ti = R.from_mont(sum(int(v) << (64 * q) for q, v in enumerate(t[i])), R.FR)
            assert ti * (pow(x, 1 << k, R.FR) - 1) % R.FR == 1
