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
