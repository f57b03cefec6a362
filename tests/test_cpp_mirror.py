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
