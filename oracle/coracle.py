"""ctypes binding of oracle/libzkb_oracle.so (TEST INFRASTRUCTURE — see oracle/zkb_oracle.c header).

Arrays are numpy ``uint64`` with the in-memory layout halo2curves uses: Fr/Fq = 4 LE limbs (Montgomery),
G1Affine = 8 limbs (x, y), G1 = 12 limbs (x, y, z Jacobian).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzkb_oracle.so")
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zkb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libzkb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def use_native() -> bool:
    """bench.py's CPU-baseline legs only: recompile the oracle ON THIS HOST with -O3 -march=native into
    oracle/_native/ and bind that build from now on.  The portable build that travels with the repository stays what the tests
    use (a -march=native binary built in one container may not run on another host).  Returns False (and keeps the portable
    build) if the compiler is missing or fails."""
    global _lib, _LIB_PATH
    out_dir = os.path.join(_HERE, "_native")
    out = os.path.join(out_dir, "libzkb_oracle_native.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-pthread", "-fPIC", "-fvisibility=hidden", "-std=c11", "-shared", "-o", out,
                               os.path.join(_HERE, "zkb_oracle.c"), "-lm", "-lpthread"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return False
    _LIB_PATH = out
    _lib = None
    return True


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.zko_num_threads.restype = ctypes.c_int
        _lib.zko_g1_is_on_curve.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _new(*shape):
    return np.zeros(shape, dtype=np.uint64)


def num_threads() -> int:
    return lib().zko_num_threads()


# ---- field helpers -----------------------------------------------------------------------------
def vec_op(field: str, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    o = _new(a.shape[0], 4)
    lib().zko_vec_op(
        ctypes.c_int(0 if field == "fr" else 1),
        ctypes.c_int({"mul": 0, "add": 1, "sub": 2}[op]),
        _p(a), _p(b), _p(o), ctypes.c_size_t(a.shape[0]),
    )
    return o


def fr_to_mont(a: np.ndarray) -> np.ndarray:
    """canonical limbs (n,4) -> Montgomery limbs"""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    r2 = np.tile(np.array([0x1BB8E645AE216DA7, 0x53FE3AB1E35C59E3, 0x8C49833D53BB8085, 0x0216D0B17F4E44A5],
                          dtype=np.uint64), (a.shape[0], 1))
    return vec_op("fr", "mul", a, r2)


def fr_from_mont(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    one = np.tile(np.array([1, 0, 0, 0], dtype=np.uint64), (a.shape[0], 1))
    return vec_op("fr", "mul", a, one)


def fr_inner_product(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    o = _new(4)
    lib().zko_fr_inner_product(_p(a), _p(b), ctypes.c_size_t(a.shape[0]), _p(o))
    return o


def fr_omega(k: int) -> np.ndarray:
    o = _new(4)
    lib().zko_fr_omega(ctypes.c_uint32(k), _p(o))
    return o


def fr_inv(a: np.ndarray) -> np.ndarray:
    o = _new(4)
    lib().zko_fr_inv(_p(np.ascontiguousarray(a, dtype=np.uint64)), _p(o))
    return o


# ---- FFT / domain ---------------------------------------------------------------------------------
def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int = 0) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    assert a.shape[0] == 1 << log_n
    lib().zko_best_fft(_p(a), _p(np.ascontiguousarray(omega, dtype=np.uint64)), ctypes.c_uint32(log_n),
                       ctypes.c_int(threads))
    return a


def lagrange_to_coeff(a: np.ndarray, k: int, threads: int = 0) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    assert a.shape[0] == 1 << k
    lib().zko_lagrange_to_coeff(_p(a), ctypes.c_uint32(k), ctypes.c_int(threads))
    return a


def coeff_to_lagrange(a: np.ndarray, k: int, threads: int = 0) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    lib().zko_coeff_to_lagrange(_p(a), ctypes.c_uint32(k), ctypes.c_int(threads))
    return a


def coeff_to_extended(a: np.ndarray, k: int, ek: int, threads: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    assert a.shape[0] == 1 << k
    o = _new(1 << ek, 4)
    lib().zko_coeff_to_extended(_p(a), _p(o), ctypes.c_uint32(k), ctypes.c_uint32(ek), ctypes.c_int(threads))
    return o


def extended_to_coeff(a: np.ndarray, k: int, ek: int, threads: int = 0) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    assert a.shape[0] == 1 << ek
    lib().zko_extended_to_coeff(_p(a), ctypes.c_uint32(k), ctypes.c_uint32(ek), ctypes.c_int(threads))
    return a


# ---- G1 / MSM -----------------------------------------------------------------------------------------
def g1_generator() -> np.ndarray:
    o = _new(8)
    lib().zko_g1_generator(_p(o))
    return o


def g1_to_affine(jac: np.ndarray) -> np.ndarray:
    o = _new(8)
    lib().zko_g1_to_affine(_p(np.ascontiguousarray(jac, dtype=np.uint64)), _p(o))
    return o


def g1_is_on_curve(aff: np.ndarray) -> bool:
    return bool(lib().zko_g1_is_on_curve(_p(np.ascontiguousarray(aff, dtype=np.uint64))))


def g1_mul(base_aff: np.ndarray, scalar_mont: np.ndarray) -> np.ndarray:
    o = _new(8)
    lib().zko_g1_mul(_p(np.ascontiguousarray(base_aff, dtype=np.uint64)),
                     _p(np.ascontiguousarray(scalar_mont, dtype=np.uint64)), _p(o))
    return o


def g1_add_affine(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    o = _new(8)
    lib().zko_g1_add_affine(_p(np.ascontiguousarray(a, dtype=np.uint64)),
                            _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(o))
    return o


def g1_fixed_base_mul(scalars_mont: np.ndarray, threads: int = 0) -> np.ndarray:
    s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    o = _new(s.shape[0], 8)
    lib().zko_g1_fixed_base_mul(_p(s), ctypes.c_size_t(s.shape[0]), _p(o), ctypes.c_int(threads))
    return o


def g1_batch_normalize(jac: np.ndarray) -> np.ndarray:
    j = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    o = _new(j.shape[0], 8)
    lib().zko_g1_batch_normalize(_p(j), ctypes.c_size_t(j.shape[0]), _p(o))
    return o


def fr_eval_polynomial(coeffs: np.ndarray, x: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    o = _new(4)
    lib().zko_fr_eval_polynomial(_p(c), ctypes.c_size_t(c.shape[0]), _p(np.ascontiguousarray(x, dtype=np.uint64)), _p(o))
    return o


def fr_kate_division(coeffs: np.ndarray, b: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    o = _new(max(c.shape[0] - 1, 0), 4)
    lib().zko_fr_kate_division(_p(c), ctypes.c_size_t(c.shape[0]), _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(o))
    return o


def fr_batch_invert(a: np.ndarray) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    lib().zko_fr_batch_invert(_p(a), ctypes.c_size_t(a.shape[0]))
    return a


def permute_expression_pair(input_: np.ndarray, table: np.ndarray, usable_rows: int):
    """-> (permuted_input, permuted_table), usable_rows elements each; raises ValueError when an input value is not in the table"""
    a = np.ascontiguousarray(input_, dtype=np.uint64).reshape(-1, 4)
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
    oa, ot = _new(usable_rows, 4), _new(usable_rows, 4)
    lib().zko_permute_expression_pair.restype = ctypes.c_int
    rc = lib().zko_permute_expression_pair(_p(a), _p(t), ctypes.c_size_t(usable_rows), _p(oa), _p(ot))
    if rc != 0:
        raise ValueError("ConstraintSystemFailure" if rc == -1 else "malformed lookup")
    return oa, ot


def g1_fft_naive(points_aff: np.ndarray, omega: np.ndarray) -> np.ndarray:
    """best_fft with G = G1 by the DFT definition (O(n^2) scalar multiplications)."""
    p = np.ascontiguousarray(points_aff, dtype=np.uint64).reshape(-1, 8)
    o = _new(p.shape[0], 8)
    lib().zko_g1_fft_naive(_p(p), ctypes.c_size_t(p.shape[0]), _p(np.ascontiguousarray(omega, dtype=np.uint64)), _p(o))
    return o


def kzg_setup(k: int, s_mont: np.ndarray, threads: int = 0):
    """ParamsKZG::setup, G1 side: (g, g_lagrange), each (2^k, 8)."""
    g, gl = _new(1 << k, 8), _new(1 << k, 8)
    rc = lib().zko_kzg_setup(ctypes.c_uint(k), _p(np.ascontiguousarray(s_mont, dtype=np.uint64)), _p(g), _p(gl), ctypes.c_int(threads))
    if rc:
        raise ValueError("s is an n-th root of unity")
    return g, gl


def msm_naive(scalars_mont: np.ndarray, bases_aff: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases_aff, dtype=np.uint64).reshape(-1, 8)
    assert s.shape[0] == b.shape[0]
    o = _new(12)
    lib().zko_msm_naive(_p(s), _p(b), ctypes.c_size_t(s.shape[0]), _p(o))
    return o


def best_multiexp(scalars_mont: np.ndarray, bases_aff: np.ndarray, threads: int = 0) -> np.ndarray:
    s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases_aff, dtype=np.uint64).reshape(-1, 8)
    assert s.shape[0] == b.shape[0]
    o = _new(12)
    lib().zko_best_multiexp(_p(s), _p(b), ctypes.c_size_t(s.shape[0]), ctypes.c_int(threads), _p(o))
    return o


def graph_evaluate(calcs: np.ndarray, num_intermediates: int, constants: np.ndarray, rotations, fixed, advice, instance,
                   challenges, beta, gamma, theta, y, rot_scale: int, values: np.ndarray, threads: int = 0) -> np.ndarray:
    """GraphEvaluator::evaluate for every row (evaluate_h's loop).  calcs: (ncalc, 11) uint32 in zkb_calculation's layout;
    columns: lists of (isize, 4) arrays; scalars (4,) or None; values: previous values, a new array is returned."""
    calcs = np.ascontiguousarray(calcs, dtype=np.uint32).reshape(-1, 11)
    consts = np.ascontiguousarray(constants, dtype=np.uint64).reshape(-1, 4)
    rots = np.ascontiguousarray(rotations, dtype=np.int32)
    out = np.array(values, dtype=np.uint64, copy=True).reshape(-1, 4)
    keep = []

    def table(cols):
        arrs = [np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4) for c in cols]
        keep.append(arrs)
        for a in arrs:
            assert a.shape[0] == out.shape[0]
        t = (_u64p * max(len(arrs), 1))(*[_p(a) for a in arrs])
        keep.append(t)
        return t

    def scalar(s):
        if s is None:
            return None
        a = np.ascontiguousarray(s, dtype=np.uint64).reshape(-1)
        keep.append(a)
        return _p(a)

    ch = np.ascontiguousarray(challenges if challenges is not None else np.zeros((0, 4)), dtype=np.uint64).reshape(-1, 4)
    rc = lib().zko_graph_evaluate(calcs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_size_t(calcs.shape[0]),
                                  ctypes.c_uint32(num_intermediates), _p(consts) if consts.size else None,
                                  rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), table(fixed), table(advice), table(instance),
                                  _p(ch) if ch.size else None, scalar(beta), scalar(gamma), scalar(theta), scalar(y),
                                  ctypes.c_int32(rot_scale), _p(out), ctypes.c_size_t(out.shape[0]), ctypes.c_int(threads))
    if rc != 0:
        raise ValueError("malformed graph")
    return out
