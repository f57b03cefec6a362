"""ctypes binding of the CPU emulator of the device code (test infrastructure, see csrc/hostemu.cu)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "zksnap-circuits-halo2_b200")
LIB = os.environ.get("ZKB200_EMU_LIB") or os.path.join(PKG, "libzkb200_hostemu.so")  # override: the ASan / UBSan build (tools/asan_emulator.sh)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_lib = None


def build(force=False):
    srcs = [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))]
    newest = max(os.path.getmtime(s) for s in srcs)
    if os.environ.get("ZKB200_EMU_LIB"):
        return LIB
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["make", "-C", PKG, "-B", "libzkb200_hostemu.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(_u64p) if a is not None else None


def vec_op(field, op, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    o = np.zeros_like(a)
    lib().zkb_emu_vec_op(ctypes.c_int(0 if field == "fr" else 1), ctypes.c_int({"mul": 0, "add": 1, "sub": 2, "mul_lazy": 4, "add_lazy": 5, "sub_lazy": 6, "is_zero_lazy": 7, "mul_lazy_raw": 8}[op]),
                         _p(a), _p(b), _p(o), ctypes.c_size_t(a.shape[0]))
    return o


def ntt(a, log_n, omega, in_scale3=None, out_scale3=None, cols=1):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(cols, -1, 4)
    in_len = a.shape[1]
    out = np.zeros((cols, 1 << log_n, 4), dtype=np.uint64)
    ins = np.ascontiguousarray(in_scale3, dtype=np.uint64) if in_scale3 is not None else None
    outs = np.ascontiguousarray(out_scale3, dtype=np.uint64) if out_scale3 is not None else None
    rc = lib().zkb_emu_ntt(_p(a), ctypes.c_uint64(in_len), _p(out), ctypes.c_uint32(log_n),
                           _p(np.ascontiguousarray(omega, dtype=np.uint64)), _p(ins), _p(outs), ctypes.c_uint32(cols))
    assert rc == 0
    return out if cols > 1 else out[0]


def msm(scalars, bases, c=0, chunk=0, table=False):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(12, dtype=np.uint64)
    fn = lib().zkb_emu_msm_table if table else lib().zkb_emu_msm
    rc = fn(_p(s), _p(b), ctypes.c_uint64(s.shape[0]), ctypes.c_uint32(c), ctypes.c_uint32(chunk), _p(out))
    assert rc == 0
    return out


def msm_sliced(scalars, bases, slices, c=0, chunk=0):
    """table-mode MSM run as `slices` point-range slices adding into one bucket array (the pipelined host commit)"""
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(12, dtype=np.uint64)
    rc = lib().zkb_emu_msm_sliced(_p(s), _p(b), ctypes.c_uint64(s.shape[0]), ctypes.c_uint32(c), ctypes.c_uint32(chunk),
                                  ctypes.c_uint32(slices), _p(out))
    assert rc == 0
    return out


def msm_batch(scalar_cols, bases, c=0, chunk=0, table=False, dev_final=False):
    s = np.ascontiguousarray(scalar_cols, dtype=np.uint64)
    ncols, n = s.shape[0], s.shape[1]
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros((ncols, 12), dtype=np.uint64)
    rc = lib().zkb_emu_msm_batch(_p(s), _p(b), ctypes.c_uint64(n), ctypes.c_uint32(ncols), ctypes.c_uint32(c), ctypes.c_uint32(chunk),
                                 _p(out), ctypes.c_int(int(table)), ctypes.c_int(int(dev_final)))
    assert rc == 0
    return out


def fixed_base_mul(scalars):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((s.shape[0], 8), dtype=np.uint64)
    lib().zkb_emu_fixed_base_mul(_p(s), ctypes.c_uint64(s.shape[0]), _p(out))
    return out


def ntt_dist_phase(phase, rank, log_g, log_n, omega, A, W, O):
    """One phase of the sharded NTT for `rank`; A/W/O are lists (one per rank) of (N/G, 4) uint64 slices (any memory,
    e.g. numpy views of POSIX shared memory)."""
    G = 1 << log_g
    arr = lambda bufs: (_u64p * G)(*[b.ctypes.data_as(_u64p) for b in bufs])
    rc = lib().zkb_emu_ntt_dist_phase(ctypes.c_int(phase), ctypes.c_uint32(rank), ctypes.c_uint32(log_g), ctypes.c_uint32(log_n),
                                      _p(np.ascontiguousarray(omega, dtype=np.uint64)), arr(A), arr(W), arr(O))
    return rc


def ntt_dist(a, log_n, omega, log_g):
    """Whole sharded NTT in one process: every rank's pass 0, barrier, every rank's remaining passes."""
    G = 1 << log_g
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    sl = a.shape[0] // G
    A = [a[r * sl:(r + 1) * sl].copy() for r in range(G)]
    W = [np.zeros((sl, 4), dtype=np.uint64) for _ in range(G)]
    O = [np.zeros((sl, 4), dtype=np.uint64) for _ in range(G)]
    for phase in (0, 1):
        for r in range(G):
            rc = ntt_dist_phase(phase, r, log_g, log_n, omega, A, W, O)
            if rc != 0:
                return rc, None
    return 0, np.concatenate(O)


def fixed_base_window(scalars):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((s.shape[0], 8), dtype=np.uint64)
    lib().zkb_emu_fixed_base_window(_p(s), ctypes.c_uint64(s.shape[0]), _p(out), ctypes.c_uint32(0))
    return out


def batch_normalize(jac):
    j = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((j.shape[0], 8), dtype=np.uint64)
    lib().zkb_emu_batch_normalize(_p(j), ctypes.c_uint64(j.shape[0]), _p(out))
    return out


def setup_scalars(kind, n, s, omega=None, c=None):
    """kind 0: s^i; kind 1: Lagrange-basis scalars l_i(s) = omega^i c / (s - omega^i)"""
    out = np.zeros((n, 4), dtype=np.uint64)
    z = np.zeros(4, dtype=np.uint64)
    rc = lib().zkb_emu_setup_scalars(ctypes.c_int(kind), ctypes.c_uint64(n), _p(np.ascontiguousarray(s, dtype=np.uint64)),
                                     _p(np.ascontiguousarray(omega if omega is not None else z, dtype=np.uint64)),
                                     _p(np.ascontiguousarray(c if c is not None else z, dtype=np.uint64)), _p(out))
    return rc, out


def poly_eval(a, x):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    lib().zkb_emu_poly_eval(_p(a), ctypes.c_uint64(a.shape[0]), _p(np.ascontiguousarray(x, dtype=np.uint64)), _p(out))
    return out


def kate_division(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((a.shape[0] - 1, 4), dtype=np.uint64)
    lib().zkb_emu_kate_division(_p(a), ctypes.c_uint64(a.shape[0]), _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(out))
    return out


def batch_invert(a):
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    lib().zkb_emu_batch_invert(_p(a), ctypes.c_uint64(a.shape[0]))
    return a


def g1_fft(points_aff, log_n, omega, scale=None):
    p = np.ascontiguousarray(points_aff, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros_like(p)
    lib().zkb_emu_g1_fft(_p(p), ctypes.c_uint32(log_n), _p(np.ascontiguousarray(omega, dtype=np.uint64)),
                         _p(np.ascontiguousarray(scale, dtype=np.uint64)) if scale is not None else None, _p(out))
    return out


def prefix_product(a):
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    lib().zkb_emu_prefix_product(_p(a), ctypes.c_uint64(a.shape[0]))
    return a


def graph_evaluate(graph, fixed, advice, instance, challenges, beta, gamma, theta, y, rot_scale, values, cta_threads=0, halo=None):
    """graph: evaluation.GraphEvaluator; columns: lists of (isize, 4) host arrays (passed as addresses where the product takes
    handles).  Returns (rc, new values, info = [instructions, slots, polys])."""
    keep = [[np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4) for c in cols] for cols in (fixed, advice, instance)]
    out = np.array(values, dtype=np.uint64, copy=True).reshape(-1, 4)
    g, kg = graph.c_graph()
    inp, ki = graph.c_inputs(*[[a.ctypes.data for a in cols] for cols in keep], challenges, beta, gamma, theta, y, rot_scale)
    info = (ctypes.c_uint32 * 3)()
    lo, hi = halo if halo is not None else (0xFFFFFFFFFFFFFFFF, 0)   # halo = (halo_lo, halo_hi): row-window mode
    rc = lib().zkb_emu_graph_evaluate(ctypes.byref(g), ctypes.byref(inp), _p(out), ctypes.c_uint64(out.shape[0]),
                                      ctypes.c_uint32(cta_threads), info, ctypes.c_uint64(lo), ctypes.c_uint64(hi))
    del kg, ki
    return rc, out, list(info)


def permute_expression_pair(input_, table, usable_rows):
    a = np.ascontiguousarray(input_, dtype=np.uint64).reshape(-1, 4)
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
    oa, ot = np.zeros((usable_rows, 4), dtype=np.uint64), np.zeros((usable_rows, 4), dtype=np.uint64)
    lib().zkb_emu_permute_expression_pair.restype = ctypes.c_int
    rc = lib().zkb_emu_permute_expression_pair(_p(a), _p(t), ctypes.c_uint64(usable_rows), _p(oa), _p(ot))
    if rc != 0:
        raise ValueError("ConstraintSystemFailure")
    return oa, ot


def bucket_sort(keys, vals, key_bits, tile=0):
    """csrc/bucket_sort.cuh run phase by phase on the CPU: returns (sorted keys, permuted values), or None when the key is too wide."""
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    rc = lib().zkb_emu_bucket_sort(k.ctypes.data_as(u32p), v.ctypes.data_as(u32p), ctypes.c_uint64(k.size), ctypes.c_uint32(key_bits), ctypes.c_uint32(tile))
    return (k, v) if rc == 0 else None
