"""ctypes loader of libzkb200.so — binds exactly the symbols declared in include/zkb200.h.

There is no fallback: if the shared library is missing or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")  # override: A/B builds of the same ABI
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "zkb200.h")

u64p = ctypes.POINTER(ctypes.c_uint64)
u64pp = ctypes.POINTER(u64p)

_lib = None


class ZkbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libzkb200 error {code}: {msg}")
        self.code = code


def header_symbols() -> list[str]:
    """Every function name declared in include/zkb200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkbError(-4, f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(libzkb200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.zkb_last_error.restype = ctypes.c_char_p
    lib.zkb_version.restype = ctypes.c_char_p
    lib.zkb_launch_count.restype = ctypes.c_uint64
    sz, u32, u64, vp, ci = ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int
    sigs = {
        "zkb_init": [ctypes.POINTER(ci), ci],
        "zkb_msm_g1": [u64p, u64p, sz, u64p],
        "zkb_srs_register": [u64p, sz, u64p],
        "zkb_srs_release": [u64],
        "zkb_srs_load_file": [ctypes.c_char_p, u64, sz, ci, u64p],
        "zkb_msm_g1_srs": [u64, u64p, sz, u64p],
        "zkb_msm_g1_srs_range": [u64, sz, u64p, sz, u64p],
        "zkb_msm_g1_srs_batch": [u64, u64pp, sz, sz, u64p],
        "zkb_g1_sum": [u64p, sz, u64p],
        "zkb_srs_set_precompute": [ci],
        "zkb_srs_precompute": [u64, ctypes.POINTER(u32), ctypes.POINTER(u64)],
        "zkb_srs_table_info": [u64, ctypes.POINTER(u32), ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "zkb_g1_fixed_base_mul": [u64p, sz, u64p],
        "zkb_g1_fixed_base_mul_naive": [u64p, sz, u64p],
        "zkb_g1_batch_normalize": [u64p, sz, u64p],
        "zkb_kzg_setup": [u32, u64p, u64p, u64p],
        "zkb_kzg_setup_resident": [u32, u64p, ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "zkb_srs_download": [u64, u64p, sz],
        "zkb_g1_ntt": [u64p, u64p, u64p, u32],
        "zkb_srs_g_to_lagrange": [u64, u32, ctypes.POINTER(u64)],
        "zkb_ntt_fr": [u64p, u64p, u32],
        "zkb_ntt_fr_batch": [u64pp, sz, u64p, u32],
        "zkb_lagrange_to_coeff": [u64p, u32],
        "zkb_lagrange_to_coeff_batch": [u64pp, sz, u32],
        "zkb_coeff_to_lagrange": [u64p, u32],
        "zkb_coeff_to_extended": [u64p, u64p, u32, u32],
        "zkb_coeff_to_extended_batch": [u64pp, u64pp, sz, u32, u32],
        "zkb_extended_to_coeff": [u64p, u32, u32],
        "zkb_fr_omega": [u32, u64p],
        "zkb_fr_zeta": [u64p],
        "zkb_bound_devices": [ctypes.POINTER(ci), ci],
        "zkb_multi_device_set": [ci, ci, ci],
        "zkb_thread_bind_device": [ci],
        "zkb_dist_create_inprocess": [u32],
        "zkb_msm_g1_srs_dev": [u64, sz, vp, sz, u64p, vp],
        "zkb_ntt_fr_dev": [vp, vp, sz, u64p, u32, vp],
        "zkb_coeff_to_extended_dev": [vp, vp, vp, sz, u32, u32, vp],
        "zkb_extended_to_coeff_dev": [vp, vp, sz, u32, u32, vp],
        "zkb_lagrange_to_coeff_dev": [vp, vp, sz, u32, vp],
        "zkb_poly_upload": [u64p, sz, ctypes.POINTER(u64)],
        "zkb_poly_alloc": [sz, ctypes.POINTER(u64)],
        "zkb_poly_load_file": [ctypes.c_char_p, u64, sz, ctypes.POINTER(u64)],
        "zkb_transfer_stats": [ctypes.POINTER(u64), ci],
        "zkb_poly_len": [u64, ctypes.POINTER(sz)],
        "zkb_poly_write": [u64, sz, u64p, sz],
        "zkb_poly_download": [u64, u64p, sz],
        "zkb_poly_free": [u64],
        "zkb_poly_commit": [u64, u64, u64p],
        "zkb_poly_lagrange_to_coeff": [u64, u32],
        "zkb_poly_coeff_to_lagrange": [u64, u32],
        "zkb_poly_coeff_to_extended": [u64, u32, u32, ctypes.POINTER(u64)],
        "zkb_poly_extended_to_coeff": [u64, u32, u32],
        "zkb_poly_eval": [u64, u64p, u64p],
        "zkb_poly_kate_division": [u64, u64p, ctypes.POINTER(u64)],
        "zkb_poly_batch_invert": [u64],
        "zkb_poly_mul": [u64, u64],
        "zkb_poly_mul_periodic": [u64, u64p, u32],
        "zkb_poly_slice": [u64, sz, sz, u64p],
        "zkb_poly_scale_add": [u64, u64p, u64],
        "zkb_poly_add_const": [u64, u64p],
        "zkb_poly_prefix_product": [u64],
        "zkb_lookup_permute_expression_pair": [u64, u64, sz, ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "zkb_fr_eval_polynomial": [u64p, sz, u64p, u64p],
        "zkb_fr_kate_division": [u64p, sz, u64p, u64p],
        "zkb_fr_batch_invert": [u64p, sz],
        "zkb_graph_evaluate": [vp, vp, u64],
        "zkb_graph_evaluate_dev": [vp, vp, vp, sz, ci, sz, sz, vp],
        "zkb_graph_last_info": [ctypes.POINTER(u32)] * 4,
        "zkb_dist_create": [ci, ci, u32, ctypes.POINTER(ctypes.c_uint8)],
        "zkb_dist_connect": [ctypes.POINTER(ctypes.c_uint8)],
        "zkb_dist_destroy": [],
        "zkb_dist_ntt_fr": [u64p, u64p, u64p, u32],
        "zkb_dist_ntt_fr_dev": [vp, vp, u64p, u32, vp],
        "zkb_dist_buffers": [ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(sz)],
        "zkb_dist_status": [vp],
        "zkb_dist_set_timeout_ms": [u64],
        "zkb_host_register": [vp, sz],
        "zkb_host_unregister": [vp],
        "zkb_pipeline_set": [ci, sz],
        "zkb_msm_set_slices": [ci],
        "zkb_field_vec_op": [ci, ci, u64p, u64p, u64p, sz],
        "zkb_bucket_sort_pairs": [ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), sz, u32, u32],
        "zkb_msm_set_params": [u32, u32],
        "zkb_msm_get_params": [sz, ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u32)],
        "zkb_msm_last_entries": [ctypes.POINTER(u64)],
        "zkb_prof_enable": [ci],
        "zkb_prof_reset": [],
        "zkb_measure_imad_peak": [ctypes.POINTER(ctypes.c_double)],
        "zkb_prof_get": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ci
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise ZkbError(rc, load().zkb_last_error().decode(errors="replace"))
