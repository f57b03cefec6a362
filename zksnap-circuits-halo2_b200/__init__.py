"""zkb200 — B200-native BN254 G1 MSM and Fr NTT behind halo2's call signatures.

Host-side mirror (Python over the C ABI in include/zkb200.h) of the halo2-axiom functions on the zksnap
provers' hot path.  The reference's toolchain (Rust) is absent from the build image, so this
mirror plays the role the Rust shim (shim/, source only — see INTEGRATION.md) plays in production: same
function names, argument meaning and panics-as-exceptions as

    halo2_proofs::arithmetic::{best_multiexp, best_fft}
    halo2_proofs::poly::EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff}
    halo2_proofs::poly::kzg::commitment::ParamsKZG::{commit, commit_lagrange}

reached in the reference from /root/reference/aggregator/src/wrapper.rs:106-137 (gen_pk, gen_proof) and the
three benches.  Arrays are numpy uint64 in halo2curves' memory layout (Montgomery limbs).
"""
from .halo2 import (  # noqa: F401
    EvaluationDomain,
    Polynomial,
    batch_invert,
    eval_polynomial,
    kate_division,
    ParamsKZG,
    batch_normalize,
    best_fft,
    best_fft_g1,
    best_multiexp,
    bound_devices,
    device_count,
    g1_fixed_base_mul,
    g1_fixed_base_mul_naive,
    g1_sum,
    init,
    launch_count,
    lib,
    omega,
    prof,
    shutdown,
)
from ._ffi import ZkbError, header_symbols  # noqa: F401
from .evaluation import GraphEvaluator, QuotientEvaluator, ValueSource  # noqa: F401

__all__ = [
    "GraphEvaluator", "QuotientEvaluator", "ValueSource", "EvaluationDomain", "ParamsKZG", "Polynomial", "batch_invert", "eval_polynomial", "kate_division", "batch_normalize", "best_fft", "best_fft_g1", "best_multiexp", "bound_devices", "device_count", "g1_fixed_base_mul", "g1_fixed_base_mul_naive", "g1_sum",
    "init", "launch_count", "lib", "omega", "prof", "shutdown", "ZkbError", "header_symbols",
]
