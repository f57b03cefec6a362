// poly.cuh — Fr vector kernels either side of the commits, device code as per-thread functions.
//
// Restates (on the device) three halo2-axiom `arithmetic` helpers the prover calls on every committed polynomial
// (sources un-vendored; the maths fixes the results uniquely, the field being exact):
//   eval_polynomial(poly, x)          sum_i a_i x^i                       — evaluations written to the transcript
//   kate_division(a, b)               quotient of a(X) / (X - b)           — GWC / SHPLONK witness polynomials
//   BatchInvert (ff::BatchInvert)     element-wise inverse, zeros stay 0  — permutation / lookup grand products
// All three are chunked: a thread owns POLY_CHUNK consecutive coefficients; chunk results are combined by applying the
// same routine one level up (x -> x^CHUNK), so a 2^22 polynomial takes four short launches.
#pragma once
#include "setup.cuh"

namespace zkb {

constexpr uint32_t POLY_CHUNK = 64;

struct PolyEvalArgs {
    const uint4* a;   // n coefficients
    uint64_t n;
    uint4* out;       // ceil(n / POLY_CHUNK) partial values: out[t] = sum_j a[t C + j] x^j
    uint32_t x[8];    // Montgomery Fr
};
ZKB_HD void poly_eval_chunk_thread(const PolyEvalArgs& p, uint64_t t) {
    const uint64_t lo = t * POLY_CHUNK;
    if (lo >= p.n) return;
    const uint64_t hi = lo + POLY_CHUNK < p.n ? lo + POLY_CHUNK : p.n;
    const Fr x = fr_from_words(p.x);
    Fr acc = fr_load2(p.a, hi - 1);
    for (uint64_t i = hi - 1; i-- > lo;) acc = fp_add(fp_mul(acc, x), fr_load2(p.a, i));
    fr_store2(p.out, t, acc);
}

// Q_{i-1} = a_i + b Q_i inside chunk t, started from the carry K_t = Q at the chunk's last index
// (K of the last chunk is 0): writes Q_i for every index of the chunk.
struct KateExpandArgs {
    const uint4* a;      // n coefficients
    uint64_t n;
    const uint4* carry;  // ceil(n / POLY_CHUNK) values K_t, or NULL when there is a single chunk (K = 0)
    uint4* q;            // n values Q_0 .. Q_{n-1} (Q_{n-1} = 0); the quotient is q[0 .. n-2]
    uint32_t b[8];
};
ZKB_HD void kate_expand_thread(const KateExpandArgs& p, uint64_t t) {
    const uint64_t lo = t * POLY_CHUNK;
    if (lo >= p.n) return;
    const uint64_t hi = lo + POLY_CHUNK < p.n ? lo + POLY_CHUNK : p.n;
    const Fr b = fr_from_words(p.b);
    Fr q = p.carry ? fr_load2(p.carry, t) : Fr::zero();
    fr_store2(p.q, hi - 1, q);
    for (uint64_t i = hi - 1; i > lo; --i) {
        q = fp_add(fr_load2(p.a, i), fp_mul(b, q));
        fr_store2(p.q, i - 1, q);
    }
}

// in place: a[i] <- a[i]^-1, zeros untouched; one Fermat inversion per POLY_CHUNK / 2 elements
constexpr uint32_t INV_CHUNK = 32;
struct BatchInvertArgs {
    uint4* a;
    uint64_t n;
};
ZKB_HD void fr_batch_invert_thread(const BatchInvertArgs& p, uint64_t t) {
    const uint64_t lo = t * INV_CHUNK;
    if (lo >= p.n) return;
    const uint64_t hi = lo + INV_CHUNK < p.n ? lo + INV_CHUNK : p.n;
    Fr pre[INV_CHUNK];
    Fr acc = Fr::one();
    for (uint64_t i = lo; i < hi; ++i) {
        pre[i - lo] = acc;
        const Fr v = fr_load2(p.a, i);
        if (!v.is_zero()) acc = fp_mul(acc, v);
    }
    Fr inv = fp_inv(acc);
    for (uint64_t i = hi; i-- > lo;) {
        const Fr v = fr_load2(p.a, i);
        if (v.is_zero()) continue;
        fr_store2(p.a, i, fp_mul(inv, pre[i - lo]));
        inv = fp_mul(inv, v);
    }
}

// ---- element-wise product and prefix products: the grand products of the permutation / lookup arguments -----------------------
// z[0] = 1, z[i] = prod_{j < i} v[j]  (halo2's `iter::once(one).chain(values).scan(one, |s, v| { *s *= v; Some(*s) })`),
// as a chunked scan: chunk products, the same scan one level up on the chunk products, then the in-chunk expansion.
struct PolyMulArgs {
    uint4* a;         // a[i] <- a[i] * b[i]
    const uint4* b;
    uint64_t n;
};
ZKB_HD void poly_mul_thread(const PolyMulArgs& p, uint64_t i) {
    if (i >= p.n) return;
    fr_store2(p.a, i, fp_mul(fr_load2(p.a, i), fr_load2(p.b, i)));
}

// a[i] <- a[i] * t[i mod period], period a power of two <= 16: EvaluationDomain::divide_by_vanishing_poly — 1 / (X^n - 1) takes
// only 2^(extended_k - k) distinct values on the extended coset, so the table travels in the kernel arguments
constexpr uint32_t POLY_PERIOD_MAX = 16;
struct PolyMulPeriodicArgs {
    uint4* a;
    uint64_t n;
    uint32_t period;
    uint32_t t[POLY_PERIOD_MAX][8];
};
ZKB_HD void poly_mul_periodic_thread(const PolyMulPeriodicArgs& p, uint64_t i) {
    if (i >= p.n) return;
    fr_store2(p.a, i, fp_mul(fr_load2(p.a, i), fr_from_words(p.t[i & (p.period - 1)])));
}

// acc[i] <- acc[i] * k + other[i]   (Horner over polynomials: the multiopen provers fold their queries with powers of v)
struct PolyScaleAddArgs {
    uint4* acc;
    const uint4* other;   // may be NULL: acc[i] <- acc[i] * k
    uint64_t n;
    uint32_t k[8];
};
ZKB_HD void poly_scale_add_thread(const PolyScaleAddArgs& p, uint64_t i) {
    if (i >= p.n) return;
    Fr v = fp_mul_lazy(fr_load2(p.acc, i), fr_from_words(p.k));
    if (p.other) v = fp_add_lazy(v, fr_load2(p.other, i));
    fr_store2(p.acc, i, fp_canon(v));
}

struct ScanChunkArgs {
    const uint4* v;   // n values
    uint64_t n;
    uint4* prod;      // ceil(n / POLY_CHUNK) chunk products
};
ZKB_HD void scan_chunk_product_thread(const ScanChunkArgs& p, uint64_t t) {
    const uint64_t lo = t * POLY_CHUNK;
    if (lo >= p.n) return;
    const uint64_t hi = lo + POLY_CHUNK < p.n ? lo + POLY_CHUNK : p.n;
    Fr acc = fr_load2(p.v, lo);
    for (uint64_t i = lo + 1; i < hi; ++i) acc = fp_mul_lazy(acc, fr_load2(p.v, i));
    fr_store2(p.prod, t, fp_canon(acc));
}

struct ScanExpandArgs {
    const uint4* v;       // n values
    uint64_t n;
    const uint4* carry;   // exclusive prefix products of the chunk products (carry[t] = product of all chunks before t), or NULL
    uint4* z;             // n outputs: z[i] = prod_{j < i} v[j]
};
ZKB_HD void scan_expand_thread(const ScanExpandArgs& p, uint64_t t) {
    const uint64_t lo = t * POLY_CHUNK;
    if (lo >= p.n) return;
    const uint64_t hi = lo + POLY_CHUNK < p.n ? lo + POLY_CHUNK : p.n;
    Fr acc = p.carry ? fr_load2(p.carry, t) : Fr::one();
    for (uint64_t i = lo; i < hi; ++i) {
        const Fr v = fr_load2(p.v, i);  // read before the store: z may alias v
        fr_store2(p.z, i, acc);
        acc = fp_mul(acc, v);
    }
}

}  // namespace zkb
