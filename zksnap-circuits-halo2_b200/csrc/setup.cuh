// setup.cuh — SRS generation and point normalisation, device code as per-thread functions.
//
// SURVEY.md §8(f) row 2 and §8(a) row a9.  Replaces (on the device)
//   ParamsKZG::<Bn256>::setup(k, rng)   /root/reference/voter/benches/voter_circuit.rs:60,
//                                       /root/reference/aggregator/benches/state_transition_circuit.rs:64
//       g[i] = [s^i] G,  g_lagrange[i] = [l_i(s)] G,  l_i(s) = omega^i (s^n - 1) / (n (s - omega^i))
//       (halo2-axiom poly/kzg/commitment.rs, un-vendored: restated from the published algorithm)
//   Curve::batch_normalize (halo2curves, Montgomery's trick)            Jacobian -> affine before transcript writes
//
//   fixed-base multiples of G: 16 unsigned 16-bit windows over a table T[w][d-1] = [d 2^(16 w)] G kept in HBM (64 MiB,
//   built once per process with the naive double-and-add kernel), i.e. <= 16 mixed additions per scalar, followed by a
//   batched conversion to affine (one inversion per FB_BATCH points per thread).
#pragma once
#include "curve.cuh"

namespace zkb {

constexpr uint32_t FB_WINDOW_BITS = 16;
constexpr uint32_t FB_WINDOWS = 16;                       // 16 x 16 = 256 >= 254 scalar bits
constexpr uint32_t FB_ROW = (1u << FB_WINDOW_BITS) - 1;   // entries per window (d = 1 .. 2^16 - 1)
constexpr int FB_BATCH = 16;                              // points per thread in the batched to-affine

// a^(m-2) (Fermat); a != 0
template <class P>
ZKB_HD_NOINLINE Fp<P> fp_inv(const Fp<P>& a) {
    Fp<P> acc = Fp<P>::one();
    for (int i = 255; i >= 0; --i) {
        acc = fp_sqr<P>(acc);
        uint32_t limb = P::M(i >> 5);
        if (i < 32) limb -= 2;  // both moduli end in ...01 / ...47: no borrow
        if ((limb >> (i & 31)) & 1) acc = fp_mul<P>(acc, a);
    }
    return acc;
}

// ---- scalars d * 2^(16 w) for the table build (Montgomery Fr, fed to the naive kernel) ---------------------------------
struct FbScalarArgs {
    uint4* out;  // FB_WINDOWS * FB_ROW Montgomery Fr
};
ZKB_HD void fb_table_scalar_thread(const FbScalarArgs& a, uint64_t t) {
    if (t >= (uint64_t)FB_WINDOWS * FB_ROW) return;
    const uint32_t w = (uint32_t)(t / FB_ROW), d = (uint32_t)(t % FB_ROW) + 1;
    Fr v = Fr::zero();
    const uint32_t bit = w * FB_WINDOW_BITS;  // multiple of 16: d sits in one 32-bit limb
    v.l[bit >> 5] = d << (bit & 31);
    fr_store2(a.out, t, fp_to_mont(v));       // top window: d 2^240 may exceed r — fp_to_mont reduces it, [x]G = [x mod r]G
}

// ---- windowed fixed-base multiplication: XYZZ result per scalar --------------------------------------------------------
struct FbMulArgs {
    const uint4* scalars;  // n Montgomery Fr
    uint64_t n;
    const uint4* table;    // FB_WINDOWS * FB_ROW affine points
    uint4* out_xyzz;       // n x 128 B
};
ZKB_HD void fb_mul_thread(const FbMulArgs& a, uint64_t i) {
    if (i >= a.n) return;
    const Fr s = fp_from_mont(fr_load2(a.scalars, i));
    XYZZ acc = XYZZ::identity();
    for (uint32_t w = 0; w < FB_WINDOWS; ++w) {
        const uint32_t d = (s.l[w >> 1] >> ((w & 1) * 16)) & 0xffffu;
        if (!d) continue;
        Affine p = affine_load(a.table + 4 * ((uint64_t)w * FB_ROW + (d - 1)));
        xyzz_add_mixed(acc, p.x, p.y);
    }
    acc.store(a.out_xyzz + 8 * i);
}

// ---- batched conversion to affine (Montgomery's trick inside one thread, strided elements for coalescing) -------------------
struct BatchAffineArgs {
    const uint4* in;     // n points: XYZZ (128 B) or Jacobian (96 B)
    uint64_t n;
    uint4* out;          // n affine points (64 B); identity -> (0, 0)
    uint32_t jacobian;   // 0: XYZZ input, 1: Jacobian input
    uint64_t nthreads;   // element j of thread t is t + j * nthreads
};
ZKB_HD bool batch_affine_denominator(const BatchAffineArgs& a, uint64_t i, Fq& d) {
    if (a.jacobian) {
        d = Fq::load(reinterpret_cast<const char*>(a.in) + 96 * i + 64);
        return !d.is_zero();
    }
    const char* p = reinterpret_cast<const char*>(a.in) + 128 * i;
    Fq zz = Fq::load(p + 64);
    if (zz.is_zero()) return false;
    d = fp_mul(zz, Fq::load(p + 96));
    return true;
}
ZKB_HD void g1_batch_to_affine_thread(const BatchAffineArgs& a, uint64_t t) {
    if (t >= a.nthreads || t >= a.n) return;
    Fq pre[FB_BATCH];
    Fq acc = Fq::one();
    for (int j = 0; j < FB_BATCH; ++j) {
        const uint64_t i = t + (uint64_t)j * a.nthreads;
        pre[j] = acc;
        Fq d;
        if (i < a.n && batch_affine_denominator(a, i, d)) acc = fp_mul(acc, d);
    }
    Fq inv = fp_inv(acc);
    for (int j = FB_BATCH - 1; j >= 0; --j) {
        const uint64_t i = t + (uint64_t)j * a.nthreads;
        if (i >= a.n) continue;
        Fq d;
        Fq x = Fq::zero(), y = Fq::zero();
        if (batch_affine_denominator(a, i, d)) {
            const Fq dinv = fp_mul(inv, pre[j]);
            inv = fp_mul(inv, d);
            if (a.jacobian) {  // (X / Z^2, Y / Z^3)
                const char* p = reinterpret_cast<const char*>(a.in) + 96 * i;
                const Fq zi2 = fp_sqr(dinv);
                x = fp_mul(Fq::load(p), zi2);
                y = fp_mul(Fq::load(p + 32), fp_mul(zi2, dinv));
            } else {           // dinv = 1 / (ZZ ZZZ):  x = X dinv ZZZ,  y = Y dinv ZZ
                const char* p = reinterpret_cast<const char*>(a.in) + 128 * i;
                x = fp_mul(Fq::load(p), fp_mul(dinv, Fq::load(p + 96)));
                y = fp_mul(Fq::load(p + 32), fp_mul(dinv, Fq::load(p + 64)));
            }
        }
        x.store(a.out + 4 * i);
        y.store(a.out + 4 * i + 2);
    }
}

// ---- SRS scalars -------------------------------------------------------------------------------------------------------------
constexpr uint32_t SETUP_CHUNK = 32;  // consecutive indices per thread

// out[i] = s^i
struct FrPowersArgs {
    uint4* out;
    uint64_t n;
    uint32_t s[8];  // Montgomery Fr
};
ZKB_HD void fr_powers_thread(const FrPowersArgs& a, uint64_t t) {
    const uint64_t lo = t * SETUP_CHUNK;
    if (lo >= a.n) return;
    const Fr s = fr_from_words(a.s);
    Fr p = fp_pow_u64(s, lo);
    for (uint64_t i = lo; i < lo + SETUP_CHUNK && i < a.n; ++i) {
        fr_store2(a.out, i, p);
        p = fp_mul(p, s);
    }
}

// out[i] = omega^i * c / (s - omega^i),  c = (s^n - 1) / n  — the Lagrange basis polynomials evaluated at s.
// status[0] is set when s lies in the domain (upstream panics on the failed inversion).
struct LagrangeScalarArgs {
    uint4* out;
    uint64_t n;
    uint32_t s[8], omega[8], c[8];  // Montgomery Fr
    uint32_t* status;
};
ZKB_HD void lagrange_scalars_thread(const LagrangeScalarArgs& a, uint64_t t) {
    const uint64_t lo = t * SETUP_CHUNK;
    if (lo >= a.n) return;
    const Fr s = fr_from_words(a.s), omega = fr_from_words(a.omega), c = fr_from_words(a.c);
    Fr w = fp_pow_u64(omega, lo);
    Fr pre[SETUP_CHUNK];
    Fr acc = Fr::one();
    uint32_t cnt = 0;
    Fr wi = w;
    for (uint64_t i = lo; i < lo + SETUP_CHUNK && i < a.n; ++i, ++cnt) {
        Fr d = fp_sub(s, wi);
        if (d.is_zero()) { a.status[0] = 1; d = Fr::one(); }
        pre[cnt] = acc;
        acc = fp_mul(acc, d);
        wi = fp_mul(wi, omega);
    }
    Fr inv = fp_inv(acc);
    // wi is now omega^(lo + cnt); walk back with omega^-1 = omega^(n-1)
    const Fr omega_inv = fp_pow_u64(omega, a.n - 1);
    for (uint32_t j = cnt; j-- > 0;) {
        wi = fp_mul(wi, omega_inv);  // omega^(lo + j)
        Fr d = fp_sub(s, wi);
        if (d.is_zero()) d = Fr::one();
        const Fr dinv = fp_mul(inv, pre[j]);
        inv = fp_mul(inv, d);
        fr_store2(a.out, lo + j, fp_mul(fp_mul(wi, c), dinv));
    }
}

// ---- points read from a file: canonical coordinates on y^2 = x^3 + 3 (or the identity (0, 0)), as SerdeFormat::RawBytes checks ---
struct OnCurveArgs {
    const uint4* pts;              // n affine points, 64 B each
    uint64_t n;
    unsigned long long* bad;       // number of points that fail
};
ZKB_HD bool g1_affine_is_valid(const Affine& p) {
    if (p.is_identity()) return true;
    if (!(fp_canon(p.x) == p.x) || !(fp_canon(p.y) == p.y)) return false;   // a limb pattern >= p is not a field element
    const Fq one = Fq::one();
    const Fq b = fp_add(fp_add(one, one), one);
    return fp_sqr(p.y) == fp_add(fp_mul(fp_sqr(p.x), p.x), b);
}

}  // namespace zkb

// ---- best_fft over G1 (halo2's FftGroup impl for curve points) — used by g_to_lagrange -------------------------------------------
// Radix-2 decimation in time on XYZZ points, one butterfly per thread and one launch per stage: the work is the scalar
// multiplication by the twiddle (~4300 Fq products per butterfly), so memory layout is irrelevant here.
namespace zkb {

// [k]P for a canonical 254-bit scalar k (8 limbs), double-and-add from the top bit
ZKB_HD_NOINLINE XYZZ xyzz_mul_scalar(const XYZZ& p, const Fr& k) {
    XYZZ r = XYZZ::identity();
    if (p.is_identity()) return r;
    for (int i = 253; i >= 0; --i) {
        r = xyzz_double(r);
        if ((k.l[i >> 5] >> (i & 31)) & 1) xyzz_add(r, p);
    }
    return r;
}
ZKB_HD XYZZ xyzz_neg(const XYZZ& p) {
    XYZZ r = p;
    if (!p.is_identity()) r.y = fp_neg(p.y);
    return r;
}

// a[i] <-> a[bitreverse(i)] (thread i swaps when i < rev)
struct G1BitrevArgs {
    uint4* a;  // n XYZZ
    uint32_t log_n;
};
ZKB_HD void g1_bitrev_thread(const G1BitrevArgs& p, uint64_t i) {
    const uint64_t n = 1ull << p.log_n;
    if (i >= n) return;
    uint64_t r = 0;
    for (uint32_t b = 0; b < p.log_n; ++b) r |= ((i >> b) & 1) << (p.log_n - 1 - b);
    if (i < r) {
        XYZZ x = XYZZ::load(p.a + 8 * i), y = XYZZ::load(p.a + 8 * r);
        y.store(p.a + 8 * i);
        x.store(p.a + 8 * r);
    }
}

// stage s (m = 2^s): butterfly t = (block k, offset j < m/2):  u = a[k m + j], v = w^(j n / m) a[k m + j + m/2]
struct G1FftStageArgs {
    uint4* a;             // n XYZZ
    const uint4* tw;      // omega^t as CANONICAL integers, t < n/2
    uint32_t log_n, stage;
};
ZKB_HD void g1_fft_stage_thread(const G1FftStageArgs& p, uint64_t t) {
    const uint64_t half_n = 1ull << (p.log_n - 1);
    if (t >= half_n) return;
    const uint32_t log_half_m = p.stage - 1;
    const uint64_t j = t & ((1ull << log_half_m) - 1), k = t >> log_half_m;
    const uint64_t i0 = (k << p.stage) + j, i1 = i0 + (1ull << log_half_m);
    XYZZ u = XYZZ::load(p.a + 8 * i0), v = XYZZ::load(p.a + 8 * i1);
    if (j) v = xyzz_mul_scalar(v, fr_load2(p.tw, j << (p.log_n - p.stage)));
    XYZZ s = u;
    xyzz_add(s, v);
    xyzz_add(u, xyzz_neg(v));
    s.store(p.a + 8 * i0);
    u.store(p.a + 8 * i1);
}

// tw[t] = omega^t as canonical integers (chunks of SETUP_CHUNK per thread)
struct FrPowCanonArgs {
    uint4* out;
    uint64_t n;
    uint32_t w[8];
};
ZKB_HD void fr_pow_canon_thread(const FrPowCanonArgs& a, uint64_t t) {
    const uint64_t lo = t * SETUP_CHUNK;
    if (lo >= a.n) return;
    const Fr w = fr_from_words(a.w);
    Fr p = fp_pow_u64(w, lo);
    for (uint64_t i = lo; i < lo + SETUP_CHUNK && i < a.n; ++i) {
        fr_store2(a.out, i, fp_from_mont(p));
        p = fp_mul(p, w);
    }
}

// a[i] <- [k] a[i] (the 1/n of an inverse transform), affine or XYZZ input -> XYZZ in place
struct G1ScaleArgs {
    uint4* a;       // n XYZZ
    uint64_t n;
    uint32_t k[8];  // canonical scalar
};
ZKB_HD void g1_scale_thread(const G1ScaleArgs& p, uint64_t i) {
    if (i >= p.n) return;
    xyzz_mul_scalar(XYZZ::load(p.a + 8 * i), fr_from_words(p.k)).store(p.a + 8 * i);
}

// affine (64 B) -> XYZZ (128 B)
struct G1LiftArgs {
    const uint4* in;
    uint4* out;
    uint64_t n;
};
ZKB_HD void g1_lift_thread(const G1LiftArgs& p, uint64_t i) {
    if (i >= p.n) return;
    Affine a = affine_load(p.in + 4 * i);
    XYZZ::from_affine(a).store(p.out + 8 * i);
}

}  // namespace zkb
