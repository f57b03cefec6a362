// field.cuh — BN254 Fr / Fq Montgomery arithmetic for sm_100a, 8 x 32-bit limbs.
//
// Replaces (on the device) halo2curves::bn256::{Fr, Fq} mul/square/add/sub/neg — SURVEY.md §8 row a8.
// Memory layout is exactly halo2curves': 4 little-endian u64 limbs = 8 little-endian u32 limbs holding
// a * 2^256 mod m, always canonical (< m).  Every routine here returns canonical values so results are
// limb-for-limb identical with the CPU path.
//
// Multiplication is an operand-scanning Montgomery product on two interleaved accumulators: products
// a[j]*b[i] with even j land on limb pairs (0,1),(2,3).. of accumulator E, odd j on pairs (1,2),(3,4)..
// held in accumulator O (O[k] sits at limb position k+1).  Each 32x32->64 product therefore adds into an
// aligned 64-bit pair and a whole row is one uninterrupted carry chain of mad.lo.cc / madc.hi.cc pairs,
// which ptxas fuses into IMAD.WIDE.U32(.X) — 16 wide MACs per row pair (a*b[i] and m*N), 128 per product.
// After each row's reduction E[0] == 0, the value is shifted one limb by renaming (E' = O, O'[k] = E[k+2]),
// so no data moves.  All partial accumulators are sums of non-negative terms of a total < 2^288, hence the
// O-chain never carries out and the E-chain's carry is absorbed by O[7].
#pragma once
#include <cstdint>
#if !defined(__CUDA_ARCH__)
#include <cassert>
#endif

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_HD_NOINLINE __host__ __device__ inline
#else
#define ZKB_HD inline
#define ZKB_HD_NOINLINE inline
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace zkb {

struct FrParams {
    __host__ __device__ static constexpr uint32_t M(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    __host__ __device__ static constexpr uint32_t R(int i) {  // 2^256 mod r  (Montgomery one)
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    __host__ __device__ static constexpr uint32_t R2(int i) {  // 2^512 mod r
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqParams {
    __host__ __device__ static constexpr uint32_t M(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    __host__ __device__ static constexpr uint32_t R(int i) {
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    __host__ __device__ static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
};

// ---- carry-chain rows (each asm statement is self-contained w.r.t. the carry flag) -------------------

// O[0..7] += {x1,x3,x5,x7} * y on limb pairs (O0,O1)..(O6,O7).  No carry out (see header).
ZKB_HD void row_odd(uint32_t (&O)[8], uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7,
                                        uint32_t y) {
#if !defined(__CUDA_ARCH__)
    const uint32_t x[4] = {x1, x3, x5, x7};
    uint64_t c = 0;
    for (int k = 0; k < 4; ++k) {
        uint64_t p = (uint64_t)x[k] * y, t;
        t = (uint64_t)O[2 * k] + (uint32_t)p + c; O[2 * k] = (uint32_t)t; c = t >> 32;
        t = (uint64_t)O[2 * k + 1] + (p >> 32) + c; O[2 * k + 1] = (uint32_t)t; c = t >> 32;
    }
    assert(c == 0);
#else
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;\n\t"
        : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
#endif
}

// e0 += carry_limb, carry into the O chain; then as row_odd.  (The deferred one-limb shift: the old E[1]
// belongs at position 0 of the new even accumulator.)
ZKB_HD void row_odd_merge(uint32_t& e0, uint32_t carry_limb, uint32_t (&O)[8], uint32_t x1,
                                              uint32_t x3, uint32_t x5, uint32_t x7, uint32_t y) {
#if !defined(__CUDA_ARCH__)
    const uint32_t x[4] = {x1, x3, x5, x7};
    uint64_t t0 = (uint64_t)e0 + carry_limb;
    e0 = (uint32_t)t0;
    uint64_t c = t0 >> 32;
    for (int k = 0; k < 4; ++k) {
        uint64_t p = (uint64_t)x[k] * y, t;
        t = (uint64_t)O[2 * k] + (uint32_t)p + c; O[2 * k] = (uint32_t)t; c = t >> 32;
        t = (uint64_t)O[2 * k + 1] + (p >> 32) + c; O[2 * k + 1] = (uint32_t)t; c = t >> 32;
    }
    assert(c == 0);
#else
    asm("add.cc.u32     %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32    %7, %13, %14, %7;\n\t"
        : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7]), "+r"(e0)
        : "r"(carry_limb), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
#endif
}

// E[0..7] += {x0,x2,x4,x6} * y on limb pairs (E0,E1)..(E6,E7); the carry out (position 8) goes to o7.
ZKB_HD void row_even(uint32_t (&E)[8], uint32_t& o7, uint32_t x0, uint32_t x2, uint32_t x4,
                                         uint32_t x6, uint32_t y) {
#if !defined(__CUDA_ARCH__)
    const uint32_t x[4] = {x0, x2, x4, x6};
    uint64_t c = 0;
    for (int k = 0; k < 4; ++k) {
        uint64_t p = (uint64_t)x[k] * y, t;
        t = (uint64_t)E[2 * k] + (uint32_t)p + c; E[2 * k] = (uint32_t)t; c = t >> 32;
        t = (uint64_t)E[2 * k + 1] + (p >> 32) + c; E[2 * k + 1] = (uint32_t)t; c = t >> 32;
    }
    uint64_t t7 = (uint64_t)o7 + c;
    assert((t7 >> 32) == 0);
    o7 = (uint32_t)t7;
#else
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;\n\t"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]), "+r"(o7)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
#endif
}

#if !defined(__CUDA_ARCH__)
// portable 8-limb chains used by the host build (logic twin of the PTX chains)
inline uint32_t host_add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) { uint64_t t = (uint64_t)a[i] + b[i] + c; r[i] = (uint32_t)t; c = t >> 32; }
    return (uint32_t)c;
}
inline uint32_t host_sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // returns 0xffffffff on borrow
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) { uint64_t t = (uint64_t)a[i] - b[i] - br; r[i] = (uint32_t)t; br = (t >> 32) & 1; }
    return br ? 0xffffffffu : 0u;
}
#endif

template <class P>
struct Fp {
    uint32_t l[8];

    ZKB_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = 0;
        return r;
    }
    ZKB_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = P::R(i);
        return r;
    }
    ZKB_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = P::R2(i);
        return r;
    }
    ZKB_HD bool is_zero() const {
        return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0;
    }
    ZKB_HD bool operator==(const Fp& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) d |= l[i] ^ o.l[i];
        return d == 0;
    }

    // 128-bit vector load/store (pointer must be 16-byte aligned: all device buffers are)
    ZKB_HD static Fp load(const void* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = q[0], b = q[1];
        Fp r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    ZKB_HD static Fp load_nc(const void* p) {  // read-only path
        const uint4* q = reinterpret_cast<const uint4*>(p);
        #if defined(__CUDA_ARCH__)
        uint4 a = __ldg(q), b = __ldg(q + 1);
#else
        uint4 a = q[0], b = q[1];
#endif
        Fp r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    ZKB_HD void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(l[0], l[1], l[2], l[3]);
        q[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
};

// limb i of K * M (K = 1 or 2; 2M < 2^255 still fits eight limbs)
template <class P, int K>
__host__ __device__ constexpr uint32_t mod_k(int i) {
    return K == 1 ? P::M(i) : ((P::M(i) << 1) | (i ? P::M(i - 1) >> 31 : 0u));
}

// r = (t >= K M) ? t - K M : t      (t < 2 K M)
template <class P, int K = 1>
ZKB_HD void reduce_once(uint32_t (&t)[8]) {
    uint32_t s[8], borrow;
#if !defined(__CUDA_ARCH__)
    uint32_t mm[8];
    for (int i = 0; i < 8; ++i) mm[i] = mod_k<P, K>(i);
    borrow = host_sub8(s, t, mm);
#else
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;\n\t"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
          "=r"(borrow)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(mod_k<P, K>(0)),
          "r"(mod_k<P, K>(1)), "r"(mod_k<P, K>(2)), "r"(mod_k<P, K>(3)), "r"(mod_k<P, K>(4)), "r"(mod_k<P, K>(5)), "r"(mod_k<P, K>(6)), "r"(mod_k<P, K>(7)));
#endif
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = borrow ? t[i] : s[i];
}

// K = 1: canonical in, canonical out.  K = 2 ("lazy"): inputs in [0, 2M), output in [0, 2M), congruent mod M.
template <class P, int K>
ZKB_HD Fp<P> fp_add_k(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[8];
#if !defined(__CUDA_ARCH__)
    host_add8(t, a.l, b.l);
#else
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;\n\t"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#endif
    reduce_once<P, K>(t);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
    return r;
}

template <class P>
ZKB_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) { return fp_add_k<P, 1>(a, b); }
template <class P>
ZKB_HD Fp<P> fp_add_lazy(const Fp<P>& a, const Fp<P>& b) { return fp_add_k<P, 2>(a, b); }

template <class P, int K>
ZKB_HD Fp<P> fp_sub_k(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[8], borrow;
    Fp<P> r;
#if !defined(__CUDA_ARCH__)
    borrow = host_sub8(t, a.l, b.l);
    uint32_t mm[8];
    for (int i = 0; i < 8; ++i) mm[i] = mod_k<P, K>(i) & borrow;
    host_add8(r.l, t, mm);
#else
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;\n\t"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
          "=r"(borrow)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // conditional + M as a predicated carry chain (no masking instructions)
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %8, 0;\n\t"
        "@p add.cc.u32  %0, %0, %9;\n\t"
        "@p addc.cc.u32 %1, %1, %10;\n\t"
        "@p addc.cc.u32 %2, %2, %11;\n\t"
        "@p addc.cc.u32 %3, %3, %12;\n\t"
        "@p addc.cc.u32 %4, %4, %13;\n\t"
        "@p addc.cc.u32 %5, %5, %14;\n\t"
        "@p addc.cc.u32 %6, %6, %15;\n\t"
        "@p addc.u32    %7, %7, %16;\n\t"
        "}\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
        : "r"(borrow), "r"(mod_k<P, K>(0)), "r"(mod_k<P, K>(1)), "r"(mod_k<P, K>(2)), "r"(mod_k<P, K>(3)), "r"(mod_k<P, K>(4)), "r"(mod_k<P, K>(5)), "r"(mod_k<P, K>(6)),
          "r"(mod_k<P, K>(7)));
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
#endif
    return r;
}

template <class P>
ZKB_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) { return fp_sub_k<P, 1>(a, b); }
template <class P>
ZKB_HD Fp<P> fp_sub_lazy(const Fp<P>& a, const Fp<P>& b) { return fp_sub_k<P, 2>(a, b); }

template <class P>
ZKB_HD Fp<P> fp_neg(const Fp<P>& a) {
    return fp_sub<P>(Fp<P>::zero(), a);
}

template <class P>
ZKB_HD Fp<P> fp_dbl(const Fp<P>& a) {
    return fp_add<P>(a, a);
}

// Montgomery product a*b*2^-256 mod M.  REDUCE: canonical result (inputs may be lazy, i.e. < 2M: 4M^2/2^256 + M < 2M since
// 4M < 2^256).  !REDUCE ("lazy"): the final conditional subtraction is skipped and the result is only < 2M — valid input for
// the next lazy product, lazy add/sub (mod 2M) and for the canonical routines' products.
template <class P, bool REDUCE>
ZKB_HD Fp<P> fp_mul_k(const Fp<P>& a, const Fp<P>& b) {
    uint32_t E[8], O[8];
    // row 0: plain 32x32->64 products into aligned pairs
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint64_t pe = (uint64_t)a.l[2 * k] * b.l[0];
        uint64_t po = (uint64_t)a.l[2 * k + 1] * b.l[0];
        E[2 * k] = (uint32_t)pe; E[2 * k + 1] = (uint32_t)(pe >> 32);
        O[2 * k] = (uint32_t)po; O[2 * k + 1] = (uint32_t)(po >> 32);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i > 0) {
            // deferred shift: E' = O (with old E[1] folded into position 0), O'[k] = E[k+2], top two limbs fresh
            uint32_t nE[8], nO[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nE[k] = O[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) nO[k] = E[k + 2];
            nO[6] = 0; nO[7] = 0;
            row_odd_merge(nE[0], E[1], nO, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            row_even(nE, nO[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) { E[k] = nE[k]; O[k] = nO[k]; }
        }
        uint32_t m = E[0] * P::INV;
        row_odd(O, P::M(1), P::M(3), P::M(5), P::M(7), m);
        row_even(E, O[7], P::M(0), P::M(2), P::M(4), P::M(6), m);
        // now E[0] == 0
    }
    // final one-limb shift and merge: t[k] = O[k] + E[k+1]
    uint32_t t[8];
#if !defined(__CUDA_ARCH__)
    {
        uint32_t sh[8] = {E[1], E[2], E[3], E[4], E[5], E[6], E[7], 0};
        uint32_t c = host_add8(t, O, sh);
        assert(c == 0 && E[0] == 0);
    }
#else
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, 0;\n\t"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(E[1]),
          "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]));
#endif
    if (REDUCE) reduce_once<P>(t);
    Fp<P> r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = t[k];
    return r;
}

template <class P>
ZKB_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) { return fp_mul_k<P, true>(a, b); }
template <class P>
ZKB_HD Fp<P> fp_mul_lazy(const Fp<P>& a, const Fp<P>& b) { return fp_mul_k<P, false>(a, b); }
// lazy value (< 2M) -> canonical
template <class P>
ZKB_HD Fp<P> fp_canon(const Fp<P>& a) {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = a.l[i];
    reduce_once<P>(t);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
    return r;
}
// a == 0 mod M for a lazy value (< 2M): a is 0 or M
template <class P>
ZKB_HD bool fp_is_zero_lazy(const Fp<P>& a) {
    uint32_t z = 0, m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { z |= a.l[i]; m |= a.l[i] ^ P::M(i); }
    return z == 0 || m == 0;
}

template <class P>
ZKB_HD Fp<P> fp_sqr(const Fp<P>& a) {
    return fp_mul<P>(a, a);
}

// canonical integer limbs from Montgomery form (multiply by 1) — halo2curves' to_repr()
template <class P>
ZKB_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> one = Fp<P>::zero();
    one.l[0] = 1;
    return fp_mul<P>(a, one);
}
template <class P>
ZKB_HD Fp<P> fp_to_mont(const Fp<P>& a) {
    return fp_mul<P>(a, Fp<P>::r2());
}

// a^e for a small public exponent (square-and-multiply, e > 0 scanned from the top bit)
template <class P>
ZKB_HD_NOINLINE Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
    Fp<P> acc = Fp<P>::one();
    if (e == 0) return acc;
    int top = 63;
    while (!((e >> top) & 1)) --top;
    for (int i = top; i >= 0; --i) {
        acc = fp_sqr<P>(acc);
        if ((e >> i) & 1) acc = fp_mul<P>(acc, a);
    }
    return acc;
}

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

ZKB_HD Fr fr_from_words(const uint32_t (&w)[8]) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = w[i];
    return r;
}

// element idx of an array of 32-byte field elements addressed as uint4 pairs
ZKB_HD Fr fr_from_u4(uint4 a, uint4 b) {
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
ZKB_HD Fr fr_load2(const uint4* base, uint64_t idx) {  // AoS global layout: element idx = uint4[2*idx], [2*idx+1]
    return fr_from_u4(base[2 * idx], base[2 * idx + 1]);
}
ZKB_HD void fr_store2(uint4* base, uint64_t idx, const Fr& v) {
    base[2 * idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    base[2 * idx + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

}  // namespace zkb
