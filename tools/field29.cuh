// field29.cuh — BN254 Fq / Fr arithmetic on 9 x 29-bit unsaturated limbs (carry-free accumulation).
//
// Why: on sm_100a the carry-chained IMAD.WIDE.U32.X that a 32-bit-limb Montgomery product compiles to occupies
// the FMA-heavy pipe about twice as long as a plain IMAD.WIDE.U32 (ncu, profiles/r1: fmaheavy 88% busy at 42% of
// the plain-IMAD.WIDE peak).  With 29-bit limbs a 64-bit column accumulator can absorb all 18 partial products
// of a Montgomery step (9 * 2^60 + 9 * 2^58 < 2^64) without any carry, so the whole product is plain
// IMAD.WIDE.U32 plus a handful of shifts on the otherwise idle ALU pipe.
//
// Representation ("loose"): value = sum l[i] * 2^(29 i), limbs < 2^29 (normalised) or < 2^30 (one pending lazy add),
// value < 2^261 but NOT necessarily < p.  The Montgomery radix is R' = 2^261 (~169.28 p): mont(a, b) = a*b/R' mod p
// with result < a*b/R' + p, i.e. < 2p whenever a*b < 169 p^2.  Memory (halo2curves) form x~ = x*2^256 mod p is
// entered by an integer shift x^ = x~ << 5 (= x*2^261 mod p, < 32 p, no reduction needed) and left by an exact
// 5-bit Montgomery step plus canonical reduction (from29), so results are bit-identical with the CPU path.
#pragma once
#include "field.cuh"

namespace zkb {

constexpr uint32_t MASK29 = (1u << 29) - 1;

struct Fq29Params {
    __host__ __device__ static constexpr uint32_t N(int i) {
        constexpr uint32_t v[9] = {0x187cfd47u, 0x010460b6u, 0x1c72a34fu, 0x02d522d0u, 0x1585d978u,
                                   0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
        return v[i];
    }
    static constexpr uint32_t INV = 0x04866389u;  // -p^-1 mod 2^29
    using P32 = FqParams;
};
struct Fr29Params {
    __host__ __device__ static constexpr uint32_t N(int i) {
        constexpr uint32_t v[9] = {0x10000001u, 0x1f0fac9fu, 0x0e5c2450u, 0x07d090f3u, 0x1585d283u,
                                   0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
        return v[i];
    }
    static constexpr uint32_t INV = 0x0fffffffu;  // -r^-1 mod 2^29
    using P32 = FrParams;
};

template <class P>
struct F29 {
    uint32_t l[9];
};

// acc += a * b as ONE plain IMAD.WIDE.U32 (written in PTX so the 64-bit multiply is never widened to 64x64)
ZKB_HD void madw(uint64_t& acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
    acc += (uint64_t)a * b;
#endif
}
ZKB_HD uint64_t mulw(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (uint64_t)a * b;
#endif
}

// memory form (8 x 32, Montgomery R = 2^256, canonical) -> loose 9 x 29 form with R' = 2^261: x^ = x~ << 5
template <class P>
ZKB_HD F29<P> to29(const Fp<typename P::P32>& a) {
    F29<P> r;
    // bits [29 i - 5, 29 i + 24) of a  (bit positions of a << 5)
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        int bit = 29 * i - 5;
        uint32_t v;
        if (bit < 0) {
            v = (a.l[0] << 5) & MASK29;
        } else {
            int w = bit >> 5, sh = bit & 31;
            uint32_t lo = a.l[w];
            uint32_t hi = (w + 1 < 8) ? a.l[w + 1] : 0u;
            v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
            v &= MASK29;
        }
        r.l[i] = v;
    }
    return r;
}

// limbs < 2^29 + small after one parallel carry pass (value unchanged; top limb keeps its excess)
template <class P>
ZKB_HD F29<P> carry29(const F29<P>& a) {
    F29<P> r;
    r.l[0] = a.l[0] & MASK29;
#pragma unroll
    for (int i = 1; i < 8; ++i) r.l[i] = (a.l[i] & MASK29) + (a.l[i - 1] >> 29);
    r.l[8] = a.l[8] + (a.l[7] >> 29);
    return r;
}

template <class P>
ZKB_HD F29<P> add29(const F29<P>& a, const F29<P>& b) {  // lazy: limbs add, no carry
    F29<P> r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + b.l[i];
    return r;
}

// a - b + K*p for normalised b < K*p, with K*p spelled so every limb dominates a 29-bit limb of b:
// limbs of K*p are re-borrowed: l_i + 2^30 - 2 (i = 0), l_i + 2^30 - 2 (middle), top limb minus the borrowed 2.
template <class P, int K>
struct KP29 {
    // limb i of K*p in plain 29-bit spelling (the top limb keeps everything above bit 232)
    __host__ __device__ static constexpr uint32_t raw(int i) {
        uint64_t carry = 0, v = 0;
        for (int j = 0; j <= i; ++j) {
            v = (uint64_t)P::N(j) * K + carry;
            carry = v >> 29;
            if (j < 8) v &= MASK29;
        }
        return (uint32_t)v;
    }
    // borrowed spelling of the same integer: every limb below the top gets +2^30 and takes 2 from the limb above,
    // so limb(i) >= 2^30 - 2 dominates any normalised limb of the subtrahend
    __host__ __device__ static constexpr uint32_t limb(int i) {
        if (i == 0) return raw(0) + (1u << 30);
        if (i < 8) return raw(i) + (1u << 30) - 2;
        return raw(8) - 2;
    }
};

template <class P, int K>
ZKB_HD F29<P> sub29(const F29<P>& a, const F29<P>& b) {  // a + K*p - b ; b normalised (carry29 / mul output), b < (K-1)*p
    F29<P> r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + KP29<P, K>::limb(i) - b.l[i];
    return r;
}

// Montgomery product a*b/2^261 mod p.  Limbs of a and b < 2^30; result normalised (limbs < 2^29 except the
// top one), value < a*b/2^261 + p.
template <class P>
ZKB_HD F29<P> mul29(const F29<P>& a, const F29<P>& b) {
    uint64_t acc[18];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (i == 0 || j == 8) acc[i + j] = mulw(a.l[j], b.l[i]);   // first touch of this column
            else madw(acc[i + j], a.l[j], b.l[i]);
        }
        uint32_t m = ((uint32_t)acc[i] * P::INV) & MASK29;
#pragma unroll
        for (int j = 0; j < 9; ++j) madw(acc[i + j], m, P::N(j));
        acc[i + 1] += acc[i] >> 29;
    }
    F29<P> r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.l[k] = (uint32_t)acc[9 + k] & MASK29;
        acc[10 + k] += acc[9 + k] >> 29;
    }
    r.l[8] = (uint32_t)acc[17];
    return r;
}

template <class P>
ZKB_HD F29<P> sqr29(const F29<P>& a) {
    return mul29<P>(a, a);
}

// loose 9 x 29 (R' = 2^261, any value < 2^261) -> canonical memory form (R = 2^256): exact division by 2^5 via a
// 5-bit Montgomery step, then full carry and conditional subtractions.
template <class P>
ZKB_HD Fp<typename P::P32> from29(const F29<P>& a) {
    // t = a + m*p with m = a * (-p^-1) mod 32, then t >> 5
    uint64_t c[9];
    uint32_t m = (a.l[0] * P::INV) & 31u;
#pragma unroll
    for (int i = 0; i < 9; ++i) { c[i] = a.l[i]; madw(c[i], m, P::N(i)); }
    // full carry propagation into 29-bit limbs
    uint32_t t[10];
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        uint64_t v = c[i] + carry;
        t[i] = (uint32_t)v & MASK29;
        carry = v >> 29;
    }
    t[9] = (uint32_t)carry;
    // pack (t >> 5) into 8 x 32-bit words plus overflow word
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        // word k holds bits [32k + 5, 32k + 37) of t
        int bit = 32 * k + 5;
        uint64_t v = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            int lo = 29 * i;  // limb i covers [lo, lo + 29)
            if (lo + 29 > bit && lo < bit + 32) {
                int sh = lo - bit;
                v |= sh >= 0 ? ((uint64_t)t[i] << sh) : ((uint64_t)t[i] >> (-sh));
            }
        }
        w[k] = (uint32_t)v;
    }
    // value < 2^261 / 32 = 2^256: fits 8 words (w[8] == 0); reduce mod p by conditional subtraction
    Fp<typename P::P32> r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = w[k];
    // value < 2^256 < 6p: subtract 4p, 2p, p conditionally
    using P32 = typename P::P32;
#pragma unroll
    for (int s = 2; s >= 0; --s) {
        uint32_t mm[8], d[8];
        uint64_t cc = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // mm = p << s
            uint64_t v = ((uint64_t)P32::M(k) << s) | cc;
            mm[k] = (uint32_t)v;
            cc = v >> 32;
        }
        uint64_t br = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t v = (uint64_t)r.l[k] - mm[k] - br;
            d[k] = (uint32_t)v;
            br = (v >> 32) & 1;
        }
        if (!br) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r.l[k] = d[k];
        }
    }
    return r;
}

using Fq29 = F29<Fq29Params>;
using Fr29 = F29<Fr29Params>;

}  // namespace zkb
