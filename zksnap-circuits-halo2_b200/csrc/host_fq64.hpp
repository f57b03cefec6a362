// host_fq64.hpp — HOST-ONLY Fq arithmetic on 4 x 64-bit limbs (unsigned __int128 products) for the per-commit host fold.
//
// Every commitment ends on the host (north_star: "tiny bucket sums combined on the host"): the c-bit Horner combination of the window
// sums when there is no SRS window table, the sum of the per-device partials, and ALWAYS the normalisation to the G1 output encoding —
// one inversion.  The portable twins of the device's 8 x 32-bit carry chains (field.cuh, shared with the CPU emulator) cost ~340 ns per
// product on a host core, 130 us per inversion: a third of a 2^13-point commit.  This file is the same arithmetic written for a
// 64-bit CPU (CIOS Montgomery product ~25 ns, a 4-bit-window Fermat inversion ~8 us).  Values are canonical Montgomery residues on
// both sides, so the results are bit-identical to the portable path (tests/test_emulator.py compares the two on random and edge inputs).
#pragma once
#include <cstdint>

#include "curve.cuh"

namespace zkb {
namespace host64 {

typedef unsigned __int128 u128;
struct F {
    uint64_t l[4];
};

inline const uint64_t* modulus() {
    static const uint64_t m[4] = {
        (uint64_t)FqParams::M(0) | ((uint64_t)FqParams::M(1) << 32), (uint64_t)FqParams::M(2) | ((uint64_t)FqParams::M(3) << 32),
        (uint64_t)FqParams::M(4) | ((uint64_t)FqParams::M(5) << 32), (uint64_t)FqParams::M(6) | ((uint64_t)FqParams::M(7) << 32)};
    return m;
}
// -p^-1 mod 2^64 by Newton iteration from the low limb (no constant to trust)
inline uint64_t neg_inv64() {
    static const uint64_t v = [] {
        const uint64_t p0 = modulus()[0];
        uint64_t x = 1;
        for (int i = 0; i < 6; ++i) x *= 2 - p0 * x;
        return (uint64_t)0 - x;
    }();
    return v;
}
inline F from_fq(const Fq& a) {
    F r;
    for (int i = 0; i < 4; ++i) r.l[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
    return r;
}
inline Fq to_fq(const F& a) {
    Fq r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)a.l[i]; r.l[2 * i + 1] = (uint32_t)(a.l[i] >> 32); }
    return r;
}
inline bool is_zero(const F& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool equal(const F& a, const F& b) { return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3]; }
inline bool geq_modulus(const F& a) {
    const uint64_t* M = modulus();
    for (int i = 3; i >= 0; --i)
        if (a.l[i] != M[i]) return a.l[i] > M[i];
    return true;
}
inline void sub_modulus(F& a) {
    const uint64_t* M = modulus();
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        const u128 t = (u128)a.l[i] - M[i] - br;
        a.l[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}
inline F add(const F& a, const F& b) {   // canonical in, canonical out (2p < 2^255: the sum fits 4 limbs)
    F r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (geq_modulus(r)) sub_modulus(r);
    return r;
}
inline F sub(const F& a, const F& b) {
    const uint64_t* M = modulus();
    F r;
    u128 br = 0;
    for (int i = 0; i < 4; ++i) { const u128 t = (u128)a.l[i] - b.l[i] - br; r.l[i] = (uint64_t)t; br = (t >> 64) & 1; }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + M[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
}
inline F dbl(const F& a) { return add(a, a); }
// Montgomery product a b 2^-256 mod p (CIOS), canonical
inline F mul(const F& a, const F& b) {
    const uint64_t* M = modulus();
    const uint64_t ninv = neg_inv64();
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * ninv;
        c = (u128)m * M[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) { c += (u128)m * M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    F r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_modulus(r)) sub_modulus(r);
    return r;
}
inline F sqr(const F& a) { return mul(a, a); }
// a^(p-2) with a 4-bit fixed window (256 squarings + <= 64 + 14 products); a = 0 gives 0
inline F inv(const F& a) {
    const uint64_t* M = modulus();
    uint64_t e[4] = {M[0] - 2, M[1], M[2], M[3]};   // p - 2: the low limb of p is odd and > 2, no borrow
    F tab[16];
    tab[1] = a;
    for (int i = 2; i < 16; ++i) tab[i] = mul(tab[i - 1], a);
    F acc = {{0, 0, 0, 0}};
    bool started = false;
    for (int nib = 63; nib >= 0; --nib) {
        const uint32_t w = (uint32_t)(e[nib >> 4] >> ((nib & 15) * 4)) & 15u;
        if (started) { acc = sqr(acc); acc = sqr(acc); acc = sqr(acc); acc = sqr(acc); }
        if (w) {
            if (started) acc = mul(acc, tab[w]);
            else { acc = tab[w]; started = true; }
        }
    }
    return acc;
}

// ---- XYZZ points over F (the formulas of curve.cuh: dbl-2008-s-1 and add-2008-s, a = 0) -----------------------------------------
struct P {
    F x, y, zz, zzz;
};
inline P identity() { return P{{{0, 0, 0, 0}}, {{0, 0, 0, 0}}, {{0, 0, 0, 0}}, {{0, 0, 0, 0}}}; }
inline bool is_identity(const P& p) { return is_zero(p.zz); }
inline P from_xyzz(const XYZZ& p) { return P{from_fq(p.x), from_fq(p.y), from_fq(p.zz), from_fq(p.zzz)}; }
inline P doubled(const P& p) {
    if (is_identity(p)) return p;
    const F u = dbl(p.y), v = sqr(u), w = mul(u, v), s = mul(p.x, v);
    const F xx = sqr(p.x), m = add(dbl(xx), xx);
    P r;
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = mul(v, p.zz);
    r.zzz = mul(w, p.zzz);
    return r;
}
inline P added(const P& a, const P& b) {
    if (is_identity(a)) return b;
    if (is_identity(b)) return a;
    const F u1 = mul(a.x, b.zz), u2 = mul(b.x, a.zz), s1 = mul(a.y, b.zzz), s2 = mul(b.y, a.zzz);
    const F p = sub(u2, u1), r = sub(s2, s1);
    if (is_zero(p)) return is_zero(r) ? doubled(a) : identity();
    const F pp = sqr(p), ppp = mul(p, pp), q = mul(u1, pp);
    P o;
    o.x = sub(sub(sqr(r), ppp), dbl(q));
    o.y = sub(mul(r, sub(q, o.x)), mul(s1, ppp));
    o.zz = mul(mul(a.zz, b.zz), pp);
    o.zzz = mul(mul(a.zzz, b.zzz), ppp);
    return o;
}
// sum_w 2^(c w) S_w by Horner from the top window
inline P combine_windows(const XYZZ* sums, uint32_t nwin, uint32_t c) {
    P acc = identity();
    for (uint32_t w = nwin; w-- > 0;) {
        for (uint32_t i = 0; i < c; ++i) acc = doubled(acc);
        acc = added(acc, from_xyzz(sums[w]));
    }
    return acc;
}
// G1 output encoding: (x, y, R) normalised, identity (0, R, 0); 12 little-endian 64-bit limbs
inline void to_out(const P& p, uint64_t out[12]) {
    const F one = from_fq(Fq::one());
    if (is_identity(p)) {
        for (int i = 0; i < 12; ++i) out[i] = 0;
        for (int i = 0; i < 4; ++i) out[4 + i] = one.l[i];
        return;
    }
    const F iv = inv(mul(p.zz, p.zzz));
    const F x = mul(p.x, mul(iv, p.zzz)), y = mul(p.y, mul(iv, p.zz));
    for (int i = 0; i < 4; ++i) { out[i] = x.l[i]; out[4 + i] = y.l[i]; out[8 + i] = one.l[i]; }
}
// a Jacobian point (x, y, z) in the output encoding -> XYZZ (x, y, z^2, z^3); z = 0: identity
inline P from_jacobian(const uint64_t p[12]) {
    P r;
    F z;
    for (int i = 0; i < 4; ++i) { r.x.l[i] = p[i]; r.y.l[i] = p[4 + i]; z.l[i] = p[8 + i]; }
    if (is_zero(z)) return identity();
    r.zz = sqr(z);
    r.zzz = mul(r.zz, z);
    return r;
}

}  // namespace host64
}  // namespace zkb
