// curve.cuh — BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Replaces halo2curves::bn256::{G1Affine, G1} add / double / mixed add on the device.
// In-memory types match halo2curves: G1Affine = (x, y) Montgomery, identity (0,0); G1 = (x, y, z) Jacobian,
// identity z = 0.  Bucket accumulators use extended Jacobian ("XYZZ": x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2,
// identity ZZ = 0): the mixed add is 8M+2S and needs no inversion; the group law is exact, so the affine
// result is independent of the representation.
#pragma once
#include "field.cuh"

namespace zkb {

struct Affine {
    Fq x, y;
    ZKB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    ZKB_HD bool is_identity() const { return zz.is_zero(); }
    ZKB_HD static XYZZ identity() {
        XYZZ r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    ZKB_HD static XYZZ from_affine(const Affine& p) {
        XYZZ r;
        if (p.is_identity()) return identity();
        r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }
    ZKB_HD static XYZZ load(const void* p) {
        const char* c = reinterpret_cast<const char*>(p);
        XYZZ r;
        r.x = Fq::load(c); r.y = Fq::load(c + 32); r.zz = Fq::load(c + 64); r.zzz = Fq::load(c + 96);
        return r;
    }
    ZKB_HD void store(void* p) const {
        char* c = reinterpret_cast<char*>(p);
        x.store(c); y.store(c + 32); zz.store(c + 64); zzz.store(c + 96);
    }
};

// 2*(x,y) for an affine point, y != 0 on this curve (odd order group)   [mdbl-2008-s-1, a = 0]
ZKB_HD XYZZ xyzz_double_affine(const Fq& x1, const Fq& y1) {
    XYZZ r;
    Fq u = fp_dbl(y1);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(x1, v);
    Fq xx = fp_sqr(x1);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, y1));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2*p   [dbl-2008-s-1, a = 0]
ZKB_HD XYZZ xyzz_double(const XYZZ& p) {
    if (p.is_identity()) return p;
    XYZZ r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// acc += (x2, y2) affine, not the identity   [madd-2008-s], with the exceptional cases handled exactly
ZKB_HD void xyzz_add_mixed(XYZZ& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_identity()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq u2 = fp_mul(x2, acc.zz);
    Fq s2 = fp_mul(y2, acc.zzz);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_double_affine(x2, y2);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}

// The same mixed addition on a LAZY accumulator: coordinates are only < 2p (congruent mod p), every product skips its final
// conditional subtraction (-17 of 184 instructions each) and add/sub work mod 2p; (x2, y2) is canonical.  A computed ZZ is
// never congruent to 0 (it is a product of non-zero factors), so the all-zero identity test stays valid.  Callers make
// the accumulator canonical (xyzz_canon) before it leaves the thread.
ZKB_HD void xyzz_add_mixed_lazy(XYZZ& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_identity()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq u2 = fp_mul_lazy(x2, acc.zz);
    Fq s2 = fp_mul_lazy(y2, acc.zzz);
    Fq p = fp_sub_lazy(u2, acc.x);
    Fq r = fp_sub_lazy(s2, acc.y);
    if (fp_is_zero_lazy(p)) {
        if (fp_is_zero_lazy(r)) acc = xyzz_double_affine(x2, y2);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = fp_mul_lazy(p, p);
    Fq ppp = fp_mul_lazy(p, pp);
    Fq q = fp_mul_lazy(acc.x, pp);
    Fq x3 = fp_sub_lazy(fp_sub_lazy(fp_mul_lazy(r, r), ppp), fp_add_lazy(q, q));
    Fq y3 = fp_sub_lazy(fp_mul_lazy(r, fp_sub_lazy(q, x3)), fp_mul_lazy(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul_lazy(acc.zz, pp);
    acc.zzz = fp_mul_lazy(acc.zzz, ppp);
}
ZKB_HD XYZZ xyzz_canon(const XYZZ& p) {
    XYZZ r;
    r.x = fp_canon(p.x); r.y = fp_canon(p.y); r.zz = fp_canon(p.zz); r.zzz = fp_canon(p.zzz);
    return r;
}

// acc += b   [add-2008-s]
ZKB_HD void xyzz_add(XYZZ& acc, const XYZZ& b) {
    if (b.is_identity()) return;
    if (acc.is_identity()) { acc = b; return; }
    Fq u1 = fp_mul(acc.x, b.zz);
    Fq u2 = fp_mul(b.x, acc.zz);
    Fq s1 = fp_mul(acc.y, b.zzz);
    Fq s2 = fp_mul(b.y, acc.zzz);
    Fq p = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_double(acc);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(u1, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(fp_mul(acc.zz, b.zz), pp);
    acc.zzz = fp_mul(fp_mul(acc.zzz, b.zzz), ppp);
}

ZKB_HD Affine affine_load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    Affine a;
    a.x = Fq::load_nc(c);
    a.y = Fq::load_nc(c + 32);
    return a;
}

}  // namespace zkb
