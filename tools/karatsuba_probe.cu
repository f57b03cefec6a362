// Karatsuba + separated Montgomery reduction prototype: SASS instruction-mix probe and correctness self-check
#include "../zksnap-circuits-halo2_b200/csrc/field.cuh"
using namespace zkb;

// o[0..7] = x[0..3] * y[0..3]   (schoolbook 4x4, row-wise, even/odd chains)
__device__ __forceinline__ void mul4(uint32_t (&o)[8], const uint32_t* x, const uint32_t* y) {
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // even chain: (o[i],o[i+1]) += x0*y_i ; (o[i+2],o[i+3]) += x2*y_i ; carry -> o[i+4] (, o[i+5])
        if (i < 3) {
            asm("mad.lo.cc.u32 %0, %6, %8, %0;\n\tmadc.hi.cc.u32 %1, %6, %8, %1;\n\t"
                "madc.lo.cc.u32 %2, %7, %8, %2;\n\tmadc.hi.cc.u32 %3, %7, %8, %3;\n\t"
                "addc.cc.u32 %4, %4, 0;\n\taddc.u32 %5, %5, 0;\n\t"
                : "+r"(o[i]), "+r"(o[i + 1]), "+r"(o[i + 2]), "+r"(o[i + 3]), "+r"(o[i + 4]), "+r"(o[i + 5])
                : "r"(x[0]), "r"(x[2]), "r"(y[i]));
            asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
                "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
                "addc.u32 %4, %4, 0;\n\t"
                : "+r"(o[i + 1]), "+r"(o[i + 2]), "+r"(o[i + 3]), "+r"(o[i + 4]), "+r"(o[i + 5])
                : "r"(x[1]), "r"(x[3]), "r"(y[i]));
        } else {
            asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
                "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
                "addc.u32 %4, %4, 0;\n\t"
                : "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
                : "r"(x[0]), "r"(x[2]), "r"(y[3]));
            asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\t"
                "madc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.u32 %3, %5, %6, %3;\n\t"
                : "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
                : "r"(x[1]), "r"(x[3]), "r"(y[3]));
        }
    }
}

template <class P>
__device__ __forceinline__ Fp<P> fp_mul_kara(const Fp<P>& a, const Fp<P>& b) {
    uint32_t z0[8], z2[8], z1[8], sa[4], sb[4], ca, cb;
    mul4(z0, a.l, b.l);
    mul4(z2, a.l + 4, b.l + 4);
    asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;\n\t"
        : "=r"(sa[0]), "=r"(sa[1]), "=r"(sa[2]), "=r"(sa[3]), "=r"(ca)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
    asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;\n\t"
        : "=r"(sb[0]), "=r"(sb[1]), "=r"(sb[2]), "=r"(sb[3]), "=r"(cb)
        : "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    mul4(z1, sa, sb);
    // mid (9 limbs) = z1 + ca*sb*2^128 + cb*sa*2^128 + ca*cb*2^256 - z0 - z2
    uint32_t m8 = ca & cb;
    uint32_t ma = 0u - ca, mb = 0u - cb;
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;\n\t"
        : "+r"(z1[4]), "+r"(z1[5]), "+r"(z1[6]), "+r"(z1[7]), "+r"(m8)
        : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma));
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;\n\t"
        : "+r"(z1[4]), "+r"(z1[5]), "+r"(z1[6]), "+r"(z1[7]), "+r"(m8)
        : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;\n\t"
        : "+r"(z1[0]), "+r"(z1[1]), "+r"(z1[2]), "+r"(z1[3]), "+r"(z1[4]), "+r"(z1[5]), "+r"(z1[6]), "+r"(z1[7]), "+r"(m8)
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]));
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;\n\t"
        : "+r"(z1[0]), "+r"(z1[1]), "+r"(z1[2]), "+r"(z1[3]), "+r"(z1[4]), "+r"(z1[5]), "+r"(z1[6]), "+r"(z1[7]), "+r"(m8)
        : "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    // t = z0 + mid * 2^128 + z2 * 2^256  (16 limbs): t[0..3] = z0[0..3]; t[4..12] += mid; carry ripples through t[13..15]
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { t[k] = z0[k]; t[8 + k] = z2[k]; }
    asm("add.cc.u32 %0, %0, %12;\n\taddc.cc.u32 %1, %1, %13;\n\taddc.cc.u32 %2, %2, %14;\n\taddc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %18;\n\taddc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\taddc.cc.u32 %9, %9, 0;\n\taddc.cc.u32 %10, %10, 0;\n\taddc.u32 %11, %11, 0;\n\t"
        : "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]),
          "+r"(t[14]), "+r"(t[15])
        : "r"(z1[0]), "r"(z1[1]), "r"(z1[2]), "r"(z1[3]), "r"(z1[4]), "r"(z1[5]), "r"(z1[6]), "r"(z1[7]), "r"(m8));
    // Montgomery reduction, carries out of each row deferred in c[]
    uint32_t c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t m = t[i] * P::INV;
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;\n\t"
            : "+r"(t[i]), "+r"(t[i + 1]), "+r"(t[i + 2]), "+r"(t[i + 3]), "+r"(t[i + 4]), "+r"(t[i + 5]), "+r"(t[i + 6]), "+r"(t[i + 7]),
              "+r"(c[i])
            : "r"(P::M(0)), "r"(P::M(2)), "r"(P::M(4)), "r"(P::M(6)), "r"(m));
        if (i < 7) {
            asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
                "addc.u32 %8, %8, 0;\n\t"
                : "+r"(t[i + 1]), "+r"(t[i + 2]), "+r"(t[i + 3]), "+r"(t[i + 4]), "+r"(t[i + 5]), "+r"(t[i + 6]), "+r"(t[i + 7]),
                  "+r"(t[i + 8]), "+r"(c[i + 1])
                : "r"(P::M(1)), "r"(P::M(3)), "r"(P::M(5)), "r"(P::M(7)), "r"(m));
        } else {
            asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
                "addc.u32 %8, %8, 0;\n\t"
                : "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15]), "+r"(c[8])
                : "r"(P::M(1)), "r"(P::M(3)), "r"(P::M(5)), "r"(P::M(7)), "r"(m));
        }
    }
    // carries: even chain of row i overflows limb i+7 -> lands at limb i+8 (c[i]); odd chain -> limb i+9 (c[i+1])
    // result = t[8..15] + (c[0] at limb 8) + (c[1] + c[1]' at limb 9) ...: c[k] holds the carries destined for limb 8 + k
    uint32_t r[8];
    asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]),
          "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]));
    reduce_once<P>(r);
    Fp<P> out;
#pragma unroll
    for (int k = 0; k < 8; ++k) out.l[k] = r[k];
    return out;
}

__global__ void __launch_bounds__(128) probe(uint4* out, const uint4* in, int iters) {
    Fq x = Fq::load(in + 2 * threadIdx.x), y = Fq::load(in + 2 * threadIdx.x + 512);
    for (int i = 0; i < iters; ++i) x = fp_mul_kara(x, y);
    x.store(out + 2 * threadIdx.x);
}
__global__ void check(const uint4* in, uint32_t* bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq x = Fq::load(in + 4 * i), y = Fq::load(in + 4 * i + 2);
    if (!(fp_mul_kara(x, y) == fp_mul(x, y))) atomicAdd(bad, 1u);
}
