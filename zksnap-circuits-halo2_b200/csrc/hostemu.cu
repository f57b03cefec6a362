// hostemu.cu — CPU emulator of the device code (TEST INFRASTRUCTURE; built as libzkb200_hostemu.so).
//
// The kernels in this package are written as __host__ __device__ per-thread phase functions.  This file runs
// exactly those functions on the CPU, thread by thread and CTA by CTA (the PTX carry chains are replaced by
// their portable twins in field.cuh), so limb arithmetic, index maths, digit reversal, chunk-boundary logic
// etc. are validated against the oracle in the `-m "not gpu"` tests of this repo, where no GPU exists.
// It is NOT a CPU fallback: libzkb200.so never links it and the product fails loudly without a CUDA device.
#include <cstdint>
#include <cstring>
#include <vector>

#include "curve.cuh"
#include "../../tools/field29.cuh"  // rejected 9 x 29-bit prototype, kept with the microbenchmarks (not on any product path)
#include "msm.cuh"
#include "ntt.cuh"
#include "ntt_plan.hpp"

using namespace zkb;

#define EMU_EXPORT extern "C" __attribute__((visibility("default")))

static Fr fr_from_u64(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}
template <class F>
static void f_to_u64(const F& v, uint64_t* p) {
    for (int i = 0; i < 4; ++i) p[i] = (uint64_t)v.l[2 * i] | ((uint64_t)v.l[2 * i + 1] << 32);
}
static Fq fq_from_u64(const uint64_t* p) {
    Fq r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}

// field: 0 Fr, 1 Fq; op: 0 mul, 1 add, 2 sub; lazy forms (inputs < 2M, result made canonical for comparison):
// 4 canon(mul_lazy), 5 canon(add_lazy), 6 canon(sub_lazy), 7 is_zero_lazy(a) (0/1 in limb 0), 8 raw mul_lazy (must be < 2M)
template <class F>
static F emu_field_op(int op, const F& x, const F& y) {
    switch (op) {
        case 0: return fp_mul(x, y);
        case 1: return fp_add(x, y);
        case 2: return fp_sub(x, y);
        case 4: return fp_canon(fp_mul_lazy(x, y));
        case 5: return fp_canon(fp_add_lazy(x, y));
        case 6: return fp_canon(fp_sub_lazy(x, y));
        case 7: { F r = F::zero(); r.l[0] = fp_is_zero_lazy(x) ? 1u : 0u; return r; }
        default: return fp_mul_lazy(x, y);
    }
}
EMU_EXPORT void zkb_emu_vec_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) f_to_u64(emu_field_op(op, fr_from_u64(a + 4 * i), fr_from_u64(b + 4 * i)), o + 4 * i);
        else f_to_u64(emu_field_op(op, fq_from_u64(a + 4 * i), fq_from_u64(b + 4 * i)), o + 4 * i);
    }
}

// 29-bit-limb arithmetic: op 0: from29(mul29(to29 a, to29 b)); 1: lazy add; 2: sub29<K=34> + carry; 3: chain
// ((a*b) - a + b) * (a - b) exercising lazy inputs to mul29.  field: 0 Fr, 1 Fq
template <class P29, class F>
static F emu29_one(int op, const F& x, const F& y) {
    F29<P29> a = to29<P29>(x), b = to29<P29>(y);
    F29<P29> r;
    if (op == 0) r = mul29<P29>(a, b);
    else if (op == 1) r = add29<P29>(a, b);
    else if (op == 2) r = carry29<P29>(sub29<P29, 34>(a, b));
    else {
        F29<P29> ab = mul29<P29>(a, b);                          // < 7.1 p
        F29<P29> t = add29<P29>(carry29<P29>(sub29<P29, 34>(ab, a)), b);   // < 7.1p + 34p + 32p, lazy limbs
        F29<P29> u = carry29<P29>(sub29<P29, 34>(a, b));         // < 66 p
        r = mul29<P29>(carry29<P29>(t), u);                      // product < 74*66 p^2 = 4884 p^2 > 169 p^2 * 28: result < 30 p
    }
    return from29<P29>(r);
}
EMU_EXPORT void zkb_emu_vec_op29(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) f_to_u64(emu29_one<Fr29Params>(op, fr_from_u64(a + 4 * i), fr_from_u64(b + 4 * i)), o + 4 * i);
        else f_to_u64(emu29_one<Fq29Params>(op, fq_from_u64(a + 4 * i), fq_from_u64(b + 4 * i)), o + 4 * i);
    }
}

// ---- NTT -----------------------------------------------------------------------------------------------------
template <int LOGR>
static void emu_run_pass(const NttPassArgs& a, uint32_t nthreads, uint64_t nctas, uint32_t cols, size_t smem_bytes) {
    std::vector<uint4> sm(smem_bytes / 16);
    for (uint32_t col = 0; col < cols; ++col)
        for (uint64_t cta_local = 0; cta_local < nctas; ++cta_local) {
            const uint64_t cta = ntt_dist_cta(a, cta_local);
            for (uint32_t t = 0; t < nthreads; ++t) ntt_phase_load<LOGR>(a, sm.data(), t, nthreads, cta, col);
            int left = LOGR;
            while (left > 0) {
                int r = left >= 3 ? 3 : left;
                for (uint32_t t = 0; t < nthreads; ++t)
                    ntt_phase_round<LOGR>(a, sm.data(), t, nthreads, (uint32_t)left, (uint32_t)r);
                left -= r;
            }
            for (uint32_t t = 0; t < nthreads; ++t) ntt_phase_store<LOGR>(a, sm.data(), t, nthreads, cta, col);
        }
}

static void emu_dispatch_pass(const NttPassArgs& a, uint32_t logr, uint32_t nthreads, uint64_t nctas, uint32_t cols,
                              size_t smem) {
    switch (logr) {
#define C(L) case L: emu_run_pass<L>(a, nthreads, nctas, cols, smem); break;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10)
#undef C
        default: break;
    }
}

static void pow_table(std::vector<uint4>& out, const Fr& omega, uint64_t count, uint32_t shift) {
    out.resize(2 * count);
    for (uint64_t t = 0; t < count; ++t) fr_store2(out.data(), t, fp_pow_u64(omega, t << shift));
}

// per-pass inter-pass twiddle table, same entry function as the device kernel
static void pass_table(std::vector<uint4>& out, const std::vector<uint4>& tw_hi, const std::vector<uint4>& tw_lo, uint32_t tw_h,
                       uint32_t log_r, uint32_t shift, uint64_t count) {
    out.resize(2 * count);
    for (uint64_t t = 0; t < count; ++t) fr_store2(out.data(), t, ntt_pass_twiddle(tw_hi.data(), tw_lo.data(), tw_h, log_r, shift, t));
}
static bool g_emu_pass_tables = true;
EMU_EXPORT void zkb_emu_set_pass_tables(int on) { g_emu_pass_tables = on != 0; }

// in: cols x in_len elements (column stride in_len); out: cols x 2^log_n.  *_scale3: 3x4 u64 Montgomery or NULL.
EMU_EXPORT int zkb_emu_ntt(const uint64_t* in, uint64_t in_len, uint64_t* out, uint32_t log_n, const uint64_t* omega,
                           const uint64_t* in_scale3, const uint64_t* out_scale3, uint32_t cols) {
    if (log_n < 1 || log_n > 28) return -1;
    const uint64_t N = 1ull << log_n;
    NttGeometry g = ntt_geometry(log_n);
    Fr w = fr_from_u64(omega);
    std::vector<uint4> tw_lo, tw_hi, tw_r[NTT_MAX_PASSES];
    pow_table(tw_lo, w, 1ull << g.tw_h, 0);
    pow_table(tw_hi, w, N >> g.tw_h ? N >> g.tw_h : 1, g.tw_h);
    for (uint32_t p = 0; p < g.npass; ++p) pow_table(tw_r[p], w, 1ull << g.lr[p], log_n - g.lr[p]);
    std::vector<uint4> work(2 * N * cols);
    for (uint32_t p = 0; p < g.npass; ++p) {
        NttPassArgs a{};
        bool fin = p + 1 == g.npass;
        a.src = p == 0 ? reinterpret_cast<const uint4*>(in) : work.data();
        a.src_col_stride = p == 0 ? in_len : N;
        a.dst = fin ? reinterpret_cast<uint4*>(out) : work.data();
        a.dst_col_stride = N;
        a.log_n = log_n; a.npass = g.npass; a.pass = p;
        for (uint32_t q = 0; q < g.npass; ++q) a.lr[q] = g.lr[q];
        a.log_t = g.log_t[p];
        a.is_final = fin;
        a.tw_r = tw_r[p].data(); a.tw_hi = tw_hi.data(); a.tw_lo = tw_lo.data(); a.tw_h = g.tw_h;
        std::vector<uint4> tw_pass;
        if (p > 0 && g_emu_pass_tables) {
            uint32_t log_q = 0;
            for (uint32_t q = 0; q <= p; ++q) log_q += g.lr[q];
            pass_table(tw_pass, tw_hi, tw_lo, g.tw_h, g.lr[p], log_n - log_q, 1ull << log_q);
            a.tw_pass = tw_pass.data();
        }
        a.in_len = p == 0 ? in_len : N;
        a.in_scale_on = (p == 0 && in_scale3) ? 1 : 0;
        a.out_scale_on = (fin && out_scale3) ? 1 : 0;
        for (int m = 0; m < 3; ++m)
            for (int i = 0; i < 4; ++i) {
                if (in_scale3) { a.in_scale[m][2 * i] = (uint32_t)in_scale3[4 * m + i]; a.in_scale[m][2 * i + 1] = (uint32_t)(in_scale3[4 * m + i] >> 32); }
                if (out_scale3) { a.out_scale[m][2 * i] = (uint32_t)out_scale3[4 * m + i]; a.out_scale[m][2 * i + 1] = (uint32_t)(out_scale3[4 * m + i] >> 32); }
            }
        emu_dispatch_pass(a, g.lr[p], ntt_cta_threads(g, p), ntt_cta_count(g, p), cols, ntt_cta_smem_bytes(g, p));
    }
    return 0;
}


// ---- distributed NTT (one transform sharded over 2^log_g ranks, peer memory emulated by plain pointers) -----------
// phase 0: pass 0 of `rank` (loads from every rank's A slice, stores to every rank's W slice);
// phase 1: the remaining passes of `rank` (middle passes local in W, final pass scatters to every rank's O slice).
// The caller provides the barrier between the phases (a loop over ranks in one process, or a gloo barrier between
// processes sharing the slices through POSIX shared memory).
EMU_EXPORT int zkb_emu_ntt_dist_phase(int phase, uint32_t rank, uint32_t log_g, uint32_t log_n, const uint64_t* omega,
                                      uint64_t* const* A, uint64_t* const* W, uint64_t* const* O) {
    if (log_n < 1 || log_n > 28 || log_g > 3) return -1;
    const uint64_t N = 1ull << log_n;
    NttGeometry g = ntt_geometry(log_n, true);
    if (!ntt_dist_supported(g, log_g)) return -2;
    Fr w = fr_from_u64(omega);
    std::vector<uint4> tw_lo, tw_hi, tw_r[NTT_MAX_PASSES];
    pow_table(tw_lo, w, 1ull << g.tw_h, 0);
    pow_table(tw_hi, w, N >> g.tw_h ? N >> g.tw_h : 1, g.tw_h);
    for (uint32_t p = 0; p < g.npass; ++p) pow_table(tw_r[p], w, 1ull << g.lr[p], log_n - g.lr[p]);
    for (uint32_t p = 0; p < g.npass; ++p) {
        if ((phase == 0) != (p == 0)) continue;
        NttPassArgs a{};
        bool fin = p + 1 == g.npass;
        a.log_n = log_n; a.npass = g.npass; a.pass = p;
        for (uint32_t q = 0; q < g.npass; ++q) a.lr[q] = g.lr[q];
        a.log_t = g.log_t[p];
        a.is_final = fin;
        a.tw_r = tw_r[p].data(); a.tw_hi = tw_hi.data(); a.tw_lo = tw_lo.data(); a.tw_h = g.tw_h;
        std::vector<uint4> tw_pass;
        if (p > 0 && g_emu_pass_tables) {
            uint32_t log_q = 0;
            for (uint32_t q = 0; q <= p; ++q) log_q += g.lr[q];
            pass_table(tw_pass, tw_hi, tw_lo, g.tw_h, g.lr[p], log_n - log_q, 1ull << log_q);
            a.tw_pass = tw_pass.data();
        }
        a.in_len = N;
        a.dist_log_g = log_g; a.dist_rank = rank; a.dist_log_slice = log_n - log_g;
        for (uint32_t r = 0; r < (1u << log_g); ++r) {
            a.peer_src[r] = reinterpret_cast<const uint4*>(p == 0 ? A[r] : W[r]);
            a.peer_dst[r] = reinterpret_cast<uint4*>(fin ? O[r] : W[r]);
        }
        emu_dispatch_pass(a, g.lr[p], ntt_cta_threads(g, p), ntt_cta_count(g, p) >> log_g, 1, ntt_cta_smem_bytes(g, p));
    }
    return 0;
}

// ---- MSM -----------------------------------------------------------------------------------------------------
#include "hostemu_msm.inc"

// ---- host_fq64.hpp (the product's host fold in 64-bit arithmetic) against the portable twins of the device code ---------------------
#include "host_fq64.hpp"
// returns the number of mismatches over `n` random cases plus the edge cases (identity operands, equal points, opposite points, zero)
EMU_EXPORT uint64_t zkb_emu_host64_check(uint64_t seed, uint64_t n) {
    uint64_t bad = 0, st = seed * 0x9e3779b97f4a7c15ull + 1;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    auto rnd_fq = [&]() {
        Fq v;
        for (int i = 0; i < 8; ++i) v.l[i] = (uint32_t)rnd();
        v.l[7] &= 0x0fffffffu;   // < 2^252 < p: a canonical residue
        return v;
    };
    auto same = [&](const host64::F& a, const Fq& b) { return host64::equal(a, host64::from_fq(b)); };
    auto same_point = [&](const host64::P& a, const XYZZ& b) {   // compare the normalised output encodings
        uint64_t oa[12], ob[12];
        host64::to_out(a, oa);
        xyzz_to_jacobian_u64(b, ob);
        return memcmp(oa, ob, 96) == 0;
    };
    for (uint64_t i = 0; i < n; ++i) {
        const Fq a = rnd_fq(), b = rnd_fq();
        const host64::F fa = host64::from_fq(a), fb = host64::from_fq(b);
        if (!same(host64::mul(fa, fb), fp_mul(a, b))) ++bad;
        if (!same(host64::add(fa, fb), fp_add(a, b))) ++bad;
        if (!same(host64::sub(fa, fb), fp_sub(a, b))) ++bad;
        XYZZ p, q;
        p.x = rnd_fq(); p.y = rnd_fq(); p.zz = rnd_fq(); p.zzz = rnd_fq();
        q.x = rnd_fq(); q.y = rnd_fq(); q.zz = rnd_fq(); q.zzz = rnd_fq();
        XYZZ s = p;
        xyzz_add(s, q);
        if (!same_point(host64::added(host64::from_xyzz(p), host64::from_xyzz(q)), s)) ++bad;
        if (!same_point(host64::doubled(host64::from_xyzz(p)), xyzz_double(p))) ++bad;
        if (i < 8) {   // the expensive ones: inversion, Horner combination
            if (!same(host64::inv(fa), fq_inv(a))) ++bad;
            XYZZ sums[5] = {p, q, s, XYZZ::identity(), xyzz_double(q)};
            if (!same_point(host64::combine_windows(sums, 5, 7), msm_combine_windows(sums, 5, 7))) ++bad;
        }
    }
    // edge cases
    const Fq zero = Fq::zero(), one = Fq::one();
    Fq pm1;   // p - 1
    for (int i = 0; i < 8; ++i) pm1.l[i] = FqParams::M(i);
    pm1.l[0] -= 1;
    const Fq edge[4] = {zero, one, pm1, fp_neg(one)};
    for (const Fq& a : edge)
        for (const Fq& b : edge) {
            const host64::F fa = host64::from_fq(a), fb = host64::from_fq(b);
            if (!same(host64::mul(fa, fb), fp_mul(a, b))) ++bad;
            if (!same(host64::add(fa, fb), fp_add(a, b))) ++bad;
            if (!same(host64::sub(fa, fb), fp_sub(a, b))) ++bad;
        }
    if (!host64::is_zero(host64::inv(host64::from_fq(zero)))) ++bad;
    XYZZ g;   // the generator (1, 2) and friends
    g.x = one; g.y = fp_dbl(one); g.zz = one; g.zzz = one;
    XYZZ neg = g;
    neg.y = fp_neg(g.y);
    const XYZZ id = XYZZ::identity();
    const XYZZ cases[4] = {g, neg, id, xyzz_double(g)};
    for (const XYZZ& x : cases)
        for (const XYZZ& y : cases) {
            XYZZ s = x;
            xyzz_add(s, y);
            if (!same_point(host64::added(host64::from_xyzz(x), host64::from_xyzz(y)), s)) ++bad;
        }
    // Jacobian input of g1_sum: (x, y, z) with z = 1 and z = 0
    uint64_t jac[12];
    xyzz_to_jacobian_u64(g, jac);
    if (!same_point(host64::from_jacobian(jac), g)) ++bad;
    memset(jac + 8, 0, 32);
    if (!host64::is_identity(host64::from_jacobian(jac))) ++bad;
    return bad;
}
