/*
 * zkb_oracle.c — CPU restatement of the MSM / NTT hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libzkb200.so) never links or calls it.
 *
 * The reference repo (aerius-labs/zksnap-circuits-halo2) contains no field, curve, FFT or MSM code; the
 * path lives in its un-vendored, un-pinned git dependencies halo2-axiom (`halo2_proofs`) and
 * halo2curves-axiom (floating branches; Cargo.lock is git-ignored: /root/reference/.gitignore:2,
 * /root/reference/aggregator/Cargo.toml:7-8).  This file restates those crates' published algorithms
 * and is anchored on the reference's call sites:
 *   keygen_vk/keygen_pk  /root/reference/aggregator/src/wrapper.rs:106-109
 *   create_proof         /root/reference/aggregator/src/wrapper.rs:129-137
 *   ParamsKZG::setup     /root/reference/voter/benches/voter_circuit.rs:60
 *
 * PARITY UNPINNED by reference fixtures (the reference holds no golden vectors for this path).  Pinned
 * instead to first-principles KATs, to the independent big-int twin oracle/pyref.py and, for the
 * curve arithmetic, to the public EIP-196 (alt_bn128) precompile vectors in tests/golden/eip196_kats.json.
 *
 * Restated upstream functions (sources not on disk):
 *   halo2curves::bn256::{Fr,Fq}            4x64 Montgomery, canonical outputs        -> fr_* / fq_*
 *   halo2curves::bn256::{G1Affine,G1}      a=0 Jacobian add / mixed add / double     -> g1_*
 *   halo2_proofs::arithmetic::best_fft      bit-reverse + radix-2 DIT + twiddle table -> zko_best_fft
 *   halo2_proofs::arithmetic::best_multiexp contiguous chunk per thread, serial
 *       Pippenger (c = ceil(ln n), unsigned digits, running-sum), fold                -> zko_best_multiexp
 *   halo2_proofs::poly::EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
 *       extended_to_coeff}                                                            -> zko_*
 *   ParamsKZG::setup's g[i] = [s^i]G (fixed-base scalar mul)                           -> zko_g1_fixed_base_mul
 *   ParamsKZG::setup (G1 side): g[i] = [s^i]G, g_lagrange[i] = [l_i(s)]G with
 *       l_i(s) = omega^i (s^n - 1) / (n (s - omega^i))                                 -> zko_kzg_setup
 *   Curve::batch_normalize (Montgomery's trick)                                        -> zko_g1_batch_normalize
 *   arithmetic::eval_polynomial (Horner)                                               -> zko_fr_eval_polynomial
 *   arithmetic::kate_division(a, b): q_{n-2} = a_{n-1}, q_{i-1} = a_i + b q_i            -> zko_fr_kate_division
 *   ff::BatchInvert (zeros are skipped and stay zero)                                   -> zko_fr_batch_invert
 *   plonk::lookup::prover::permute_expression_pair                                      -> zko_permute_expression_pair
 *   best_fft with G = G1 (FftGroup for curve points), by the DFT definition             -> zko_g1_fft_naive
 *   plonk::evaluation::GraphEvaluator::evaluate in evaluate_h's row loop (intermediates
 *       numbered as upstream numbers them, rotated rows by rem_euclid)                  -> zko_graph_evaluate
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct { u64 l[4]; } fe;

/* ---- minimal pthread parallel-for (static ranges); stands in for rayon's `parallelize` ---- */
typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } pf_job;
static void *pf_tramp(void *p) { pf_job *j = (pf_job *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(0, n, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    pf_job *jobs = (pf_job *)malloc(sizeof(pf_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * (size_t)t / threads; jobs[t].hi = n * (size_t)(t + 1) / threads;
        if (t == threads - 1) pf_tramp(&jobs[t]);
        else pthread_create(&th[t], NULL, pf_tramp, &jobs[t]);
    }
    for (int t = 0; t + 1 < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
static int resolve_threads(int threads) {
    if (threads > 0) return threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    fe m;      /* modulus */
    fe r;      /* 2^256 mod m  (Montgomery one) */
    fe r2;     /* 2^512 mod m */
    u64 inv;   /* -m^-1 mod 2^64 */
} field_t;

static const field_t FQ = {
    {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}},
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}},
    0x87d20782e4866389ULL};

static const field_t FR = {
    {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}},
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}},
    0xc2e1f593efffffffULL};

/* Fr::ROOT_OF_UNITY (2^28-th primitive root, = 7^((r-1)/2^28)) and Fr::ZETA, canonical integers. */
static const fe FR_ROOT_OF_UNITY_CANON =
    {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const fe FR_ZETA_CANON =
    {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};
#define FR_S 28

/* ------------------------------------------------------------------------------------------ */
/* 4x64 Montgomery arithmetic                                                                  */
/* ------------------------------------------------------------------------------------------ */
static inline int fe_geq(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; --i) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline u64 fe_sub_raw(fe *o, const fe *a, const fe *b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a->l[i] - b->l[i] - borrow;
        o->l[i] = (u64)t;
        borrow = (u64)(t >> 64) & 1;
    }
    return borrow;
}
static inline u64 fe_add_raw(fe *o, const fe *a, const fe *b) {
    u64 carry = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a->l[i] + b->l[i] + carry;
        o->l[i] = (u64)t;
        carry = (u64)(t >> 64);
    }
    return carry;
}
static inline void f_add(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t;
    fe_add_raw(&t, a, b); /* both < m < 2^254: no carry out */
    if (fe_geq(&t, &F->m)) fe_sub_raw(&t, &t, &F->m);
    *o = t;
}
static inline void f_sub(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t;
    if (fe_sub_raw(&t, a, b)) fe_add_raw(&t, &t, &F->m);
    *o = t;
}
static inline void f_neg(const field_t *F, fe *o, const fe *a) {
    if (fe_is_zero(a)) { *o = *a; return; }
    fe_sub_raw(o, &F->m, a);
}
static inline void f_dbl(const field_t *F, fe *o, const fe *a) { f_add(F, o, a, a); }

/* CIOS Montgomery product: a*b*2^-256 mod m, canonical. */
static inline void f_mul(const field_t *F, fe *o, const fe *a, const fe *b) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->inv;
        c = (u128)m * F->m.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->m.l[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&r, &F->m)) fe_sub_raw(&r, &r, &F->m);
    *o = r;
}
static inline void f_sqr(const field_t *F, fe *o, const fe *a) { f_mul(F, o, a, a); }
static inline void f_to_mont(const field_t *F, fe *o, const fe *a) { f_mul(F, o, a, &F->r2); }
static inline void f_from_mont(const field_t *F, fe *o, const fe *a) {
    fe one = {{1, 0, 0, 0}};
    f_mul(F, o, a, &one);
}
static void f_pow(const field_t *F, fe *o, const fe *a, const fe *e) {
    fe acc = F->r, base = *a;
    for (int i = 0; i < 256; ++i) {
        if ((e->l[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, &base);
        f_sqr(F, &base, &base);
    }
    *o = acc;
}
static void f_inv(const field_t *F, fe *o, const fe *a) {
    fe e, two = {{2, 0, 0, 0}};
    fe_sub_raw(&e, &F->m, &two);
    f_pow(F, o, a, &e);
}

/* exported scalar helpers (ctypes-visible) */
#define EXPORT __attribute__((visibility("default")))
#define FE(p) ((fe *)(p))
#define CFE(p) ((const fe *)(p))

EXPORT void zko_fr_mul(const u64 *a, const u64 *b, u64 *o) { f_mul(&FR, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fr_add(const u64 *a, const u64 *b, u64 *o) { f_add(&FR, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fr_sub(const u64 *a, const u64 *b, u64 *o) { f_sub(&FR, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fr_inv(const u64 *a, u64 *o) { f_inv(&FR, FE(o), CFE(a)); }
EXPORT void zko_fr_to_mont(const u64 *a, u64 *o) { f_to_mont(&FR, FE(o), CFE(a)); }
EXPORT void zko_fr_from_mont(const u64 *a, u64 *o) { f_from_mont(&FR, FE(o), CFE(a)); }
EXPORT void zko_fq_mul(const u64 *a, const u64 *b, u64 *o) { f_mul(&FQ, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fq_add(const u64 *a, const u64 *b, u64 *o) { f_add(&FQ, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fq_sub(const u64 *a, const u64 *b, u64 *o) { f_sub(&FQ, FE(o), CFE(a), CFE(b)); }
EXPORT void zko_fq_inv(const u64 *a, u64 *o) { f_inv(&FQ, FE(o), CFE(a)); }
EXPORT void zko_fq_to_mont(const u64 *a, u64 *o) { f_to_mont(&FQ, FE(o), CFE(a)); }
EXPORT void zko_fq_from_mont(const u64 *a, u64 *o) { f_from_mont(&FQ, FE(o), CFE(a)); }

/* vector forms so numpy callers do not pay one ctypes call per element; op: 0 mul 1 add 2 sub.
 * field: 0 = Fr, 1 = Fq */
EXPORT void zko_vec_op(int field, int op, const u64 *a, const u64 *b, u64 *o, size_t n) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; ++i) {
        if (op == 0) f_mul(F, FE(o + 4 * i), CFE(a + 4 * i), CFE(b + 4 * i));
        else if (op == 1) f_add(F, FE(o + 4 * i), CFE(a + 4 * i), CFE(b + 4 * i));
        else f_sub(F, FE(o + 4 * i), CFE(a + 4 * i), CFE(b + 4 * i));
    }
}

/* sum_i a_i * b_i in Fr (Montgomery in, Montgomery out) — the known-dlog check for big MSMs */
EXPORT void zko_fr_inner_product(const u64 *a, const u64 *b, size_t n, u64 *o) {
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = 0; i < n; ++i) {
        fe t;
        f_mul(&FR, &t, CFE(a + 4 * i), CFE(b + 4 * i));
        f_add(&FR, &acc, &acc, &t);
    }
    *FE(o) = acc;
}

/* ------------------------------------------------------------------------------------------ */
/* Fr FFT                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; ++i) {
        r = (r << 1) | (n & 1);
        n >>= 1;
    }
    return r;
}

/* omega(k) = ROOT_OF_UNITY^(2^(S-k)), Montgomery form */
static void fr_omega(unsigned k, fe *o) {
    fe w;
    f_to_mont(&FR, &w, &FR_ROOT_OF_UNITY_CANON);
    for (unsigned i = k; i < FR_S; ++i) f_sqr(&FR, &w, &w);
    *o = w;
}
EXPORT void zko_fr_omega(uint32_t k, u64 *o) { fr_omega(k, FE(o)); }
EXPORT void zko_fr_zeta(u64 *o) { f_to_mont(&FR, FE(o), &FR_ZETA_CANON); }

/* halo2_proofs::arithmetic::best_fft: in-place, natural order in/out, Montgomery elements.
 * threads <= 0 -> all available. */
typedef struct { fe *a; fe *tw; const fe *omega; size_t half_n, half, chunk, tchunk, blk; } fft_ctx;
static void fft_twiddle_range(size_t lo, size_t hi, void *p) {
    fft_ctx *c = (fft_ctx *)p;
    for (size_t b = lo; b < hi; ++b) {
        fe e = {{b * c->blk, 0, 0, 0}}, w;
        f_pow(&FR, &w, c->omega, &e);
        size_t end = (b + 1) * c->blk < c->half_n ? (b + 1) * c->blk : c->half_n;
        for (size_t i = b * c->blk; i < end; ++i) {
            c->tw[i] = w;
            f_mul(&FR, &w, &w, c->omega);
        }
    }
}
static void fft_stage_range(size_t lo, size_t hi, void *p) {
    fft_ctx *c = (fft_ctx *)p;
    fe *a = c->a;
    for (size_t bf = lo; bf < hi; ++bf) {
        size_t start = (bf / c->half) * c->chunk, i = bf % c->half;
        fe t, u = a[start + i];
        f_mul(&FR, &t, &a[start + c->half + i], &c->tw[i * c->tchunk]);
        f_add(&FR, &a[start + i], &u, &t);
        f_sub(&FR, &a[start + c->half + i], &u, &t);
    }
}
EXPORT void zko_best_fft(u64 *data, const u64 *omega, uint32_t log_n, int threads) {
    fe *a = FE(data);
    size_t n = (size_t)1 << log_n;
    threads = resolve_threads(threads);
    for (size_t k = 0; k < n; ++k) {
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    fft_ctx c;
    c.a = a; c.omega = CFE(omega);
    c.half_n = n / 2 ? n / 2 : 1;
    c.tw = (fe *)malloc(c.half_n * sizeof(fe));
    c.blk = 1 << 12;
    /* table of omega^i, filled in parallel blocks so the baseline is not serial-bound */
    parallel_for((c.half_n + c.blk - 1) / c.blk, threads, fft_twiddle_range, &c);
    c.chunk = 2; c.tchunk = n / 2;
    for (uint32_t s = 0; s < log_n; ++s) {
        c.half = c.chunk / 2;
        parallel_for(n / 2, threads, fft_stage_range, &c);
        c.chunk *= 2;
        c.tchunk /= 2;
    }
    free(c.tw);
}

typedef struct { fe *a; const fe *s, *c1, *c2; } vec_ctx;
static void scale_range(size_t lo, size_t hi, void *p) {
    vec_ctx *c = (vec_ctx *)p;
    for (size_t i = lo; i < hi; ++i) f_mul(&FR, &c->a[i], &c->a[i], c->s);
}
static void fr_scale_all(fe *a, size_t n, const fe *s, int threads) {
    vec_ctx c = {a, s, NULL, NULL};
    parallel_for(n, threads, scale_range, &c);
}
static void fr_pow2_inv(unsigned k, fe *o) { /* (2^k)^-1 */
    fe c = {{(u64)1 << k, 0, 0, 0}}, m;
    f_to_mont(&FR, &m, &c);
    f_inv(&FR, o, &m);
}

/* EvaluationDomain::lagrange_to_coeff: ifft(a, omega_inv, k, 1/n) */
EXPORT void zko_lagrange_to_coeff(u64 *data, uint32_t k, int threads) {
    threads = resolve_threads(threads);
    fe w, wi, d;
    fr_omega(k, &w);
    f_inv(&FR, &wi, &w);
    zko_best_fft(data, wi.l, k, threads);
    fr_pow2_inv(k, &d);
    fr_scale_all(FE(data), (size_t)1 << k, &d, threads);
}
/* EvaluationDomain::coeff_to_lagrange */
EXPORT void zko_coeff_to_lagrange(u64 *data, uint32_t k, int threads) {
    fe w;
    fr_omega(k, &w);
    zko_best_fft(data, w.l, k, resolve_threads(threads));
}

/* distribute_powers_zeta: a[i] *= [1, zeta, zeta^2][i%3] (into coset) or [1, zeta^2, zeta][i%3] */
static void zeta_range(size_t lo, size_t hi, void *p) {
    vec_ctx *c = (vec_ctx *)p;
    for (size_t i = lo; i < hi; ++i) {
        size_t m = i % 3;
        if (m == 1) f_mul(&FR, &c->a[i], &c->a[i], c->c1);
        else if (m == 2) f_mul(&FR, &c->a[i], &c->a[i], c->c2);
    }
}
static void distribute_powers_zeta(fe *a, size_t n, int into_coset, int threads) {
    fe z, z2;
    f_to_mont(&FR, &z, &FR_ZETA_CANON);
    f_sqr(&FR, &z2, &z);
    vec_ctx c = {a, NULL, into_coset ? &z : &z2, into_coset ? &z2 : &z};
    parallel_for(n, threads, zeta_range, &c);
}

/* EvaluationDomain::coeff_to_extended: in = 2^k coeffs, out = 2^ek coset evaluations */
EXPORT void zko_coeff_to_extended(const u64 *in, u64 *out, uint32_t k, uint32_t ek, int threads) {
    threads = resolve_threads(threads);
    size_t n = (size_t)1 << k, en = (size_t)1 << ek;
    memcpy(out, in, n * sizeof(fe));
    memset(out + 4 * n, 0, (en - n) * sizeof(fe));
    distribute_powers_zeta(FE(out), n, 1, threads);
    fe w;
    fr_omega(ek, &w);
    zko_best_fft(out, w.l, ek, threads);
}

/* EvaluationDomain::extended_to_coeff: in place over 2^ek; caller truncates to n*(j-1) */
EXPORT void zko_extended_to_coeff(u64 *data, uint32_t k, uint32_t ek, int threads) {
    (void)k;
    threads = resolve_threads(threads);
    size_t en = (size_t)1 << ek;
    fe w, wi, d;
    fr_omega(ek, &w);
    f_inv(&FR, &wi, &w);
    zko_best_fft(data, wi.l, ek, threads);
    fr_pow2_inv(ek, &d);
    fr_scale_all(FE(data), en, &d, threads);
    distribute_powers_zeta(FE(data), en, 0, threads);
}

/* ------------------------------------------------------------------------------------------ */
/* G1: y^2 = x^3 + 3 over Fq; Jacobian (x,y,z); affine identity (0,0); Jacobian identity z = 0  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { fe x, y; } g1a;
typedef struct { fe x, y, z; } g1j;

static inline int g1a_is_identity(const g1a *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline int g1j_is_identity(const g1j *p) { return fe_is_zero(&p->z); }
static inline void g1j_set_identity(g1j *p) { memset(p, 0, sizeof(*p)); p->y = FQ.r; }

static void g1j_double(g1j *o, const g1j *p) {
    if (g1j_is_identity(p)) { *o = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    f_sqr(&FQ, &a, &p->x);
    f_sqr(&FQ, &b, &p->y);
    f_sqr(&FQ, &c, &b);
    f_add(&FQ, &t, &p->x, &b);
    f_sqr(&FQ, &t, &t);
    f_sub(&FQ, &t, &t, &a);
    f_sub(&FQ, &t, &t, &c);
    f_dbl(&FQ, &d, &t);
    f_dbl(&FQ, &e, &a);
    f_add(&FQ, &e, &e, &a);
    f_sqr(&FQ, &f, &e);
    f_dbl(&FQ, &t, &d);
    f_sub(&FQ, &x3, &f, &t);
    f_mul(&FQ, &z3, &p->y, &p->z);
    f_dbl(&FQ, &z3, &z3);
    f_sub(&FQ, &t, &d, &x3);
    f_mul(&FQ, &y3, &e, &t);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_sub(&FQ, &y3, &y3, &c);
    o->x = x3; o->y = y3; o->z = z3;
}

static void g1j_add_mixed(g1j *o, const g1j *p, const g1a *q) {
    if (g1a_is_identity(q)) { *o = *p; return; }
    if (g1j_is_identity(p)) { o->x = q->x; o->y = q->y; o->z = FQ.r; return; }
    fe z1z1, u2, s2, h, hh, i, j, r, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    f_sub(&FQ, &h, &u2, &p->x);
    f_sub(&FQ, &r, &s2, &p->y);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&r)) { g1j_double(o, p); return; }
        g1j_set_identity(o);
        return;
    }
    f_sqr(&FQ, &hh, &h);
    f_dbl(&FQ, &i, &hh);
    f_dbl(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_dbl(&FQ, &r, &r);
    f_mul(&FQ, &v, &p->x, &i);
    f_sqr(&FQ, &x3, &r);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &r, &t);
    f_mul(&FQ, &t, &p->y, &j);
    f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &h);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &hh);
    o->x = x3; o->y = y3; o->z = z3;
}

static void g1j_add(g1j *o, const g1j *p, const g1j *q) {
    if (g1j_is_identity(q)) { *o = *p; return; }
    if (g1j_is_identity(p)) { *o = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_sqr(&FQ, &z2z2, &q->z);
    f_mul(&FQ, &u1, &p->x, &z2z2);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s1, &p->y, &q->z);
    f_mul(&FQ, &s1, &s1, &z2z2);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    f_sub(&FQ, &h, &u2, &u1);
    f_sub(&FQ, &r, &s2, &s1);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&r)) { g1j_double(o, p); return; }
        g1j_set_identity(o);
        return;
    }
    f_dbl(&FQ, &i, &h);
    f_sqr(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_dbl(&FQ, &r, &r);
    f_mul(&FQ, &v, &u1, &i);
    f_sqr(&FQ, &x3, &r);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &r, &t);
    f_mul(&FQ, &t, &s1, &j);
    f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &q->z);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &z2z2);
    f_mul(&FQ, &z3, &z3, &h);
    o->x = x3; o->y = y3; o->z = z3;
}

static void g1j_to_affine(g1a *o, const g1j *p) {
    if (g1j_is_identity(p)) { memset(o, 0, sizeof(*o)); return; }
    fe zi, zi2, zi3;
    f_inv(&FQ, &zi, &p->z);
    f_sqr(&FQ, &zi2, &zi);
    f_mul(&FQ, &zi3, &zi2, &zi);
    f_mul(&FQ, &o->x, &p->x, &zi2);
    f_mul(&FQ, &o->y, &p->y, &zi3);
}

/* Curve::batch_normalize (Montgomery's trick) */
static void g1j_batch_normalize(g1a *out, const g1j *in, size_t n) {
    fe *pre = (fe *)malloc((n + 1) * sizeof(fe));
    fe acc = FQ.r;
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!g1j_is_identity(&in[i])) f_mul(&FQ, &acc, &acc, &in[i].z);
    }
    fe inv;
    f_inv(&FQ, &inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (g1j_is_identity(&in[i])) { memset(&out[i], 0, sizeof(g1a)); continue; }
        fe zi, zi2, zi3;
        f_mul(&FQ, &zi, &inv, &pre[i]);
        f_mul(&FQ, &inv, &inv, &in[i].z);
        f_sqr(&FQ, &zi2, &zi);
        f_mul(&FQ, &zi3, &zi2, &zi);
        f_mul(&FQ, &out[i].x, &in[i].x, &zi2);
        f_mul(&FQ, &out[i].y, &in[i].y, &zi3);
    }
    free(pre);
}

static void g1_generator(g1a *g) {
    fe one = {{1, 0, 0, 0}}, two = {{2, 0, 0, 0}};
    f_to_mont(&FQ, &g->x, &one);
    f_to_mont(&FQ, &g->y, &two);
}

/* double-and-add; scalar canonical (non-Montgomery) limbs */
static void g1_mul_canon(g1j *o, const g1a *p, const fe *s) {
    g1j acc;
    g1j_set_identity(&acc);
    for (int i = 255; i >= 0; --i) {
        g1j_double(&acc, &acc);
        if ((s->l[i >> 6] >> (i & 63)) & 1) g1j_add_mixed(&acc, &acc, p);
    }
    *o = acc;
}

EXPORT void zko_g1_generator(u64 *out_aff) { g1_generator((g1a *)out_aff); }
EXPORT void zko_g1_to_affine(const u64 *jac, u64 *aff) { g1j_to_affine((g1a *)aff, (const g1j *)jac); }
EXPORT int zko_g1_is_on_curve(const u64 *aff) {
    const g1a *p = (const g1a *)aff;
    if (g1a_is_identity(p)) return 1;
    fe y2, x3, three = {{3, 0, 0, 0}}, b;
    f_to_mont(&FQ, &b, &three);
    f_sqr(&FQ, &y2, &p->y);
    f_sqr(&FQ, &x3, &p->x);
    f_mul(&FQ, &x3, &x3, &p->x);
    f_add(&FQ, &x3, &x3, &b);
    return fe_eq(&y2, &x3);
}
/* [s]P, s Montgomery Fr, P affine -> affine */
EXPORT void zko_g1_mul(const u64 *base_aff, const u64 *scalar_mont, u64 *out_aff) {
    fe s;
    f_from_mont(&FR, &s, CFE(scalar_mont));
    g1j r;
    g1_mul_canon(&r, (const g1a *)base_aff, &s);
    g1j_to_affine((g1a *)out_aff, &r);
}
EXPORT void zko_g1_add_affine(const u64 *a, const u64 *b, u64 *out_aff) {
    g1j r;
    r.x = ((const g1a *)a)->x; r.y = ((const g1a *)a)->y; r.z = FQ.r;
    if (g1a_is_identity((const g1a *)a)) g1j_set_identity(&r);
    g1j_add_mixed(&r, &r, (const g1a *)b);
    g1j_to_affine((g1a *)out_aff, &r);
}
/* out[i] = [s_i]G (affine), s Montgomery Fr — ParamsKZG::setup's g[i] = [s^i]G building block */
typedef struct { const u64 *scalars; g1j *tmp; g1a g; } fbm_ctx;
static void fbm_range(size_t lo, size_t hi, void *p) {
    fbm_ctx *c = (fbm_ctx *)p;
    for (size_t i = lo; i < hi; ++i) {
        fe s;
        f_from_mont(&FR, &s, CFE(c->scalars + 4 * i));
        g1_mul_canon(&c->tmp[i], &c->g, &s);
    }
}
EXPORT void zko_g1_fixed_base_mul(const u64 *scalars_mont, size_t n, u64 *out_aff, int threads) {
    threads = resolve_threads(threads);
    fbm_ctx c;
    c.scalars = scalars_mont;
    g1_generator(&c.g);
    c.tmp = (g1j *)malloc((n ? n : 1) * sizeof(g1j));
    parallel_for(n, threads, fbm_range, &c);
    g1j_batch_normalize((g1a *)out_aff, c.tmp, n);
    free(c.tmp);
}


/* Curve::batch_normalize: n Jacobian points -> affine, identity -> (0,0) */
EXPORT void zko_g1_batch_normalize(const u64 *jac, size_t n, u64 *out_aff) {
    g1j_batch_normalize((g1a *)out_aff, (const g1j *)jac, n);
}

/* arithmetic::eval_polynomial: sum_i coeffs[i] x^i by Horner; Montgomery in, Montgomery out */
EXPORT void zko_fr_eval_polynomial(const u64 *coeffs, size_t n, const u64 *x, u64 *out) {
    fe acc;
    memset(&acc, 0, sizeof(acc));
    for (size_t i = n; i-- > 0;) {
        f_mul(&FR, &acc, &acc, CFE(x));
        f_add(&FR, &acc, &acc, CFE(coeffs + 4 * i));
    }
    *(fe *)out = acc;
}


/* arithmetic::kate_division: a has n coefficients, out receives the n-1 coefficients of a(X) / (X - b) */
EXPORT void zko_fr_kate_division(const u64 *a, size_t n, const u64 *b, u64 *out) {
    if (n < 2) return;
    fe q = *CFE(a + 4 * (n - 1));
    ((fe *)out)[n - 2] = q;
    for (size_t i = n - 2; i >= 1; --i) {
        f_mul(&FR, &q, &q, CFE(b));
        f_add(&FR, &q, &q, CFE(a + 4 * i));
        ((fe *)out)[i - 1] = q;
    }
}

/* ff::BatchInvert: element-wise inverse in place, zeros untouched (each element inverted on its own here) */
EXPORT void zko_fr_batch_invert(u64 *a, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        fe *v = (fe *)(a + 4 * i);
        if (!fe_is_zero(v)) f_inv(&FR, v, v);
    }
}


/* plonk::lookup::prover::permute_expression_pair (halo2-axiom plonk/lookup/prover.rs, un-vendored; restated from the published
 * halo2 algorithm):
 *   permuted_input = input[..usable_rows] sorted (Fr's Ord = canonical integer value)
 *   leftover_table_map: BTreeMap value -> count over table[..usable_rows]
 *   for every row: first occurrence of its value in permuted_input -> permuted_table[row] = value, one copy leaves the map
 *                  (absent: Err(ConstraintSystemFailure), returns -1 here); otherwise the row is pushed on repeated_input_rows
 *   for (value, count) in the map, ascending: count times permuted_table[repeated_input_rows.pop()] = value
 * Outputs hold usable_rows elements each (the caller appends the blinding rows). */
typedef struct { fe canon; fe mont; } pe_item;
static int pe_cmp(const void *x, const void *y) {
    const pe_item *a = (const pe_item *)x, *b = (const pe_item *)y;
    for (int i = 3; i >= 0; --i) {
        if (a->canon.l[i] != b->canon.l[i]) return a->canon.l[i] < b->canon.l[i] ? -1 : 1;
    }
    return 0;
}
EXPORT int zko_permute_expression_pair(const u64 *input, const u64 *table, size_t usable_rows, u64 *out_input, u64 *out_table) {
    const size_t u = usable_rows;
    if (u == 0) return 0;
    pe_item *a = (pe_item *)malloc(sizeof(pe_item) * u), *t = (pe_item *)malloc(sizeof(pe_item) * u);
    size_t *repeated = (size_t *)malloc(sizeof(size_t) * u);
    unsigned char *taken = (unsigned char *)calloc(u, 1);
    for (size_t i = 0; i < u; ++i) {
        a[i].mont = *CFE(input + 4 * i); f_from_mont(&FR, &a[i].canon, &a[i].mont);
        t[i].mont = *CFE(table + 4 * i); f_from_mont(&FR, &t[i].canon, &t[i].mont);
    }
    qsort(a, u, sizeof(pe_item), pe_cmp);
    qsort(t, u, sizeof(pe_item), pe_cmp);   /* the BTreeMap's iteration order; equal keys are adjacent = its counts */
    size_t nrep = 0;
    int rc = 0;
    for (size_t row = 0; row < u && rc == 0; ++row) {
        *FE(out_input + 4 * row) = a[row].mont;
        if (row == 0 || pe_cmp(&a[row], &a[row - 1]) != 0) {
            *FE(out_table + 4 * row) = a[row].mont;
            /* remove one instance from the map: the first not yet taken element equal to the value */
            size_t lo = 0, hi = u;
            while (lo < hi) { size_t mid = (lo + hi) / 2; if (pe_cmp(&t[mid], &a[row]) < 0) lo = mid + 1; else hi = mid; }
            if (lo >= u || pe_cmp(&t[lo], &a[row]) != 0) rc = -1;
            else taken[lo] = 1;   /* distinct input values hit distinct table values, so one copy each */
        } else {
            repeated[nrep++] = row;
        }
    }
    for (size_t i = 0; i < u && rc == 0; ++i) {
        if (taken[i]) continue;
        if (nrep == 0) { rc = -2; break; }
        *FE(out_table + 4 * repeated[--nrep]) = t[i].mont;
    }
    if (rc == 0 && nrep != 0) rc = -2;
    free(a); free(t); free(repeated); free(taken);
    return rc;
}

/* plonk::evaluation::GraphEvaluator::evaluate for every row idx of the extended domain, as evaluate_h calls it:
 *   values[idx] = graph.evaluate(data, fixed, advice, instance, challenges, beta, gamma, theta, y, &values[idx], idx, rot_scale, isize)
 * calcs: ncalc x 11 u32 = { op, target, a.kind, a.index, a.rotation, b.(3), c.(3) }  (the layout of zkb_calculation)
 *   op: 0 Add 1 Sub 2 Mul 3 Square 4 Double 5 Negate 6 Store 7 one Horner step (target = a * b + c)
 *   kind: 0 Constant 1 Intermediate 2 Fixed 3 Advice 4 Instance 5 Challenge 6 Beta 7 Gamma 8 Theta 9 Y 10 PreviousValue
 * Every intermediate starts at zero (upstream's EvaluationData::intermediates), one array per row; the row's result is the
 * target of the last calculation, or zero for an empty graph.  Returns -1 on a malformed graph. */
typedef struct {
    const uint32_t *calcs; size_t ncalc; uint32_t nint;
    const fe *constants; const int32_t *rotations;
    const fe *const *fixed; const fe *const *advice; const fe *const *instance;
    const fe *challenges; const fe *beta, *gamma, *theta, *y;
    int64_t rot_scale; fe *values; int64_t isize; int bad;
} graph_ctx;
static const fe *graph_get(const graph_ctx *g, const uint32_t *src, const fe *inter, const fe *prev, int64_t idx) {
    const uint32_t kind = src[0], index = src[1];
    int64_t row = 0;
    if (kind >= 2 && kind <= 4) {   /* get_rotation_idx: (idx + rot * rot_scale).rem_euclid(isize) */
        row = (idx + (int64_t)g->rotations[src[2]] * g->rot_scale) % g->isize;
        if (row < 0) row += g->isize;
    }
    switch (kind) {
        case 0: return &g->constants[index];
        case 1: return &inter[index];
        case 2: return &g->fixed[index][row];
        case 3: return &g->advice[index][row];
        case 4: return &g->instance[index][row];
        case 5: return &g->challenges[index];
        case 6: return g->beta;
        case 7: return g->gamma;
        case 8: return g->theta;
        case 9: return g->y;
        case 10: return prev;
        default: return NULL;
    }
}
static void graph_range(size_t lo, size_t hi, void *p) {
    graph_ctx *g = (graph_ctx *)p;
    fe *inter = (fe *)calloc(g->nint ? g->nint : 1, sizeof(fe));
    for (size_t idx = lo; idx < hi; ++idx) {
        memset(inter, 0, (g->nint ? g->nint : 1) * sizeof(fe));
        const fe prev = g->values[idx];
        for (size_t i = 0; i < g->ncalc; ++i) {
            const uint32_t *c = g->calcs + 11 * i;
            const fe *a = graph_get(g, c + 2, inter, &prev, (int64_t)idx);
            const fe *b = graph_get(g, c + 5, inter, &prev, (int64_t)idx);
            const fe *d = graph_get(g, c + 8, inter, &prev, (int64_t)idx);
            fe r;
            switch (c[0]) {
                case 0: f_add(&FR, &r, a, b); break;
                case 1: f_sub(&FR, &r, a, b); break;
                case 2: f_mul(&FR, &r, a, b); break;
                case 3: f_sqr(&FR, &r, a); break;
                case 4: f_dbl(&FR, &r, a); break;
                case 5: f_neg(&FR, &r, a); break;
                case 6: r = *a; break;
                case 7: f_mul(&FR, &r, a, b); f_add(&FR, &r, &r, d); break;
                default: g->bad = 1; r = *a; break;
            }
            inter[c[1]] = r;
        }
        if (g->ncalc) g->values[idx] = inter[g->calcs[11 * (g->ncalc - 1) + 1]];
        else memset(&g->values[idx], 0, sizeof(fe));
    }
    free(inter);
}
EXPORT int zko_graph_evaluate(const uint32_t *calcs, size_t ncalc, uint32_t num_intermediates, const u64 *constants,
                              const int32_t *rotations, const u64 *const *fixed, const u64 *const *advice,
                              const u64 *const *instance, const u64 *challenges, const u64 *beta, const u64 *gamma,
                              const u64 *theta, const u64 *y, int32_t rot_scale, u64 *values, size_t isize, int threads) {
    for (size_t i = 0; i < ncalc; ++i) {
        if (calcs[11 * i] > 7 || calcs[11 * i + 1] >= num_intermediates) return -1;
        for (int k = 0; k < 3; ++k)
            if (calcs[11 * i + 2 + 3 * k] > 10) return -1;
    }
    graph_ctx g = { calcs, ncalc, num_intermediates, (const fe *)constants, rotations, (const fe *const *)fixed,
                    (const fe *const *)advice, (const fe *const *)instance, (const fe *)challenges, (const fe *)beta,
                    (const fe *)gamma, (const fe *)theta, (const fe *)y, rot_scale, (fe *)values, (int64_t)isize, 0 };
    parallel_for(isize, resolve_threads(threads), graph_range, &g);
    return g.bad ? -1 : 0;
}


/* best_fft::<Fr, G1> by the definition: out[i] = sum_j [omega^(i j)] P_j   (O(n^2) scalar multiplications; tiny n only) */
EXPORT void zko_g1_fft_naive(const u64 *points_aff, size_t n, const u64 *omega_mont, u64 *out_aff) {
    const g1a *P = (const g1a *)points_aff;
    g1j *res = (g1j *)malloc((n ? n : 1) * sizeof(g1j));
    fe wi = FR.r; /* omega^i */
    for (size_t i = 0; i < n; ++i) {
        g1j acc;
        g1j_set_identity(&acc);
        fe wij = FR.r; /* omega^(i j) */
        for (size_t j = 0; j < n; ++j) {
            fe sc;
            f_from_mont(&FR, &sc, &wij);
            g1j t;
            g1_mul_canon(&t, &P[j], &sc);
            g1j_add(&acc, &acc, &t);
            f_mul(&FR, &wij, &wij, &wi);
        }
        res[i] = acc;
        f_mul(&FR, &wi, &wi, CFE(omega_mont));
    }
    g1j_batch_normalize((g1a *)out_aff, res, n);
    free(res);
}

/* ParamsKZG::setup, G1 side (halo2-axiom poly/kzg/commitment.rs; reference call sites voter_circuit.rs:60,
 * state_transition_circuit.rs:64): every scalar by the definition (one exponentiation / inversion per index), every
 * point by double-and-add.  Returns 1 if s is an n-th root of unity (upstream panics). */
EXPORT int zko_kzg_setup(unsigned k, const u64 *s_mont, u64 *g_out, u64 *gl_out, int threads) {
    size_t n = (size_t)1 << k;
    fe s = *CFE(s_mont);
    fe *sc = (fe *)malloc(n * sizeof(fe));
    if (g_out) {
        fe p = FR.r;
        for (size_t i = 0; i < n; ++i) { sc[i] = p; f_mul(&FR, &p, &p, &s); }
        zko_g1_fixed_base_mul((const u64 *)sc, n, g_out, threads);
    }
    if (gl_out) {
        fe omega, sn = s, nf = FR.r, ninv, c, one = FR.r;
        fr_omega(k, &omega);
        for (unsigned i = 0; i < k; ++i) { f_sqr(&FR, &sn, &sn); f_dbl(&FR, &nf, &nf); }
        f_inv(&FR, &ninv, &nf);
        f_sub(&FR, &c, &sn, &one);
        f_mul(&FR, &c, &c, &ninv);
        fe w = FR.r;
        for (size_t i = 0; i < n; ++i) {
            fe d, dinv;
            f_sub(&FR, &d, &s, &w);
            if (fe_is_zero(&d)) { free(sc); return 1; }
            f_inv(&FR, &dinv, &d);
            f_mul(&FR, &sc[i], &w, &c);
            f_mul(&FR, &sc[i], &sc[i], &dinv);
            f_mul(&FR, &w, &w, &omega);
        }
        zko_g1_fixed_base_mul((const u64 *)sc, n, gl_out, threads);
    }
    free(sc);
    return 0;
}

/* naive sum of double-and-add products (ground truth for tiny n) */
EXPORT void zko_msm_naive(const u64 *scalars_mont, const u64 *bases_aff, size_t n, u64 *out_jac) {
    g1j acc;
    g1j_set_identity(&acc);
    for (size_t i = 0; i < n; ++i) {
        fe s;
        f_from_mont(&FR, &s, CFE(scalars_mont + 4 * i));
        g1j t;
        g1_mul_canon(&t, (const g1a *)bases_aff + i, &s);
        g1j_add(&acc, &acc, &t);
    }
    *(g1j *)out_jac = acc;
}

/* halo2_proofs::arithmetic::multiexp_serial: coeffs are canonical (to_repr) limbs */
static void multiexp_serial(const fe *coeffs, const g1a *bases, size_t n, g1j *acc) {
    unsigned c = n < 4 ? 1 : n < 32 ? 3 : (unsigned)ceil(log((double)n));
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    g1j *buckets = (g1j *)malloc(nb * sizeof(g1j));
    for (unsigned seg = segments; seg-- > 0;) {
        for (unsigned i = 0; i < c; ++i) g1j_double(acc, acc);
        for (size_t b = 0; b < nb; ++b) g1j_set_identity(&buckets[b]);
        unsigned bit = seg * c;
        for (size_t i = 0; i < n; ++i) {
            if (bit >= 256) continue;
            unsigned w = bit >> 6, sh = bit & 63;
            u64 d = coeffs[i].l[w] >> sh;
            if (sh + c > 64 && w < 3) d |= coeffs[i].l[w + 1] << (64 - sh);
            d &= ((u64)1 << c) - 1;
            if (d) g1j_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        }
        g1j running;
        g1j_set_identity(&running);
        for (size_t b = nb; b-- > 0;) {
            g1j_add(&running, &running, &buckets[b]);
            g1j_add(acc, acc, &running);
        }
    }
    free(buckets);
}

/* halo2_proofs::arithmetic::best_multiexp; result Jacobian, normalised so z = R (or identity) */
typedef struct { const u64 *scalars; fe *coeffs; const g1a *bases; g1j *res; size_t n, chunk; } me_ctx;
static void me_repr_range(size_t lo, size_t hi, void *p) {
    me_ctx *c = (me_ctx *)p;
    for (size_t i = lo; i < hi; ++i) f_from_mont(&FR, &c->coeffs[i], CFE(c->scalars + 4 * i));
}
static void me_chunk_range(size_t lo, size_t hi, void *p) {
    me_ctx *c = (me_ctx *)p;
    for (size_t ci = lo; ci < hi; ++ci) {
        size_t a = ci * c->chunk, b = a + c->chunk < c->n ? a + c->chunk : c->n;
        g1j_set_identity(&c->res[ci]);
        multiexp_serial(c->coeffs + a, c->bases + a, b - a, &c->res[ci]);
    }
}
EXPORT void zko_best_multiexp(const u64 *scalars_mont, const u64 *bases_aff, size_t n, int threads,
                              u64 *out_jac) {
    threads = resolve_threads(threads);
    me_ctx c;
    c.scalars = scalars_mont;
    c.coeffs = (fe *)malloc((n ? n : 1) * sizeof(fe));
    c.bases = (const g1a *)bases_aff;
    c.n = n;
    parallel_for(n, threads, me_repr_range, &c);
    g1j total;
    g1j_set_identity(&total);
    if (n > (size_t)threads) {
        c.chunk = n / threads;
        size_t nchunks = (n + c.chunk - 1) / c.chunk;
        c.res = (g1j *)malloc(nchunks * sizeof(g1j));
        parallel_for(nchunks, (int)nchunks, me_chunk_range, &c); /* one OS thread per chunk, like rayon::scope */
        for (size_t ci = 0; ci < nchunks; ++ci) g1j_add(&total, &total, &c.res[ci]);
        free(c.res);
    } else {
        multiexp_serial(c.coeffs, c.bases, n, &total);
    }
    free(c.coeffs);
    g1a aff;
    g1j_to_affine(&aff, &total);
    g1j *o = (g1j *)out_jac;
    if (g1j_is_identity(&total)) { g1j_set_identity(o); return; }
    o->x = aff.x; o->y = aff.y; o->z = FQ.r;
}

EXPORT int zko_num_threads(void) { return resolve_threads(0); }
