// ntt.cuh — Fr NTT passes (device code written as per-thread phase functions).
//
// Replaces halo2_proofs::arithmetic::best_fft: natural order in, natural order out,
// out[i] = sum_j a[j] * omega^(i*j).  The field is exact, so any correct factorisation gives bit-identical
// output; this one is a mixed-radix Cooley-Tukey ("four-step" generalised to P passes):
//
//   N = R_1 * R_2 * ... * R_P          (R_p = 2^lr[p], <= 1024)
//   n = sum_p n_p * S_p ,  S_p = prod_{q>p} R_q      (n_1 most significant input digit)
//   k = sum_p k_p * Q_p ,  Q_p = prod_{q<p} R_q      (k_1 least significant output digit)
//
// Pass p transforms digit n_p -> k_p (an R_p-point NTT per (hi, lo) pair) in place: k_p is stored where n_p
// was (stride S_p).  Before the transform each element is multiplied by the inter-pass twiddle
//   omega_N ^ ( n_p * K_{p-1} * N / Q_{p+1} ),    K_{p-1} = sum_{q<p} k_q * Q_q ,
// read from a two-level table (hi/lo split of the exponent).  The last pass writes out of place to the
// digit-reversed address k = K_{P-1} + Q_P * k_P, so the result lands in natural order with no separate
// transpose pass.  One CTA stages R*T elements (T adjacent columns of the strided dimension, so every global
// row it touches is T*32 contiguous bytes) in shared memory, runs radix-8 DIF rounds with 8 elements per thread
// in registers, and reads the digit-reversed shared-memory position on the way out.
//
// Fusions: first pass can zero-extend (in_len < N) and scale by c[i mod 3] (coset shift of
// coeff_to_extended); last pass can scale by c[i mod 3] (1/N and the inverse coset shift of extended_to_coeff,
// or the plain 1/N of lagrange_to_coeff).
//
// Arithmetic inside and between the passes is LAZY: values are only < 2r (congruent mod r), products skip their final
// conditional subtraction and add/sub work mod 2r (field.cuh); the last pass stores canonical values.
//
// Every phase is a __host__ __device__ function of (cta, tid): the kernels in ntt.cu call them with
// __syncthreads() in between, and the CPU emulator (hostemu.cu, test infrastructure) runs the very same code
// thread by thread so index logic is checked without a GPU.
#pragma once
#include "field.cuh"

namespace zkb {

constexpr int NTT_MAX_PASSES = 4;
constexpr int NTT_MAX_RANKS = 8;

struct NttPassArgs {
    const uint4* src;     // column 0 of the input  (element = 2 x uint4)
    uint4* dst;           // column 0 of the output
    uint64_t src_col_stride;  // elements between columns in src
    uint64_t dst_col_stride;
    uint32_t log_n;
    uint32_t npass;           // P
    uint32_t pass;            // p (0-based)
    uint32_t lr[NTT_MAX_PASSES];  // log2 R_q
    uint32_t ncols;           // columns of the batch: block b of the 1-D grid is (cta b / ncols, column b % ncols), so the
                              // columns of one tile run back to back and share its twiddle-table lines in L2
    uint32_t log_t;           // log2 T (tile width)
    uint32_t is_final;        // last pass: transposed (digit-reversed) store
    // twiddles (device pointers)
    const uint4* tw_r;        // omega_R^t, t < R for this pass' R
    const uint4* tw_hi;       // omega_N^(t << tw_h)
    const uint4* tw_lo;       // omega_N^t, t < 2^tw_h
    uint32_t tw_h;
    const uint4* tw_pass;     // optional: the inter-pass twiddles of THIS pass, omega_N^((r*K) << shift) at [K * R + r]
                              // (Q_p * R_p entries): one lookup and one product instead of the hi/lo pair
    // fusions
    uint64_t in_len;          // first pass: elements >= in_len read as zero (in_len == N otherwise)
    uint32_t in_scale_on;     // first pass: multiply by in_scale[i % 3]
    uint32_t out_scale_on;    // last pass: multiply by out_scale[i % 3]
    uint32_t in_scale[3][8];
    uint32_t out_scale[3][8];
    // distributed mode (one NTT sharded over 2^dist_log_g GPUs, rank r owns the contiguous slice
    // [r * 2^dist_log_slice, (r+1) * 2^dist_log_slice) of every buffer): loads and stores address element gi through
    // peer_src/peer_dst[gi >> dist_log_slice] — peer HBM mapped over NVLink (CUDA IPC).  src/dst are unused.
    uint32_t dist_log_g;      // 0 = single GPU
    uint32_t dist_rank;
    uint32_t dist_log_slice;
    const uint4* peer_src[NTT_MAX_RANKS];
    uint4* peer_dst[NTT_MAX_RANKS];
    const uint32_t* dist_abort;  // status word of the device-side barriers: non-zero = a peer never arrived, skip the pass
};

// shared memory is split in two planes (low / high 16 bytes) so a warp's 128-bit accesses are conflict-free
ZKB_HD Fr sm_load(const uint4* lo, const uint4* hi, uint32_t i) { return fr_from_u4(lo[i], hi[i]); }
ZKB_HD void sm_store(uint4* lo, uint4* hi, uint32_t i, const Fr& v) {
    lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ---- element addressing (single GPU: column-major batch; distributed: owner slice over peer memory) -----------
ZKB_HD const uint4* ntt_src_elem(const NttPassArgs& a, uint32_t col, uint64_t gi) {
    if (a.dist_log_g) return a.peer_src[gi >> a.dist_log_slice] + 2 * (gi & ((1ull << a.dist_log_slice) - 1));
    return a.src + 2 * (a.src_col_stride * col + gi);
}
ZKB_HD uint4* ntt_dst_elem(const NttPassArgs& a, uint32_t col, uint64_t go) {
    if (a.dist_log_g) return a.peer_dst[go >> a.dist_log_slice] + 2 * (go & ((1ull << a.dist_log_slice) - 1));
    return a.dst + 2 * (a.dst_col_stride * col + go);
}

// Distributed mode: CTA `local` of this rank's grid -> CTA index of the single-GPU enumeration.
//   pass 0        : any split of the lo tiles works (every CTA reads all ranks); contiguous share.
//   middle passes : after pass 0 rank r holds k_1 in [r R_1/G, (r+1) R_1/G) — the leading part of the stored hi
//                   index, i.e. a contiguous CTA range: these passes touch local memory only.
//   final pass    : CTAs enumerate K = k_1 + R_1 * rest (k_1 least significant); the rank's share is the K with
//                   k_1 in its block.  Its stores scatter to every rank's output slice (natural order).
ZKB_HD uint64_t ntt_dist_cta(const NttPassArgs& a, uint64_t local) {
    if (!a.dist_log_g) return local;
    if (!a.is_final) {
        const uint64_t per_rank = ((1ull << a.log_n) >> (a.lr[a.pass] + a.log_t)) >> a.dist_log_g;
        return (uint64_t)a.dist_rank * per_rank + local;
    }
    const uint32_t log_b = a.lr[0] - a.dist_log_g - a.log_t;  // T-wide k_1 tiles per rank
    const uint64_t k1_tile = local & ((1ull << log_b) - 1), rest = local >> log_b;
    const uint64_t K0 = ((uint64_t)a.dist_rank << (a.lr[0] - a.dist_log_g)) + (k1_tile << a.log_t) + (rest << a.lr[0]);
    return K0 >> a.log_t;
}

// ---- geometry helpers ---------------------------------------------------------------------------------
// stored hi index (k_1 most significant, radices R_1..R_{p-1}) -> K = sum k_q Q_q (k_1 least significant)
ZKB_HD uint64_t ntt_hi_to_K(const NttPassArgs& a, uint64_t hi) {
    uint64_t K = 0;
    uint32_t shift = 0;
    for (uint32_t q = 0; q < a.pass; ++q) shift += a.lr[q];
    for (int q = (int)a.pass - 1; q >= 0; --q) {
        shift -= a.lr[q];
        uint64_t d = hi & ((1ull << a.lr[q]) - 1);
        hi >>= a.lr[q];
        K |= d << shift;
    }
    return K;
}
ZKB_HD uint64_t ntt_K_to_hi(const NttPassArgs& a, uint64_t K) {
    uint64_t hi = 0;
    for (uint32_t q = 0; q < a.pass; ++q) {
        uint64_t d = K & ((1ull << a.lr[q]) - 1);
        K >>= a.lr[q];
        hi = (hi << a.lr[q]) | d;
    }
    return hi;
}

// position of output k inside the R-point block after DIF rounds of radix 8,8,..,rem
template <int LOGR>
ZKB_HD uint32_t ntt_dif_pos(uint32_t k) {
    uint32_t pos = 0;
    int left = LOGR;
    while (left > 0) {
        int t = left >= 3 ? 3 : left;
        uint32_t d = k & ((1u << t) - 1);
        k >>= t;
        left -= t;
        pos |= d << left;
    }
    return pos;
}

// Strided layout with fewer than 8 columns per row (R >= 256: T = 1024 / R): a row covers only T/8 of the 32 banks, so the 8/T
// rows one quarter-warp touches in a 128-bit access must fall into different bank groups.  Which row bits vary inside a
// quarter-warp depends on the phase: the low bits in the load and the first DIF rounds, bits [L+t ..] in the rounds whose
// sub-block is shorter than the group of adjacent threads, the top digit in the digit-reversed store.  XOR-ing those bits
// into the low 3 - log_t row bits makes every phase conflict-free (a bijection on the rows: only low bits change, from
// higher bits).  Rounds are 3,3,2 (R = 2^8), 3,3,3 (2^9), 3,3,3,1 (2^10).
template <int LOGR>
ZKB_HD uint32_t ntt_row_swizzle(uint32_t r) {
    if (LOGR == 8) return r ^ (((r >> 2) ^ (r >> 5)) & 1u);                                  // T = 4: one bit
    if (LOGR == 9) return r ^ (((r >> 3) ^ (r >> 6)) & 3u);                                  // T = 2: two bits (r3,r4 / r6,r7)
    if (LOGR == 10) return r ^ (((r >> 3) & 1u) | (((r >> 4) & 3u) << 1)) ^ ((r >> 7) & 7u);  // T = 1: three bits
    return r;
}

// Column pitch of the transposed (final-pass) layout: R + 1 for T >= 8 (8 adjacent threads = 8 adjacent columns land 1 slot
// apart mod 8); for T < 8 the T columns of a row are spread 8/T bank groups apart and the swizzled row supplies the rest.
template <int LOGR>
ZKB_HD uint32_t ntt_final_pitch(uint32_t log_t) {
    constexpr uint32_t R = 1u << LOGR;
    return (LOGR >= 8 && log_t + LOGR == 10) ? R + ((8u >> log_t) & 7u) : R + 1;
}

// shared-memory index of (row r, column c)
template <int LOGR>
ZKB_HD uint32_t ntt_sm_index(const NttPassArgs& a, uint32_t r, uint32_t c) {
    if (LOGR >= 8 && a.log_t + LOGR == 10) r = ntt_row_swizzle<LOGR>(r);
    if (a.is_final) return c * ntt_final_pitch<LOGR>(a.log_t) + r;
    return (r << a.log_t) + c;
}
template <int LOGR>
ZKB_HD uint32_t ntt_sm_plane(const NttPassArgs& a) {  // uint4 elements per plane
    constexpr uint32_t R = 1u << LOGR;
    return a.is_final ? (ntt_final_pitch<LOGR>(a.log_t) << a.log_t) : (R << a.log_t);
}

// entry t = K * R + r of pass p's twiddle table: omega_N^((r*K) << shift), from the two-level power tables
ZKB_HD Fr ntt_pass_twiddle(const uint4* tw_hi, const uint4* tw_lo, uint32_t tw_h, uint32_t log_r, uint32_t shift, uint64_t t) {
    const uint64_t K = t >> log_r, r = t & ((1ull << log_r) - 1);
    const uint64_t e = (r * K) << shift;
    const uint64_t eh = e >> tw_h, el = e & ((1ull << tw_h) - 1);
    Fr v = fr_load2(tw_lo, el);
    if (eh) v = fp_mul(v, fr_load2(tw_hi, eh));
    return v;
}

// ---- phase 1: global -> shared, with zero-extension, coset scale and inter-pass twiddle ------------------
template <int LOGR>
ZKB_HD void ntt_phase_load(const NttPassArgs& a, uint4* sm, uint32_t tid, uint32_t nthreads, uint64_t cta,
                           uint32_t col) {
    constexpr uint32_t R = 1u << LOGR;
    const uint32_t T = 1u << a.log_t;
    const uint32_t E = R << a.log_t;
    uint4* lo = sm;
    uint4* hi = sm + ntt_sm_plane<LOGR>(a);

    uint32_t log_stride = 0;  // log2 S_p
    for (uint32_t q = a.pass + 1; q < a.npass; ++q) log_stride += a.lr[q];
    uint32_t log_q_next = 0;  // log2 Q_{p+1}
    for (uint32_t q = 0; q <= a.pass; ++q) log_q_next += a.lr[q];
    const uint32_t tw_shift = a.log_n - log_q_next;

    uint64_t base, K0;
    if (!a.is_final) {
        uint64_t tiles_per_hi = (1ull << log_stride) >> a.log_t;
        uint64_t hidx = cta / tiles_per_hi;
        uint64_t lo0 = (cta % tiles_per_hi) << a.log_t;
        base = ((hidx << LOGR) << log_stride) + lo0;
        K0 = ntt_hi_to_K(a, hidx);
    } else {
        K0 = cta << a.log_t;
        base = 0;
    }

    for (uint32_t q = tid; q < E; q += nthreads) {
        uint32_t r, c;
        uint64_t gi, K;
        if (!a.is_final) {
            r = q >> a.log_t; c = q & (T - 1);
            gi = base + ((uint64_t)r << log_stride) + c;
            K = K0;
        } else {
            c = q >> LOGR; r = q & (R - 1);
            K = K0 + c;
            gi = (ntt_K_to_hi(a, K) << LOGR) + r;
        }
        Fr v;
        if (gi < a.in_len) {
            v = fr_load2(ntt_src_elem(a, col, gi), 0);
            if (a.in_scale_on) {
                uint32_t m = (uint32_t)(gi % 3);
                if (m) v = fp_mul_lazy(v, fr_from_words(a.in_scale[m]));
            }
        } else {
            v = Fr::zero();
        }
        if (a.pass > 0) {
            if (a.tw_pass) {
                if (r && K) v = fp_mul_lazy(v, fr_load2(a.tw_pass, (K << LOGR) + r));
            } else {
                uint64_t e = ((uint64_t)r * K) << tw_shift;  // < N
                uint64_t eh = e >> a.tw_h, el = e & ((1ull << a.tw_h) - 1);
                if (eh) v = fp_mul_lazy(v, fr_load2(a.tw_hi, eh));
                if (el) v = fp_mul_lazy(v, fr_load2(a.tw_lo, el));
            }
        }
        sm_store(lo, hi, ntt_sm_index<LOGR>(a, r, c), v);
    }
}

// ---- phase 2: one DIF round of radix 2^t on blocks of size L = 2^log_l -------------------------------------
// group g handles inputs pos_j = b*L + i + j*(L/rho); output p goes to b*L + p*(L/rho) + i times omega_L^(i*p)
template <int LOGR>
ZKB_HD void ntt_phase_round(const NttPassArgs& a, uint4* sm, uint32_t tid, uint32_t nthreads, uint32_t log_l,
                            uint32_t t) {
    constexpr uint32_t R = 1u << LOGR;
    const uint32_t T = 1u << a.log_t;
    uint4* lo = sm;
    uint4* hi = sm + ntt_sm_plane<LOGR>(a);
    const uint32_t log_sub = log_l - t;             // log2 (L / rho)
    const uint32_t groups = (R >> t) << a.log_t;    // (R/rho) * T
    const uint32_t tw_step = LOGR - log_l;          // omega_L = omega_R^(R/L)

    for (uint32_t g = tid; g < groups; g += nthreads) {
        // column fastest in both layouts: 8 adjacent threads touch 8 adjacent columns, i.e. adjacent uint4 slots
        // (strided layout) or slots R+1 apart (final layout, R+1 = 1 mod 8): conflict-free in every round
        const uint32_t u = g >> a.log_t, c = g & (T - 1);
        uint32_t b = u >> log_sub, i = u & ((1u << log_sub) - 1);
        uint32_t p0 = (b << log_l) + i;
        if (t == 3) {
            Fr x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = sm_load(lo, hi, ntt_sm_index<LOGR>(a, p0 + ((uint32_t)j << log_sub), c));
            const Fr w1 = fr_load2(a.tw_r, R / 8), w2 = fr_load2(a.tw_r, R / 4), w3 = fr_load2(a.tw_r, 3 * (R / 8));
            Fr a0 = fp_add_lazy(x[0], x[4]), a1 = fp_add_lazy(x[1], x[5]), a2 = fp_add_lazy(x[2], x[6]), a3 = fp_add_lazy(x[3], x[7]);
            Fr b0 = fp_sub_lazy(x[0], x[4]);
            Fr b1 = fp_mul_lazy(fp_sub_lazy(x[1], x[5]), w1);
            Fr b2 = fp_mul_lazy(fp_sub_lazy(x[2], x[6]), w2);
            Fr b3 = fp_mul_lazy(fp_sub_lazy(x[3], x[7]), w3);
            Fr c0 = fp_add_lazy(a0, a2), c1 = fp_add_lazy(a1, a3), d0 = fp_sub_lazy(a0, a2), d1 = fp_mul_lazy(fp_sub_lazy(a1, a3), w2);
            Fr e0 = fp_add_lazy(b0, b2), e1 = fp_add_lazy(b1, b3), f0 = fp_sub_lazy(b0, b2), f1 = fp_mul_lazy(fp_sub_lazy(b1, b3), w2);
            x[0] = fp_add_lazy(c0, c1); x[4] = fp_sub_lazy(c0, c1);
            x[2] = fp_add_lazy(d0, d1); x[6] = fp_sub_lazy(d0, d1);
            x[1] = fp_add_lazy(e0, e1); x[5] = fp_sub_lazy(e0, e1);
            x[3] = fp_add_lazy(f0, f1); x[7] = fp_sub_lazy(f0, f1);
            if (log_sub > 0 && i > 0) {
#pragma unroll
                for (int p = 1; p < 8; ++p) x[p] = fp_mul_lazy(x[p], fr_load2(a.tw_r, (uint64_t)(i * p) << tw_step));
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) sm_store(lo, hi, ntt_sm_index<LOGR>(a, p0 + ((uint32_t)p << log_sub), c), x[p]);
        } else if (t == 2) {
            Fr x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = sm_load(lo, hi, ntt_sm_index<LOGR>(a, p0 + ((uint32_t)j << log_sub), c));
            const Fr w4 = fr_load2(a.tw_r, R / 4);
            Fr a0 = fp_add_lazy(x[0], x[2]), a1 = fp_add_lazy(x[1], x[3]), d0 = fp_sub_lazy(x[0], x[2]);
            Fr d1 = fp_mul_lazy(fp_sub_lazy(x[1], x[3]), w4);
            x[0] = fp_add_lazy(a0, a1); x[2] = fp_sub_lazy(a0, a1);
            x[1] = fp_add_lazy(d0, d1); x[3] = fp_sub_lazy(d0, d1);
            if (log_sub > 0 && i > 0) {
#pragma unroll
                for (int p = 1; p < 4; ++p) x[p] = fp_mul_lazy(x[p], fr_load2(a.tw_r, (uint64_t)(i * p) << tw_step));
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) sm_store(lo, hi, ntt_sm_index<LOGR>(a, p0 + ((uint32_t)p << log_sub), c), x[p]);
        } else {
            Fr x0 = sm_load(lo, hi, ntt_sm_index<LOGR>(a, p0, c));
            Fr x1 = sm_load(lo, hi, ntt_sm_index<LOGR>(a, p0 + (1u << log_sub), c));
            Fr y0 = fp_add_lazy(x0, x1), y1 = fp_sub_lazy(x0, x1);
            if (log_sub > 0 && i > 0) y1 = fp_mul_lazy(y1, fr_load2(a.tw_r, (uint64_t)i << tw_step));
            sm_store(lo, hi, ntt_sm_index<LOGR>(a, p0, c), y0);
            sm_store(lo, hi, ntt_sm_index<LOGR>(a, p0 + (1u << log_sub), c), y1);
        }
    }
}

// ---- phase 3: shared -> global (in place for inner passes, digit-reversed address for the last) ------------
template <int LOGR>
ZKB_HD void ntt_phase_store(const NttPassArgs& a, const uint4* sm, uint32_t tid, uint32_t nthreads, uint64_t cta,
                            uint32_t col) {
    constexpr uint32_t R = 1u << LOGR;
    const uint32_t T = 1u << a.log_t;
    const uint32_t E = R << a.log_t;
    const uint4* lo = sm;
    const uint4* hi = sm + ntt_sm_plane<LOGR>(a);

    uint32_t log_stride = 0;
    for (uint32_t q = a.pass + 1; q < a.npass; ++q) log_stride += a.lr[q];
    uint32_t log_q = 0;  // log2 Q_p
    for (uint32_t q = 0; q < a.pass; ++q) log_q += a.lr[q];

    uint64_t base = 0, K0 = 0;
    if (!a.is_final) {
        uint64_t tiles_per_hi = (1ull << log_stride) >> a.log_t;
        uint64_t hidx = cta / tiles_per_hi;
        uint64_t lo0 = (cta % tiles_per_hi) << a.log_t;
        base = ((hidx << LOGR) << log_stride) + lo0;
    } else {
        K0 = cta << a.log_t;
    }
    for (uint32_t q = tid; q < E; q += nthreads) {
        uint32_t k = q >> a.log_t, c = q & (T - 1);
        Fr v = sm_load(lo, hi, ntt_sm_index<LOGR>(a, ntt_dif_pos<LOGR>(k), c));
        uint64_t go;
        if (!a.is_final) go = base + ((uint64_t)k << log_stride) + c;
        else go = K0 + c + ((uint64_t)k << log_q);
        // values travel between the passes in lazy form (< 2r); the last pass makes them canonical — through the scaling
        // product when there is one, else explicitly
        if (a.is_final) v = a.out_scale_on ? fp_mul(v, fr_from_words(a.out_scale[go % 3])) : fp_canon(v);
        fr_store2(ntt_dst_elem(a, col, go), 0, v);
    }
}

// number of DIF rounds and their radices for an R = 2^LOGR block: 3,3,...,rem
ZKB_HD int ntt_num_rounds(int logr) { return (logr + 2) / 3; }

// the whole CTA program (used by the kernel with a real barrier, by the emulator with a thread loop)
template <int LOGR, class Barrier>
ZKB_HD void ntt_cta_program(const NttPassArgs& a, uint4* sm, uint32_t tid, uint32_t nthreads, uint64_t cta,
                            uint32_t col, Barrier& bar) {
    cta = ntt_dist_cta(a, cta);
    ntt_phase_load<LOGR>(a, sm, tid, nthreads, cta, col);
    bar.sync();
    int left = LOGR;
    while (left > 0) {
        int t = left >= 3 ? 3 : left;
        ntt_phase_round<LOGR>(a, sm, tid, nthreads, (uint32_t)left, (uint32_t)t);
        bar.sync();
        left -= t;
    }
    ntt_phase_store<LOGR>(a, sm, tid, nthreads, cta, col);
}

}  // namespace zkb
