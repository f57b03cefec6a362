// msm.cuh — BN254 G1 multi-scalar multiplication, device code as per-thread functions.
//
// Replaces halo2_proofs::arithmetic::best_multiexp.  The result is a group element, so
// it does not depend on the algorithm; this one is a sort-based Pippenger:
//
//   1. digits   : scalar -> canonical integer (one Montgomery mul by 1, == Fr::to_repr) -> W signed c-bit digits
//                 d_w in [-2^(c-1)+1, 2^(c-1)];  emits key = w*2^(c-1) + |d_w|-1, value = point index | sign<<31
//                 (zero digits get key = INVALID and sort to the end).
//   2. sort     : radix sort of (key, value) pairs — all points of one bucket become one contiguous run.
//   3. accumulate: the sorted array is cut into fixed chunks of S entries, one thread per chunk, so the load is
//                 balanced for ANY digit distribution (witness-like columns put most points in a few buckets).
//                 A thread sums each run of equal keys with mixed additions (XYZZ accumulator).  Runs strictly
//                 inside a chunk are complete buckets and are written to the bucket array; the first and last
//                 run of a chunk may continue in the neighbours: the threads of a CTA combine theirs by a tree in
//                 shared memory (warp-/CTA-cooperative XYZZ additions), the CTA's own first and last run go to a
//                 (key, partial) list reduced by the same routine one level up until one CTA remains.
//   4. reduce   : per window sum_b (b+1)*B[b] by segments: running sums inside a segment of m buckets, plus
//                 (segment offset)*(segment sum) by a short double-and-add; segment results are summed by
//                 CTA-cooperative trees.
//   5. the W window sums (128 B each) go to the host, which does the c-bit Horner combination and the final
//                 normalisation (north_star: "tiny bucket sums combined on the host").
#pragma once
#include "curve.cuh"

namespace zkb {

constexpr uint32_t MSM_INVALID_KEY = 0xffffffffu;

struct MsmDigitArgs {
    const uint4* scalars;  // ncols columns of n Montgomery Fr, column-major contiguous
    uint64_t n;            // points per column
    uint32_t ncols;        // independent MSMs against the same bases (the prover's per-column commits) in one pass
    uint32_t sets_per_col; // bucket sets per column: W, or 1 in table mode
    uint32_t c;            // window bits
    uint32_t nwin;         // W, W*c >= 255
    uint32_t* keys;        // [<= W * ncols * n] compacted: only non-zero digits, in arbitrary order
    uint32_t* vals;
    unsigned long long* counter;  // number of entries written (device)
    uint32_t invalid_key;  // number of buckets: W << (c-1), or 1 << (c-1) in table mode
    uint32_t table_mode;   // 1: bases are the precomputed rows T[w][i] = 2^(c w) P_i, all windows share one bucket set
    uint64_t row_stride;   // table mode: points per row (the registered SRS length)
};

ZKB_HD uint32_t msm_extract_bits(const Fr& s, uint32_t bit, uint32_t c) {  // c <= 24
    if (bit >= 256) return 0;
    uint32_t w = bit >> 5, sh = bit & 31;
    uint64_t v = s.l[w];
    if (w + 1 < 8) v |= (uint64_t)s.l[w + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1);
}

constexpr uint32_t MSM_MAX_WINDOWS = 128;  // c >= 2

// Digits of scalar t: fills (key, value) for every NON-ZERO digit and returns how many there are.  Zero digits (half of a
// witness column is zeros, most of the rest are small values) produce no entry at all, so they are neither sorted nor
// visited by the accumulation.
// calls f(key, value) for every non-zero signed digit of the canonical scalar s (scalar index t).
// CT_C > 0: window width known at compile time (the loop unrolls, limb indices become constants and `s` stays in
// registers); CT_C == 0: runtime a.c / a.nwin.
template <uint32_t CT_C, class F>
ZKB_HD void msm_digits_foreach(const MsmDigitArgs& a, uint64_t t, const Fr& s, F&& f) {
    const uint32_t col = (uint32_t)(t / a.n);
    const uint64_t i = t - (uint64_t)col * a.n;
    const uint32_t c = CT_C ? CT_C : a.c;
    const uint32_t nwin = CT_C ? (255 + CT_C - 1) / CT_C : a.nwin;
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t w = 0; w < nwin; ++w) {
        uint32_t v = msm_extract_bits(s, w * c, c) + carry;
        uint32_t neg = v > half;
        uint32_t mag = neg ? (1u << c) - v : v;
        carry = neg;
        if (!mag) continue;
        const uint32_t set = col * a.sets_per_col + (a.table_mode ? 0u : w);
        const uint32_t idx = a.table_mode ? (uint32_t)((uint64_t)w * a.row_stride + i) : (uint32_t)i;
        f((set << (c - 1)) + (mag - 1), idx | (neg << 31));
    }
}

ZKB_HD uint32_t msm_digits_compute(const MsmDigitArgs& a, uint64_t t, uint32_t* keys, uint32_t* vals) {
    const Fr s = fp_from_mont(fr_load2(a.scalars, t));
    uint32_t cnt = 0;
    msm_digits_foreach<0>(a, t, s, [&](uint32_t k, uint32_t v) { keys[cnt] = k; vals[cnt] = v; ++cnt; });
    return cnt;
}

// writes the entries of one scalar at [base, base + cnt)
ZKB_HD void msm_digits_emit(const MsmDigitArgs& a, uint64_t base, const uint32_t* keys, const uint32_t* vals, uint32_t cnt) {
    for (uint32_t j = 0; j < cnt; ++j) {
        a.keys[base + j] = keys[j];
        a.vals[base + j] = vals[j];
    }
}

// ---- accumulate ----------------------------------------------------------------------------------------------------
struct MsmAccArgs {
    const uint32_t* keys;   // level 0: sorted keys; level >= 1: partial keys (INVALID entries are skipped)
    const uint32_t* vals;   // level 0: point index | sign<<31
    const uint4* bases;     // level 0: affine points (64 B each)
    const uint4* pin;       // level >= 1: partial XYZZ values (128 B each)
    uint64_t count;         // entries at this level (level 0 with count_dev: an upper bound, the launches' geometry)
    const unsigned long long* count_dev;   // level 0, latency regime: the entry count is read on the device (no host round trip)
    uint32_t chunk;         // S
    uint32_t invalid_key;
    uint32_t last_level;    // single CTA: its head/tail are complete and go to the bucket array
    uint4* buckets;         // XYZZ [nbuckets], zero-initialised
    uint32_t* pkeys_out;    // [2 * CTAs]
    uint4* pvals_out;       // XYZZ [2 * CTAs]
};

ZKB_HD void msm_store_xyzz(uint4* base, uint64_t idx, const XYZZ& p) { p.store(base + 8 * idx); }
ZKB_HD XYZZ msm_load_xyzz(const uint4* base, uint64_t idx) { return XYZZ::load(base + 8 * idx); }

// ---- lane-cooperative XYZZ addition (the serial tails of a commit) ---------------------------------------------------------------
// One full addition is 14 field products of which at most 4 are independent at any time; on one thread that is 14 dependent
// product-times (~4 us), and the finishing levels of a commit are chains of such additions with most lanes idle.  Here FOUR
// ADJACENT LANES share one addition: four product phases (u1 u2 s1 s2 | pp rr zz12 zzz12 | ppp q zz3 | t v zzz3) exchanged through a
// small shared-memory scratch, so the chain is 4 product-times long.  Same formulas as xyzz_add (add-2008-s), same canonical
// values, exceptional cases handled exactly: an identity operand returns the other one, equal points are doubled (serially, by
// lane 0: it does not occur for bucket sums of an honest commit but must be right), opposite points give the identity.
// Every phase is a per-thread function of (role = lane & 3, operands, scratch); the kernel puts a barrier between phases and
// the CPU emulator runs them in the same order.
constexpr uint32_t COOP_SCRATCH_FQ = 14;   // Fq slots per addition in flight
enum : uint32_t { COOP_U1, COOP_U2, COOP_S1, COOP_S2, COOP_PP, COOP_RR, COOP_ZZ12, COOP_ZZZ12, COOP_PPP, COOP_Q, COOP_ZZ3, COOP_T, COOP_V, COOP_ZZZ3 };
enum : uint32_t { COOP_ADD = 0, COOP_TAKE_A = 1, COOP_TAKE_B = 2 };

struct CoopOperands {
    const uint4* a;   // XYZZ (8 uint4)
    const uint4* b;
    uint4* scr;       // COOP_SCRATCH_FQ x 2 uint4
    uint4* out;       // XYZZ, may alias a
};
ZKB_HD Fq coop_get(const uint4* scr, uint32_t slot) { return Fq::load(scr + 2 * slot); }
ZKB_HD void coop_put(uint4* scr, uint32_t slot, const Fq& v) { v.store(scr + 2 * slot); }
ZKB_HD Fq coop_coord(const uint4* p, uint32_t k) { return Fq::load(p + 2 * k); }   // 0 x, 1 y, 2 zz, 3 zzz
// which of the three outcomes the operands' identity flags select (every lane evaluates it: two 32-byte loads)
ZKB_HD uint32_t coop_case(const CoopOperands& o) {
    if (coop_coord(o.b, 2).is_zero()) return COOP_TAKE_A;
    if (coop_coord(o.a, 2).is_zero()) return COOP_TAKE_B;
    return COOP_ADD;
}
// The four product phases.  Every lane runs the SAME instruction stream — operands are selected by role with predicated moves,
// then one fp_mul — so the four products of a phase issue as one warp instruction stream (a switch over the role would
// serialise them: lanes of one warp cannot run different code at the same time).
ZKB_HD Fq coop_sel(bool c, const Fq& x, const Fq& y) {
    Fq r;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; ++i) r.l[i] = c ? x.l[i] : y.l[i];
    return r;
}
ZKB_HD void coop_add_p1(const CoopOperands& o, uint32_t role) {   // u1 = X1 ZZ2, u2 = X2 ZZ1, s1 = Y1 ZZZ2, s2 = Y2 ZZZ1
    if (coop_case(o) != COOP_ADD) return;
    const uint4* first = (role & 1) ? o.b : o.a;
    const uint4* second = (role & 1) ? o.a : o.b;
    const uint32_t hi = role >> 1;   // 0: x, zz   1: y, zzz
    coop_put(o.scr, COOP_U1 + role, fp_mul(coop_coord(first, hi), coop_coord(second, 2 + hi)));
}
ZKB_HD void coop_add_p2(const CoopOperands& o, uint32_t role) {   // pp = (u2 - u1)^2, rr = (s2 - s1)^2, zz12, zzz12
    if (coop_case(o) != COOP_ADD) return;
    const bool diff = role < 2;
    const uint32_t pr = 2 * (role & 1);
    const Fq d = fp_sub(coop_get(o.scr, COOP_U1 + pr + 1), coop_get(o.scr, COOP_U1 + pr));   // p (role 0) / r (role 1)
    const Fq x = coop_sel(diff, d, coop_coord(o.a, role));       // roles 2, 3: ZZ1 / ZZZ1
    const Fq y = coop_sel(diff, d, coop_coord(o.b, role));       //             ZZ2 / ZZZ2
    coop_put(o.scr, COOP_PP + role, fp_mul(x, y));
}
ZKB_HD bool coop_p_is_zero(const CoopOperands& o) { return coop_get(o.scr, COOP_U2) == coop_get(o.scr, COOP_U1); }
ZKB_HD void coop_add_p3(const CoopOperands& o, uint32_t role) {   // ppp = p pp, q = u1 pp, zz3 = zz12 pp   (role 3: nothing to do)
    if (coop_case(o) != COOP_ADD || coop_p_is_zero(o)) return;
    const Fq pp = coop_get(o.scr, COOP_PP);
    const Fq p = fp_sub(coop_get(o.scr, COOP_U2), coop_get(o.scr, COOP_U1));
    const Fq x = coop_sel(role == 0, p, coop_get(o.scr, role == 1 ? COOP_U1 : COOP_ZZ12));
    const Fq z = fp_mul(x, pp);
    if (role < 3) coop_put(o.scr, COOP_PPP + role, z);
}
ZKB_HD void coop_add_p4(const CoopOperands& o, uint32_t role) {   // t = r (q - x3), v = s1 ppp, zzz3 = zzz12 ppp
    if (coop_case(o) != COOP_ADD || coop_p_is_zero(o)) return;
    const Fq ppp = coop_get(o.scr, COOP_PPP);
    const Fq q = coop_get(o.scr, COOP_Q);
    const Fq x3 = fp_sub(fp_sub(coop_get(o.scr, COOP_RR), ppp), fp_dbl(q));
    const Fq r = fp_sub(coop_get(o.scr, COOP_S2), coop_get(o.scr, COOP_S1));
    const Fq x = coop_sel(role == 0, r, coop_get(o.scr, role == 1 ? COOP_S1 : COOP_ZZZ12));
    const Fq y = coop_sel(role == 0, fp_sub(q, x3), ppp);
    const Fq z = fp_mul(x, y);
    if (role < 3) coop_put(o.scr, COOP_T + role, z);
    if (role == 3) coop_put(o.scr, COOP_PP, x3);   // pp is dead (read above, before any lane of this group stores): its slot carries x3
}
// lane 0 assembles the result (and handles the outcomes that need no products, or the serial doubling)
ZKB_HD void coop_add_p5(const CoopOperands& o, uint32_t role) {
    if (role != 0) return;
    const uint32_t c = coop_case(o);
    if (c == COOP_TAKE_A) { if (o.out != o.a) msm_store_xyzz(o.out, 0, msm_load_xyzz(o.a, 0)); return; }
    if (c == COOP_TAKE_B) { msm_store_xyzz(o.out, 0, msm_load_xyzz(o.b, 0)); return; }
    if (coop_p_is_zero(o)) {
        if (coop_get(o.scr, COOP_S2) == coop_get(o.scr, COOP_S1)) msm_store_xyzz(o.out, 0, xyzz_double(msm_load_xyzz(o.a, 0)));
        else msm_store_xyzz(o.out, 0, XYZZ::identity());
        return;
    }
    XYZZ r;
    r.x = coop_get(o.scr, COOP_PP);
    r.y = fp_sub(coop_get(o.scr, COOP_T), coop_get(o.scr, COOP_V));
    r.zz = coop_get(o.scr, COOP_ZZ3);
    r.zzz = coop_get(o.scr, COOP_ZZZ3);
    msm_store_xyzz(o.out, 0, r);
}
ZKB_HD void coop_add_phase(uint32_t phase, const CoopOperands& o, uint32_t role) {
    switch (phase) {
        case 0: coop_add_p1(o, role); break;
        case 1: coop_add_p2(o, role); break;
        case 2: coop_add_p3(o, role); break;
        case 3: coop_add_p4(o, role); break;
        default: coop_add_p5(o, role); break;
    }
}
constexpr uint32_t COOP_PHASES = 5;

// One CTA of MSM_ACC_CTA threads.  Phase 1: every thread sums its chunk; runs strictly inside the chunk are complete buckets
// (stored), the first and the last run (which may continue in the neighbouring chunks) are left as the thread's SUMMARY in shared
// memory: skey / sval [2 tid] = head run, [2 tid + 1] = tail run (INVALID key = absent; a chunk with one run has a head only).
// Phase 2: the summaries of the CTA are combined by a binary tree (log2 MSM_ACC_CTA steps, at most one full addition per
// combination): a run that turns out to lie strictly inside the combined range is complete and is stored, so the CTA ends
// with ONE head and ONE tail.  Phase 3: thread 0 writes them to the partial list of the next level (2 entries per CTA — the
// list shrinks by chunk x 64 per level, so two or three launches finish any MSM), or to the buckets when this CTA is the last.
// Every bucket is written exactly once over all levels.
constexpr uint32_t MSM_ACC_CTA = 128;
constexpr uint32_t COOP_GROUPS = MSM_ACC_CTA / 4;   // lane-cooperative additions in flight per CTA

template <bool LEVEL0>
ZKB_HD void msm_acc_phase_chunk(const MsmAccArgs& a, uint64_t t, uint32_t tid, uint32_t* skey, uint4* sval) {
    const uint64_t lo = t * a.chunk;
    uint32_t head_key = MSM_INVALID_KEY, tail_key = MSM_INVALID_KEY;
    if (lo < a.count) {
        const uint64_t hi = lo + a.chunk < a.count ? lo + a.chunk : a.count;
        XYZZ acc = XYZZ::identity();
        uint32_t cur = MSM_INVALID_KEY;
        uint32_t runs_done = 0;       // completed runs before the current one
        // level 0 keeps the accumulator lazy (coordinates < 2p, see xyzz_add_mixed_lazy); it becomes canonical when it leaves
        auto out = [](const XYZZ& p) { return LEVEL0 ? xyzz_canon(p) : p; };
        for (uint64_t i = lo; i < hi; ++i) {
            uint32_t k = a.keys[i];
            if (k >= a.invalid_key) {
                if (LEVEL0) break;   // sorted: only invalid entries follow
                continue;            // partial lists carry holes
            }
            if (k != cur) {
                if (cur != MSM_INVALID_KEY) {
                    if (runs_done == 0) { head_key = cur; msm_store_xyzz(sval, 2 * tid, out(acc)); }  // may continue to the left
                    else msm_store_xyzz(a.buckets, cur, out(acc));   // strictly interior run == complete bucket
                    ++runs_done;
                }
                acc = XYZZ::identity();
                cur = k;
            }
            if (LEVEL0) {
                uint32_t v = a.vals[i];
                Affine p = affine_load(a.bases + 4 * (uint64_t)(v & 0x7fffffffu));
                if (!p.is_identity()) {
                    if (v >> 31) p.y = fp_neg(p.y);
                    xyzz_add_mixed_lazy(acc, p.x, p.y);
                }
            } else {
                xyzz_add(acc, msm_load_xyzz(a.pin, i));
            }
        }
        if (cur != MSM_INVALID_KEY) {
            if (runs_done == 0) { head_key = cur; msm_store_xyzz(sval, 2 * tid, out(acc)); }
            else { tail_key = cur; msm_store_xyzz(sval, 2 * tid + 1, out(acc)); }
        }
    }
    skey[2 * tid] = head_key;
    skey[2 * tid + 1] = tail_key;
}

// Throughput regime (many waves of CTAs): no tree — every thread writes its summary straight to the partial list (two entries
// per thread, slots 2t and 2t+1), and the next level, which is small, does the combining.  `skey` / `sval` of the chunk phase
// then point into the global lists.
template <bool LEVEL0>
ZKB_HD void msm_acc_thread_direct(const MsmAccArgs& a, uint64_t t) {
    msm_acc_phase_chunk<LEVEL0>(a, t, 0, a.pkeys_out + 2 * t, a.pvals_out + 16 * t);
}

// tree step d, round `base`: summary[l] (threads l .. l+d-1) absorbs summary[l+d] (threads l+d .. l+2d-1), l = 2 d pair,
// pair = base + tid / 4.  The (at most one) full addition of a combination is shared by FOUR lanes (lane-cooperative addition
// below: 4 product-times instead of 14) and the pairs of a step are taken by the first threads, so the additions fill whole warps;
// `phase` runs 0 .. COOP_PHASES-1 with a barrier in between, the bookkeeping (keys, copies, stores of runs that became complete)
// is done by lane 0 in the last phase.
ZKB_HD void msm_acc_phase_combine(const MsmAccArgs& a, uint32_t tid, uint32_t d, uint32_t base, uint32_t phase, uint32_t nthreads,
                                  uint32_t* skey, uint4* sval, uint4* scr) {
    const uint32_t pair = base + (tid >> 2), role = tid & 3;
    const uint32_t left = 2 * d * pair;
    if (left + d >= nthreads) return;
    const uint32_t ia = 2 * left, ib = 2 * (left + d);
    const uint32_t ah = skey[ia], at = skey[ia + 1], bh = skey[ib], bt = skey[ib + 1];
    if (bh == MSM_INVALID_KEY) return;  // right half empty
    const bool book = phase + 1 == COOP_PHASES && role == 0;
    auto copy = [&](uint32_t dst, uint32_t src) {
        for (int q = 0; q < 8; ++q) sval[8 * dst + q] = sval[8 * src + q];
    };
    if (ah == MSM_INVALID_KEY) {        // left half empty
        if (book) {
            skey[ia] = bh; skey[ia + 1] = bt;
            copy(ia, ib);
            if (bt != MSM_INVALID_KEY) copy(ia + 1, ib + 1);
        }
        return;
    }
    const bool a_multi = at != MSM_INVALID_KEY, b_multi = bt != MSM_INVALID_KEY;
    const uint32_t a_last = a_multi ? at : ah;
    if (a_last == bh) {                 // the run continues across the boundary: one full addition
        const uint32_t sa = a_multi ? ia + 1 : ia;
        // both halves multi-run: the merged run is now strictly inside the combined range, i.e. a complete bucket
        uint4* out = (a_multi && b_multi) ? a.buckets + 8 * (uint64_t)a_last : sval + 8 * sa;
        CoopOperands o{sval + 8 * sa, sval + 8 * ib, scr + 2 * COOP_SCRATCH_FQ * (tid >> 2), out};
        coop_add_phase(phase, o, role);
        if (book && b_multi) { skey[ia + 1] = bt; copy(ia + 1, ib + 1); }
    } else if (book) {
        if (a_multi) msm_store_xyzz(a.buckets, at, msm_load_xyzz(sval, ia + 1));
        if (b_multi) {
            msm_store_xyzz(a.buckets, bh, msm_load_xyzz(sval, ib));
            skey[ia + 1] = bt;
            copy(ia + 1, ib + 1);
        } else {
            skey[ia + 1] = bh;
            copy(ia + 1, ib);
        }
    }
}

// thread 0, after the tree
ZKB_HD void msm_acc_phase_emit(const MsmAccArgs& a, uint64_t cta, const uint32_t* skey, const uint4* sval) {
    const uint32_t hk = skey[0], tk = skey[1];
    if (a.last_level) {
        if (hk != MSM_INVALID_KEY) msm_store_xyzz(a.buckets, hk, msm_load_xyzz(sval, 0));
        if (tk != MSM_INVALID_KEY) msm_store_xyzz(a.buckets, tk, msm_load_xyzz(sval, 1));
        return;
    }
    a.pkeys_out[2 * cta] = hk;
    a.pkeys_out[2 * cta + 1] = tk;
    if (hk != MSM_INVALID_KEY) msm_store_xyzz(a.pvals_out, 2 * cta, msm_load_xyzz(sval, 0));
    if (tk != MSM_INVALID_KEY) msm_store_xyzz(a.pvals_out, 2 * cta + 1, msm_load_xyzz(sval, 1));
}

// ---- slice merge: main[b] += part[b] -----------------------------------------------------------------------------------
// A pipelined host-scalar commit runs as point-range slices; every slice after the first accumulates into a scratch bucket
// array which is then folded into the main one by this uniform kernel (one thread per bucket).  Doing the add inside the
// accumulate kernel instead was measured at +20 ms per doubling of the slice count (the run-end path diverges the warp).
struct MsmMergeArgs {
    uint4* main;
    const uint4* part;
    uint64_t nbuckets;
};
ZKB_HD void msm_merge_thread(const MsmMergeArgs& a, uint64_t b) {
    if (b >= a.nbuckets) return;
    XYZZ p = msm_load_xyzz(a.part, b);
    if (p.is_identity()) return;
    XYZZ m = msm_load_xyzz(a.main, b);
    xyzz_add(m, p);
    msm_store_xyzz(a.main, b, m);
}

// ---- bucket reduction ------------------------------------------------------------------------------------------------
struct MsmReduceArgs {
    const uint4* buckets;   // XYZZ [nwin << (c-1)]
    uint32_t c;
    uint32_t nwin;
    uint32_t log_m;         // segment length m = 2^log_m buckets
    uint4* seg_out;         // XYZZ [nwin * J], J = 2^(c-1-log_m)
};

// k*P for a small public k (double-and-add from the top bit)
ZKB_HD XYZZ xyzz_mul_small(const XYZZ& p, uint32_t k) {
    XYZZ r = XYZZ::identity();
    if (k == 0 || p.is_identity()) return r;
    int top = 31;
    while (!((k >> top) & 1)) --top;
    for (int i = top; i >= 0; --i) {
        r = xyzz_double(r);
        if ((k >> i) & 1) xyzz_add(r, p);
    }
    return r;
}

// thread t = w*J + j : sum_{i<m} (j*m + i + 1) * B[w][j*m + i]
ZKB_HD void msm_reduce_segment_thread(const MsmReduceArgs& a, uint64_t t) {
    const uint32_t log_j = a.c - 1 - a.log_m;
    const uint64_t total = (uint64_t)a.nwin << log_j;
    if (t >= total) return;
    const uint32_t w = (uint32_t)(t >> log_j), j = (uint32_t)(t & ((1u << log_j) - 1));
    const uint32_t m = 1u << a.log_m;
    const uint64_t first = ((uint64_t)w << (a.c - 1)) + ((uint64_t)j << a.log_m);
    XYZZ run = XYZZ::identity(), acc = XYZZ::identity();
    for (uint32_t i = m; i-- > 0;) {
        xyzz_add(run, msm_load_xyzz(a.buckets, first + i));
        xyzz_add(acc, run);
    }
    XYZZ off = xyzz_mul_small(run, j << a.log_m);
    xyzz_add(acc, off);
    msm_store_xyzz(a.seg_out, t, acc);
}

// ---- the same segment sum shared by four lanes per segment (latency regime) ----------------------------------------------------
// Lane-cooperative doubling [dbl-2008-s-1, a = 0]: three product phases (v xx | w s mm zz3 | t wy zzz3) + assembly.
enum : uint32_t { COOPD_V = 0, COOPD_XX, COOPD_W, COOPD_S, COOPD_MM, COOPD_ZZ3, COOPD_T, COOPD_WY, COOPD_ZZZ3, COOPD_X3 };
constexpr uint32_t COOP_DBL_PHASES = 4;
ZKB_HD void coop_dbl_phase(uint32_t phase, const uint4* p, uint4* scr, uint4* out, uint32_t role) {
    if (coop_coord(p, 2).is_zero()) {   // identity: stays the identity
        if (phase + 1 == COOP_DBL_PHASES && role == 0 && out != p) msm_store_xyzz(out, 0, XYZZ::identity());
        return;
    }
    if (phase == 0) {            // v = (2Y)^2, xx = X^2
        const Fq y = coop_coord(p, 1);
        const Fq x = coop_sel(role == 0, fp_dbl(y), coop_coord(p, 0));
        const Fq z = fp_mul(x, x);
        if (role < 2) coop_put(scr, COOPD_V + role, z);
    } else if (phase == 1) {     // w = u v, s = X v, mm = (3 xx)^2, zz3 = v ZZ
        const Fq v = coop_get(scr, COOPD_V), xx = coop_get(scr, COOPD_XX);
        const Fq m = fp_add(fp_dbl(xx), xx);
        const Fq u = fp_dbl(coop_coord(p, 1));
        const Fq x = coop_sel(role == 0, u, coop_sel(role == 1, coop_coord(p, 0), coop_sel(role == 2, m, coop_coord(p, 2))));
        const Fq y = coop_sel(role == 2, m, v);
        coop_put(scr, COOPD_W + role, fp_mul(x, y));
    } else if (phase == 2) {     // t = m (s - x3), wy = w Y, zzz3 = w ZZZ
        const Fq xx = coop_get(scr, COOPD_XX), sv = coop_get(scr, COOPD_S), w = coop_get(scr, COOPD_W);
        const Fq m = fp_add(fp_dbl(xx), xx);
        const Fq x3 = fp_sub(coop_get(scr, COOPD_MM), fp_dbl(sv));
        const Fq x = coop_sel(role == 0, m, w);
        const Fq y = coop_sel(role == 0, fp_sub(sv, x3), coop_coord(p, role == 1 ? 1 : 3));
        const Fq z = fp_mul(x, y);
        if (role < 3) coop_put(scr, COOPD_T + role, z);
        if (role == 3) coop_put(scr, COOPD_X3, x3);
    } else if (role == 0) {
        XYZZ r;
        r.x = coop_get(scr, COOPD_X3);
        r.y = fp_sub(coop_get(scr, COOPD_T), coop_get(scr, COOPD_WY));
        r.zz = coop_get(scr, COOPD_ZZ3);
        r.zzz = coop_get(scr, COOPD_ZZZ3);
        msm_store_xyzz(out, 0, r);
    }
}

// Number of lock-step steps of one segment: 2m additions, then (segment offset) * (segment sum) by double-and-add over the log_j
// bits of j (every bit pays its conditional addition's phases: the lanes of a warp hold different j), log_m more doublings, and the
// final addition.
ZKB_HD uint32_t msm_reduce_coop_steps(const MsmReduceArgs& a) {
    const uint32_t log_j = a.c - 1 - a.log_m, m = 1u << a.log_m;
    return 2 * m * COOP_PHASES + log_j * (COOP_DBL_PHASES + COOP_PHASES) + a.log_m * COOP_DBL_PHASES + COOP_PHASES;
}
// state: run, acc, off (3 XYZZ) of the group; scr: COOP_SCRATCH_FQ Fq; segment t = w*J + j as in msm_reduce_segment_thread
ZKB_HD void msm_reduce_coop_init(uint32_t role, uint4* state) {
    if (role < 3) msm_store_xyzz(state, role, XYZZ::identity());
}
ZKB_HD void msm_reduce_coop_step(const MsmReduceArgs& a, uint64_t t, uint32_t role, uint32_t step, uint4* state, uint4* scr) {
    const uint32_t log_j = a.c - 1 - a.log_m;
    if (t >= ((uint64_t)a.nwin << log_j)) return;
    const uint32_t w = (uint32_t)(t >> log_j), j = (uint32_t)(t & ((1u << log_j) - 1));
    const uint32_t m = 1u << a.log_m;
    const uint64_t first = ((uint64_t)w << (a.c - 1)) + ((uint64_t)j << a.log_m);
    uint4* run = state;
    uint4* acc = state + 8;
    uint4* off = state + 16;
    uint32_t s = step;
    if (s < 2 * m * COOP_PHASES) {                      // run += B[i]; acc += run   for i = m-1 .. 0
        const uint32_t k = s / COOP_PHASES, i = m - 1 - k / 2;
        CoopOperands o = (k & 1) ? CoopOperands{acc, run, scr, acc} : CoopOperands{run, a.buckets + 8 * (first + i), scr, run};
        coop_add_phase(s % COOP_PHASES, o, role);
        return;
    }
    s -= 2 * m * COOP_PHASES;
    const uint32_t per_bit = COOP_DBL_PHASES + COOP_PHASES;
    if (s < log_j * per_bit) {                          // off = 2 off (+ run when the bit of j is set), from the top bit
        const uint32_t b = log_j - 1 - s / per_bit, q = s % per_bit;
        if (q < COOP_DBL_PHASES) coop_dbl_phase(q, off, scr, off, role);
        else if ((j >> b) & 1) coop_add_phase(q - COOP_DBL_PHASES, CoopOperands{off, run, scr, off}, role);
        return;
    }
    s -= log_j * per_bit;
    if (s < a.log_m * COOP_DBL_PHASES) { coop_dbl_phase(s % COOP_DBL_PHASES, off, scr, off, role); return; }   // * m
    s -= a.log_m * COOP_DBL_PHASES;
    coop_add_phase(s, CoopOperands{acc, off, scr, a.seg_out + 8 * t}, role);
}

// CTA-cooperative sum: CTA (set, tile) adds up to MSM_ACC_CTA * group consecutive elements of its set — `group` serial
// additions per thread, then a log2(MSM_ACC_CTA)-step tree in shared memory — so a whole set of segment sums is folded in one
// or two launches whose dependent chains are ~15 additions long (the serial fan-in-8 chain it replaces took five launches).
struct MsmSumTreeArgs {
    const uint4* in;     // XYZZ [sets * in_per_set]
    uint4* out;          // XYZZ [sets * out_per_set]
    uint64_t in_per_set, out_per_set;
    uint32_t group;
};
ZKB_HD void msm_sum_tree_phase_load(const MsmSumTreeArgs& a, uint64_t cta, uint32_t tid, uint4* sval) {
    const uint64_t set = cta / a.out_per_set, tile = cta % a.out_per_set;
    const uint64_t first = (tile * MSM_ACC_CTA + tid) * a.group;
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = 0; i < a.group && first + i < a.in_per_set; ++i) xyzz_add(acc, msm_load_xyzz(a.in, set * a.in_per_set + first + i));
    msm_store_xyzz(sval, tid, acc);
}
// tree step d, round `base`: pair t = base + tid / 4 (t < d) is sval[t] += sval[t + d], shared by four lanes; `phase` of COOP_PHASES
ZKB_HD void msm_sum_tree_phase_step(uint32_t tid, uint32_t d, uint32_t base, uint32_t phase, uint4* sval, uint4* scr) {
    const uint32_t t = base + (tid >> 2);
    if (t >= d) return;
    CoopOperands o{sval + 8 * t, sval + 8 * (t + d), scr + 2 * COOP_SCRATCH_FQ * (tid >> 2), sval + 8 * t};
    coop_add_phase(phase, o, tid & 3);
}
ZKB_HD void msm_sum_tree_phase_store(const MsmSumTreeArgs& a, uint64_t cta, const uint4* sval) {
    msm_store_xyzz(a.out, cta, msm_load_xyzz(sval, 0));
}

// ---- window combination + normalisation (host side of step 5; also usable on the device) ----------------------------
// a^(p-2)
ZKB_HD_NOINLINE Fq fq_inv(const Fq& a) {
    Fq acc = Fq::one();
    for (int i = 255; i >= 0; --i) {
        acc = fp_sqr(acc);
        uint32_t limb = FqParams::M(i >> 5);
        if (i < 32) limb -= 2;  // p - 2 : low limb 0xd87cfd47 - 2, no borrow
        if ((limb >> (i & 31)) & 1) acc = fp_mul(acc, a);
    }
    return acc;
}

// XYZZ -> affine (identity -> (0,0))
ZKB_HD_NOINLINE Affine xyzz_to_affine(const XYZZ& p) {
    Affine r;
    if (p.is_identity()) { r.x = Fq::zero(); r.y = Fq::zero(); return r; }
    Fq inv = fq_inv(fp_mul(p.zz, p.zzz));
    r.x = fp_mul(p.x, fp_mul(inv, p.zzz));
    r.y = fp_mul(p.y, fp_mul(inv, p.zz));
    return r;
}

// sum_w 2^(c*w) * S_w by Horner from the top window
ZKB_HD_NOINLINE XYZZ msm_combine_windows(const XYZZ* sums, uint32_t nwin, uint32_t c) {
    XYZZ acc = XYZZ::identity();
    for (uint32_t w = nwin; w-- > 0;) {
        for (uint32_t i = 0; i < c; ++i) acc = xyzz_double(acc);
        xyzz_add(acc, sums[w]);
    }
    return acc;
}

// ---- device-side finalisation for batches: column col folds its `sets` window sums (Horner, c doublings per window)
// and normalises to the G1 output encoding (x, y, R) / identity (0, R, 0) -------------------------------------------------
struct MsmFinalArgs {
    const uint4* sums;   // XYZZ [ncols * sets]
    uint32_t ncols, sets, c;
    uint4* out;          // ncols x 96 B
};
ZKB_HD void msm_finalize_thread(const MsmFinalArgs& a, uint64_t col) {
    if (col >= a.ncols) return;
    XYZZ acc = XYZZ::identity();
    for (uint32_t w = a.sets; w-- > 0;) {
        if (w + 1 < a.sets)
            for (uint32_t i = 0; i < a.c; ++i) acc = xyzz_double(acc);
        xyzz_add(acc, msm_load_xyzz(a.sums, col * a.sets + w));
    }
    uint4* o = a.out + 6 * col;
    Fq one = Fq::one();
    if (acc.is_identity()) {
        Fq::zero().store(o); one.store(o + 2); Fq::zero().store(o + 4);
        return;
    }
    Affine r = xyzz_to_affine(acc);
    r.x.store(o); r.y.store(o + 2); one.store(o + 4);
}

// ---- SRS window table: next[i] = 2^c * prev[i] (affine in, affine out) — built once per registered SRS so that every
// window of every scalar can share ONE bucket set (no per-window reduction, no Horner fold) ------------------------------------
struct SrsTableArgs {
    const uint4* bases;  // n affine points (row 0), read when first != 0
    uint4* acc;          // n XYZZ points, updated in place: acc[i] = 2^c * (first ? bases[i] : acc[i])
    uint64_t n;
    uint32_t c;
    uint32_t first;
};
// One row step of the table build; the XYZZ row is then converted to affine by the batched normalisation kernel
// (setup.cuh: one inversion per 16 points instead of one per point — the row build was 2.4x slower with per-point inversions).
ZKB_HD void srs_table_thread(const SrsTableArgs& a, uint64_t i) {
    if (i >= a.n) return;
    XYZZ acc;
    uint32_t left = a.c;
    if (a.first) {
        Affine p = affine_load(a.bases + 4 * i);
        if (p.is_identity()) { XYZZ::identity().store(a.acc + 8 * i); return; }
        acc = xyzz_double_affine(p.x, p.y);
        --left;
    } else {
        acc = msm_load_xyzz(a.acc, i);
    }
    for (uint32_t k = 0; k < left; ++k) acc = xyzz_double(acc);
    msm_store_xyzz(a.acc, i, acc);
}

// ---- fixed-base multiples of the generator: out[i] = [s_i] G (affine) — ParamsKZG::setup building block, and the
// generator of synthetic bases with known discrete logs for full-size parity checks --------------------------------------
struct FixedBaseArgs {
    const uint4* scalars;  // n Montgomery Fr
    uint64_t n;
    uint4* out;            // n affine points
};
ZKB_HD void g1_fixed_base_mul_thread(const FixedBaseArgs& a, uint64_t i) {
    if (i >= a.n) return;
    Fr s = fp_from_mont(fr_load2(a.scalars, i));
    Fq gx = Fq::one();
    Fq gy = fp_dbl(gx);  // G = (1, 2)
    XYZZ acc = XYZZ::identity();
    for (int b = 253; b >= 0; --b) {
        acc = xyzz_double(acc);
        if ((s.l[b >> 5] >> (b & 31)) & 1) xyzz_add_mixed(acc, gx, gy);
    }
    Affine r = xyzz_to_affine(acc);
    r.x.store(a.out + 4 * i);
    r.y.store(a.out + 4 * i + 2);
}

}  // namespace zkb
