// lookup.cuh — halo2's lookup argument, prover side: permute_expression_pair (plonk/lookup/prover.rs, un-vendored halo2-axiom;
// reached from create_proof, /root/reference/aggregator/src/wrapper.rs:129-137).  SURVEY.md §8f row 3.
//
// Given the compressed input expression A and table expression S on the usable rows, upstream produces
//   A' = A sorted ascending (Fr's Ord: canonical integer value),
//   S' with S'[i] = A'[i] on every row where A'[i] differs from A'[i-1] (first occurrence of a value; one copy of that value
//      is taken out of the table multiset — missing value = ConstraintSystemFailure), and the remaining table elements,
//      in ascending order, assigned to the repeated rows from the LAST repeated row backwards (`repeated_input_rows.pop()`).
// On the device: both columns are sorted by canonical value (a bitonic network over (value, row) records, below),
// first occurrences are flagged, each first occurrence finds the first equal element of the sorted table by binary search and
// marks it consumed, two exclusive scans compact the leftover table elements and the repeated rows, and one thread per leftover
// element writes it to its row.  Per-thread functions so that the CPU emulator runs the same code.
#pragma once
#include "field.cuh"

namespace zkb {

struct LookupArgs {
    const uint4* input;       // u Montgomery Fr (usable rows of the compressed input expression)
    const uint4* table;       // u Montgomery Fr
    uint64_t u;
    uint4* canon_in;          // u canonical values (integer limbs), input
    uint4* canon_tab;         // u canonical values, table
    const uint32_t* idx_in;   // sorted order of the input (by canonical value)
    const uint32_t* idx_tab;  // sorted order of the table
    uint32_t* first;          // [u] 1 = first occurrence of its value in the sorted input
    uint32_t* consumed;       // [u] 1 = sorted-table element taken by a first occurrence
    const uint32_t* rep_rank; // exclusive scan of (1 - first)
    const uint32_t* left_rank;// exclusive scan of (1 - consumed)
    uint32_t repeated;        // number of repeated rows (== number of leftover table elements)
    uint4* out_in;            // A' (Montgomery), n elements; rows >= u untouched
    uint4* out_tab;           // S'
    uint32_t* rep_rows;       // [repeated] rows of the repeated input values, ascending
    uint32_t* missing;        // set to 1 when an input value does not occur in the table
};

// canonical (integer) limbs of both columns
ZKB_HD void lookup_canon_thread(const LookupArgs& a, uint64_t i) {
    if (i >= a.u) return;
    fr_store2(a.canon_in, i, fp_from_mont(fr_load2(a.input, i)));
    fr_store2(a.canon_tab, i, fp_from_mont(fr_load2(a.table, i)));
}

// ---- the sort: a bitonic network over (canonical value, row) records -----------------------------------------------------------------
// The records are compared as the pair (256-bit value, original row), a TOTAL order, so the result is exactly what a stable sort by
// value gives and does not depend on the schedule.  P = 2^log_p >= u records (the padding carries the all-ones key, above every
// canonical value, and row 0xffffffff), structure-of-arrays so that every access is a full 16-byte or 4-byte vector.  Steps (k, j)
// with j >= LOOKUP_SORT_BLOCK exchange partners that are far apart: they run on global memory, TWO consecutive j per launch (a thread
// owns the four records that differ in bits j and j / 2, so the array is read and written once for two steps); all the steps with
// j < LOOKUP_SORT_BLOCK of one k (and the whole network up to k = LOOKUP_SORT_BLOCK) run inside one CTA on a block staged in shared
// memory.  Written by hand instead of calling a library sort: no library kernel is left anywhere in this package.
constexpr uint32_t LOOKUP_SORT_BLOCK = 4096;                 // records per CTA block (144 KB of shared memory: one CTA per SM)
constexpr uint32_t LOOKUP_SORT_LOG_BLOCK = 12;
constexpr uint32_t LOOKUP_SORT_THREADS = 1024;               // two pairs per thread and step
struct LookupSortArgs {
    uint4* klo;        // [P] low 128 bits of the value
    uint4* khi;        // [P] high 128 bits
    uint32_t* row;     // [P]
    uint32_t log_p;
    const uint4* canon;   // init: u canonical values (2 uint4 each)
    uint64_t u;
};
ZKB_HD bool lookup_rec_less(const uint4& alo, const uint4& ahi, uint32_t ar, const uint4& blo, const uint4& bhi, uint32_t br) {
    if (ahi.w != bhi.w) return ahi.w < bhi.w;
    if (ahi.z != bhi.z) return ahi.z < bhi.z;
    if (ahi.y != bhi.y) return ahi.y < bhi.y;
    if (ahi.x != bhi.x) return ahi.x < bhi.x;
    if (alo.w != blo.w) return alo.w < blo.w;
    if (alo.z != blo.z) return alo.z < blo.z;
    if (alo.y != blo.y) return alo.y < blo.y;
    if (alo.x != blo.x) return alo.x < blo.x;
    return ar < br;
}
ZKB_HD void lookup_sort_init_thread(const LookupSortArgs& a, uint64_t i) {
    if (i >= ((uint64_t)1 << a.log_p)) return;
    if (i < a.u) { a.klo[i] = a.canon[2 * i]; a.khi[i] = a.canon[2 * i + 1]; a.row[i] = (uint32_t)i; }
    else { a.klo[i] = make_uint4(~0u, ~0u, ~0u, ~0u); a.khi[i] = make_uint4(~0u, ~0u, ~0u, ~0u); a.row[i] = ~0u; }
}
// pair t of step (k, j): positions i < i | j, ascending when bit k of i is clear
ZKB_HD void lookup_sort_pair(uint64_t t, uint64_t j, uint64_t* i, uint64_t* p) {
    *i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    *p = *i | j;
}
ZKB_HD void lookup_sort_global_thread(const LookupSortArgs& a, uint64_t k, uint64_t j, uint64_t t) {
    if (t >= ((uint64_t)1 << a.log_p) / 2) return;
    uint64_t i, p;
    lookup_sort_pair(t, j, &i, &p);
    const uint4 alo = a.klo[i], ahi = a.khi[i], blo = a.klo[p], bhi = a.khi[p];
    const uint32_t ar = a.row[i], br = a.row[p];
    const bool up = (i & k) == 0;
    if (lookup_rec_less(blo, bhi, br, alo, ahi, ar) == up) {   // out of order for this direction: exchange
        a.klo[i] = blo; a.khi[i] = bhi; a.row[i] = br;
        a.klo[p] = alo; a.khi[p] = ahi; a.row[p] = ar;
    }
}
// steps (k, j) and (k, j / 2) in one pass: thread t owns the records at q, q | j/2, q | j, q | j | j/2 (q: bits j and j/2 clear)
ZKB_HD void lookup_sort_global2_thread(const LookupSortArgs& a, uint64_t k, uint32_t log_j, uint64_t t) {
    if (t >= ((uint64_t)1 << a.log_p) / 4) return;
    const uint64_t j = (uint64_t)1 << log_j, h = j >> 1;
    const uint64_t low = t & (h - 1), rest = t >> (log_j - 1);   // two zero bits inserted at log2(h) and log2(j)
    const uint64_t q = (rest << (log_j + 1)) | low;
    const uint64_t pos[4] = {q, q | h, q | j, q | j | h};
    uint4 lo[4], hi[4];
    uint32_t row[4];
    for (int e = 0; e < 4; ++e) { lo[e] = a.klo[pos[e]]; hi[e] = a.khi[pos[e]]; row[e] = a.row[pos[e]]; }
    const bool up = (q & k) == 0;   // k > j: the same direction for the four records
    auto cx = [&](int x, int y) {
        if (lookup_rec_less(lo[y], hi[y], row[y], lo[x], hi[x], row[x]) == up) {
            const uint4 tl = lo[x], th = hi[x]; const uint32_t tr = row[x];
            lo[x] = lo[y]; hi[x] = hi[y]; row[x] = row[y];
            lo[y] = tl; hi[y] = th; row[y] = tr;
        }
    };
    cx(0, 2); cx(1, 3);   // step j
    cx(0, 1); cx(2, 3);   // step j / 2
    for (int e = 0; e < 4; ++e) { a.klo[pos[e]] = lo[e]; a.khi[pos[e]] = hi[e]; a.row[pos[e]] = row[e]; }
}
// block kernel phases: shared arrays slo / shi / srow of LOOKUP_SORT_BLOCK records; block b covers [b, b + 1) * LOOKUP_SORT_BLOCK
ZKB_HD void lookup_sort_block_load(const LookupSortArgs& a, uint64_t block, uint32_t tid, uint4* slo, uint4* shi, uint32_t* srow) {
    for (uint32_t r = tid; r < LOOKUP_SORT_BLOCK; r += LOOKUP_SORT_THREADS) {
        const uint64_t g = block * LOOKUP_SORT_BLOCK + r;
        slo[r] = a.klo[g]; shi[r] = a.khi[g]; srow[r] = a.row[g];
    }
}
ZKB_HD void lookup_sort_block_step(uint64_t block, uint32_t tid, uint64_t k, uint32_t j, uint4* slo, uint4* shi, uint32_t* srow) {
    for (uint32_t t = tid; t < LOOKUP_SORT_BLOCK / 2; t += LOOKUP_SORT_THREADS) {
        uint64_t i, p;
        lookup_sort_pair(t, j, &i, &p);
        const bool up = ((block * LOOKUP_SORT_BLOCK + i) & k) == 0;
        const uint4 alo = slo[i], ahi = shi[i], blo = slo[p], bhi = shi[p];
        const uint32_t ar = srow[i], br = srow[p];
        if (lookup_rec_less(blo, bhi, br, alo, ahi, ar) == up) {
            slo[i] = blo; shi[i] = bhi; srow[i] = br;
            slo[p] = alo; shi[p] = ahi; srow[p] = ar;
        }
    }
}
ZKB_HD void lookup_sort_block_store(const LookupSortArgs& a, uint64_t block, uint32_t tid, const uint4* slo, const uint4* shi, const uint32_t* srow) {
    for (uint32_t r = tid; r < LOOKUP_SORT_BLOCK; r += LOOKUP_SORT_THREADS) {
        const uint64_t g = block * LOOKUP_SORT_BLOCK + r;
        a.klo[g] = slo[r]; a.khi[g] = shi[r]; a.row[g] = srow[r];
    }
}
// the schedule, shared by the driver and the emulator: calls global(k, j) / block(k_first, k_last, j_first) in network order
template <class G, class G2, class B>
inline void lookup_sort_schedule(uint32_t log_p, G&& global, G2&& global2, B&& block) {
    const uint64_t P = (uint64_t)1 << log_p;
    block(2, P < LOOKUP_SORT_BLOCK ? P : (uint64_t)LOOKUP_SORT_BLOCK, 0);     // the whole network up to k = block size (j_first 0: from k / 2)
    for (uint64_t k = 2 * (uint64_t)LOOKUP_SORT_BLOCK; k <= P; k <<= 1) {
        uint64_t j = k / 2;
        for (; j >= 2 * (uint64_t)LOOKUP_SORT_BLOCK; j >>= 2) {   // steps j and j / 2, both still >= the block size
            uint32_t log_j = 0;
            while (((uint64_t)1 << log_j) < j) ++log_j;
            global2(k, log_j);
        }
        if (j >= LOOKUP_SORT_BLOCK) global(k, j);
        block(k, k, LOOKUP_SORT_BLOCK / 2);
    }
}
inline uint32_t lookup_sort_log_p(uint64_t u) {
    uint32_t lp = LOOKUP_SORT_LOG_BLOCK;   // at least one block
    while (((uint64_t)1 << lp) < u) ++lp;
    return lp;
}

ZKB_HD int lookup_cmp(const Fr& x, const Fr& y) {  // canonical integers
    for (int k = 7; k >= 0; --k) {
        if (x.l[k] != y.l[k]) return x.l[k] < y.l[k] ? -1 : 1;
    }
    return 0;
}

// A'[i], first[i]
ZKB_HD void lookup_first_thread(const LookupArgs& a, uint64_t i) {
    if (i >= a.u) return;
    const uint32_t src = a.idx_in[i];
    fr_store2(a.out_in, i, fr_load2(a.input, src));
    uint32_t f = 1;
    if (i > 0) f = lookup_cmp(fr_load2(a.canon_in, src), fr_load2(a.canon_in, a.idx_in[i - 1])) != 0 ? 1u : 0u;
    a.first[i] = f;
}

// every first occurrence takes the first equal element of the sorted table; S'[i] = A'[i]
ZKB_HD void lookup_match_thread(const LookupArgs& a, uint64_t i) {
    if (i >= a.u || !a.first[i]) return;
    const Fr v = fr_load2(a.canon_in, a.idx_in[i]);
    uint64_t lo = 0, hi = a.u;   // lower_bound in the sorted table
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (lookup_cmp(fr_load2(a.canon_tab, a.idx_tab[mid]), v) < 0) lo = mid + 1;
        else hi = mid;
    }
    if (lo >= a.u || lookup_cmp(fr_load2(a.canon_tab, a.idx_tab[lo]), v) != 0) { *a.missing = 1; return; }
    a.consumed[lo] = 1;
    fr_store2(a.out_tab, i, fr_load2(a.input, a.idx_in[i]));
}

// repeated rows, ascending: rep_rows[rank] = row
ZKB_HD void lookup_rep_rows_thread(const LookupArgs& a, uint64_t i) {
    if (i >= a.u || a.first[i]) return;
    a.rep_rows[a.rep_rank[i]] = (uint32_t)i;
}

// the k-th leftover table element (ascending) goes to the k-th repeated row counted from the end
ZKB_HD void lookup_leftover_thread(const LookupArgs& a, uint64_t t) {
    if (t >= a.u || a.consumed[t]) return;
    const uint32_t k = a.left_rank[t];
    if (k >= a.repeated) { *a.missing = 1; return; }   // cannot happen when every first occurrence found its element
    fr_store2(a.out_tab, a.rep_rows[a.repeated - 1 - k], fr_load2(a.table, a.idx_tab[t]));
}

}  // namespace zkb
