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

// ---- the sort: (canonical value, row) records, block-wise bitonic network + merge-path merges --------------------------------------------
// The records are compared as the pair (256-bit value, original row), a TOTAL order, so the result is exactly what a stable sort by
// value gives and does not depend on the schedule.  P = 2^log_p >= u records (the padding carries the all-ones key, above every
// canonical value, and row 0xffffffff), structure-of-arrays so that every access is a full 16-byte or 4-byte vector.
//   1. every block of LOOKUP_SORT_BLOCK records is sorted ascending inside one CTA: the bitonic network on shared memory;
//   2. log2(P / block) merge passes, runs of length L pairwise into runs of 2 L (ping-pong buffers): a partition kernel finds, for
//      every output tile of LOOKUP_MERGE_TILE records, how many of the records before it come from the left run (merge path: one
//      binary search per tile boundary, all boundaries in parallel); a CTA then stages its tile's two input pieces in shared memory,
//      every thread finds its own split the same way and merges LOOKUP_MERGE_VT records serially.  The array is read and written
//      once per pass (a bitonic network over global memory needs ~3 passes per doubling at 2^22 records: measured 9.1 ms for the
//      whole permute_expression_pair against the merges' figure in DESIGN.md).
// Written by hand instead of calling a library sort: no library kernel is left anywhere in this package.
#ifndef ZKB_LOOKUP_SORT_LOG_BLOCK
#define ZKB_LOOKUP_SORT_LOG_BLOCK 10
#endif
constexpr uint32_t LOOKUP_SORT_LOG_BLOCK = ZKB_LOOKUP_SORT_LOG_BLOCK;
constexpr uint32_t LOOKUP_SORT_BLOCK = 1u << LOOKUP_SORT_LOG_BLOCK;   // records per bitonic block (36 bytes of shared memory each)
constexpr uint32_t LOOKUP_SORT_THREADS = LOOKUP_SORT_BLOCK / 2 < 1024 ? LOOKUP_SORT_BLOCK / 2 : 1024;   // pairs per step / threads
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
// block kernel phases: shared arrays slo / shi / srow of LOOKUP_SORT_BLOCK records; block b covers [b, b + 1) * LOOKUP_SORT_BLOCK
ZKB_HD void lookup_sort_block_load(const LookupSortArgs& a, uint64_t block, uint32_t tid, uint4* slo, uint4* shi, uint32_t* srow) {
    for (uint32_t r = tid; r < LOOKUP_SORT_BLOCK; r += LOOKUP_SORT_THREADS) {
        const uint64_t g = block * LOOKUP_SORT_BLOCK + r;
        slo[r] = a.klo[g]; shi[r] = a.khi[g]; srow[r] = a.row[g];
    }
}
ZKB_HD void lookup_sort_block_step(uint32_t tid, uint32_t k, uint32_t j, uint4* slo, uint4* shi, uint32_t* srow) {
    for (uint32_t t = tid; t < LOOKUP_SORT_BLOCK / 2; t += LOOKUP_SORT_THREADS) {
        uint64_t i, p;
        lookup_sort_pair(t, j, &i, &p);
        const bool up = (i & k) == 0;   // local index: at k = block size every block ends ascending
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
// ---- merge passes ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t LOOKUP_MERGE_TILE = 1024;     // output records per CTA (36 KB of shared memory)
constexpr uint32_t LOOKUP_MERGE_THREADS = 256;
constexpr uint32_t LOOKUP_MERGE_VT = LOOKUP_MERGE_TILE / LOOKUP_MERGE_THREADS;
struct LookupMergeArgs {
    const uint4* ilo; const uint4* ihi; const uint32_t* irow;   // input: sorted runs of `run` records
    uint4* olo; uint4* ohi; uint32_t* orow;                     // output: sorted runs of 2 * run
    uint64_t run;
    uint32_t log_p;
    uint32_t* part;   // [P / LOOKUP_MERGE_TILE] records taken from the LEFT run before every tile boundary
};
ZKB_HD bool lookup_merge_in_less(const LookupMergeArgs& a, uint64_t x, uint64_t y) {
    return lookup_rec_less(a.ilo[x], a.ihi[x], a.irow[x], a.ilo[y], a.ihi[y], a.irow[y]);
}
// merge path: of the first d records of merge(A, B), how many come from A (|A| = na, |B| = nb; A wins ties — there are none)
template <class Less>
ZKB_HD uint64_t lookup_merge_path(uint64_t d, uint64_t na, uint64_t nb, Less&& b_less_a) {
    uint64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (!b_less_a(d - 1 - mid, mid)) lo = mid + 1;   // A[mid] <= B[d - 1 - mid]: A[mid] is among the first d
        else hi = mid;
    }
    return lo;
}
ZKB_HD void lookup_merge_partition_thread(const LookupMergeArgs& a, uint64_t c) {
    const uint64_t P = (uint64_t)1 << a.log_p;
    if (c >= P / LOOKUP_MERGE_TILE) return;
    const uint64_t g = c * LOOKUP_MERGE_TILE, base = g & ~(2 * a.run - 1), d = g - base;
    a.part[c] = (uint32_t)lookup_merge_path(d, a.run, a.run, [&](uint64_t y, uint64_t x) { return lookup_merge_in_less(a, base + a.run + y, base + x); });
}
// the two input pieces of tile c: A[a0, a0 + na) and B[b0, b0 + nb) of the pair of runs starting at `base`
ZKB_HD void lookup_merge_tile_geometry(const LookupMergeArgs& a, uint64_t c, uint64_t* base, uint64_t* a0, uint64_t* b0, uint32_t* na, uint32_t* nb) {
    const uint64_t g = c * LOOKUP_MERGE_TILE;
    *base = g & ~(2 * a.run - 1);
    const uint64_t d = g - *base;
    *a0 = a.part[c];
    const uint64_t a1 = (d + LOOKUP_MERGE_TILE == 2 * a.run) ? a.run : a.part[c + 1];
    *b0 = d - *a0;
    *na = (uint32_t)(a1 - *a0);
    *nb = LOOKUP_MERGE_TILE - *na;
}
ZKB_HD void lookup_merge_tile_load(const LookupMergeArgs& a, uint64_t c, uint32_t tid, uint4* slo, uint4* shi, uint32_t* srow) {
    uint64_t base, a0, b0;
    uint32_t na, nb;
    lookup_merge_tile_geometry(a, c, &base, &a0, &b0, &na, &nb);
    for (uint32_t r = tid; r < LOOKUP_MERGE_TILE; r += LOOKUP_MERGE_THREADS) {
        const uint64_t src = r < na ? base + a0 + r : base + a.run + b0 + (r - na);
        slo[r] = a.ilo[src]; shi[r] = a.ihi[src]; srow[r] = a.irow[src];
    }
}
ZKB_HD void lookup_merge_tile_merge(const LookupMergeArgs& a, uint64_t c, uint32_t tid, const uint4* slo, const uint4* shi, const uint32_t* srow) {
    uint64_t base, a0, b0;
    uint32_t na, nb;
    lookup_merge_tile_geometry(a, c, &base, &a0, &b0, &na, &nb);
    auto less = [&](uint32_t x, uint32_t y) { return lookup_rec_less(slo[x], shi[x], srow[x], slo[y], shi[y], srow[y]); };
    const uint32_t d = tid * LOOKUP_MERGE_VT;
    uint32_t ta = (uint32_t)lookup_merge_path(d, na, nb, [&](uint64_t y, uint64_t x) { return less(na + (uint32_t)y, (uint32_t)x); });
    uint32_t tb = d - ta;
    const uint64_t out = c * LOOKUP_MERGE_TILE + d;
    for (uint32_t e = 0; e < LOOKUP_MERGE_VT; ++e) {
        const bool take_a = tb >= nb || (ta < na && !less(na + tb, ta));
        const uint32_t src = take_a ? ta++ : na + tb++;
        a.olo[out + e] = slo[src]; a.ohi[out + e] = shi[src]; a.orow[out + e] = srow[src];
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
