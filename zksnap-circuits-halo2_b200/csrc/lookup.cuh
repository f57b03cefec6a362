// lookup.cuh — halo2's lookup argument, prover side: permute_expression_pair (plonk/lookup/prover.rs, un-vendored halo2-axiom;
// reached from create_proof, /root/reference/aggregator/src/wrapper.rs:129-137).  SURVEY.md §8f row 3.
//
// Given the compressed input expression A and table expression S on the usable rows, upstream produces
//   A' = A sorted ascending (Fr's Ord: canonical integer value),
//   S' with S'[i] = A'[i] on every row where A'[i] differs from A'[i-1] (first occurrence of a value; one copy of that value
//      is taken out of the table multiset — missing value = ConstraintSystemFailure), and the remaining table elements,
//      in ascending order, assigned to the repeated rows from the LAST repeated row backwards (`repeated_input_rows.pop()`).
// On the device: both columns are sorted by canonical value (radix sort on the four 64-bit limbs, least significant first),
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

// 64-bit limb `limb` of canon[idx[i]] -> keys[i]  (one LSD radix pass sorts by it)
ZKB_HD void lookup_gather_limb_thread(const uint4* canon, const uint32_t* idx, uint64_t u, uint32_t limb, unsigned long long* keys, uint64_t i) {
    if (i >= u) return;
    const Fr v = fr_load2(canon, idx[i]);
    keys[i] = (unsigned long long)v.l[2 * limb] | ((unsigned long long)v.l[2 * limb + 1] << 32);
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
