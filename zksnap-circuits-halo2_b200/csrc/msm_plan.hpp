// msm_plan.hpp — host-side geometry of the Pippenger pipeline (shared by the CUDA driver and the CPU emulator).
#pragma once
#include <cstdint>
#include <cstdlib>

namespace zkb {

struct MsmGeometry {
    uint32_t c;            // window bits
    uint32_t nwin;         // W = ceil(255 / c) digit windows per scalar
    uint32_t bucket_sets;  // per column: W, or 1 when the SRS window table is used (all windows share one bucket set)
    uint32_t ncols;        // columns accumulated in one pass
    uint32_t total_sets;   // ncols * bucket_sets
    uint32_t nbuckets;     // total_sets << (c-1)
    uint32_t invalid_key;  // == nbuckets
    uint32_t key_bits;     // radix-sort bits covering [0, invalid_key]
    uint32_t chunk0;       // entries per thread, level 0
    uint32_t log_m;        // bucket-reduction segment length 2^log_m
};

// cost model: W * (n mixed adds + ~2.8 * 2^(c-1) full-add equivalents for the reduction)
inline uint32_t msm_pick_window(uint64_t n) {
    double best = 1e300;
    uint32_t best_c = 8;
    for (uint32_t c = 4; c <= 22; ++c) {
        uint32_t w = (255 + c - 1) / c;
        double cost = (double)w * ((double)n + 2.8 * (double)(1ull << (c - 1)));
        if (cost < best) { best = cost; best_c = c; }
    }
    return best_c;
}

// table mode: one reduction for all windows: W * n mixed adds + ~2.8 * 2^(c-1)
inline uint32_t msm_pick_window_table(uint64_t n) {
    double best = 1e300;
    uint32_t best_c = 8;
    for (uint32_t c = 4; c <= 22; ++c) {
        uint32_t w = (255 + c - 1) / c;
        double cost = (double)w * (double)n + 2.8 * (double)(1ull << (c - 1));
        if (cost < best) { best = cost; best_c = c; }
    }
    if (const char* e = getenv("ZKB_MSM_TABLE_C_DELTA")) {  // experiment hook: shift the table window width
        int v = (int)best_c + atoi(e);
        best_c = (uint32_t)(v < 4 ? 4 : (v > 22 ? 22 : v));
    }
    return best_c;
}

// Fan-in of one CTA-tree sum level over J elements per set (msm_sum_tree_kernel: `group` serial additions per thread, then a
// 7-step tree): the dependent chain is group + 7 per level, so spread the elements over as many CTAs as a single-CTA second
// level can absorb (128 x 128 elements -> group 1 twice: 16 additions deep; before: 8 + 7, then 1..8 + 7).
inline uint32_t msm_sum_tree_group(uint64_t J) {
    if (J <= 128) return 1;
    const uint64_t g = (J + 128ull * 128 - 1) / (128ull * 128);
    return (uint32_t)(g < 1 ? 1 : (g > 64 ? 64 : g));
}

// Launch plan of the accumulation levels (msm.cuh): level 0 cuts `entries` sorted entries into chunks of chunk0 per thread, 128
// threads per CTA; every CTA leaves two partial entries for the next level; a level that fits one CTA is the last.
// direct0: level 0 runs without the in-CTA tree (throughput regime: its threads write two partial entries each and the
// next, much smaller, level combines them — the tree would keep three of a CTA's four warps idle for ~2 % of a long chunk loop).
struct MsmAccPlan {
    uint32_t levels;
    uint32_t chunk[8];
    uint64_t ctas[8];
    bool direct0;
    uint64_t partials0;   // partial entries level 0 leaves
};
inline MsmAccPlan msm_acc_plan(uint64_t entries, uint32_t chunk0, bool direct0 = false) {
    MsmAccPlan p{};
    uint64_t count = entries;
    uint32_t chunk = chunk0 ? chunk0 : 1;
    while (count > 0 && p.levels < 8) {
        if (p.levels > 0) chunk = count <= 128ull * 16 ? (uint32_t)((count + 127) / 128) : 16u;
        const uint64_t threads = (count + chunk - 1) / chunk;
        const uint64_t ctas = (threads + 127) / 128;
        p.chunk[p.levels] = chunk;
        p.ctas[p.levels] = ctas;
        const bool direct = p.levels == 0 && direct0 && ctas > 1;
        if (p.levels == 0) { p.direct0 = direct; p.partials0 = direct ? 2 * ctas * 128 : 2 * ctas; }
        ++p.levels;
        if (ctas == 1) break;
        count = direct ? 2 * ctas * 128 : 2 * ctas;
    }
    return p;
}

// n = points per column.  For a batch of ncols columns the bucket array holds ncols * bucket_sets sets.
inline MsmGeometry msm_geometry(uint64_t n, uint32_t c_override = 0, uint32_t chunk_override = 0, bool table = false,
                                uint32_t ncols = 1) {
    MsmGeometry g{};
    g.ncols = ncols ? ncols : 1;
    g.c = c_override ? c_override : (table ? msm_pick_window_table(n ? n : 1) : msm_pick_window(n ? n : 1));
    if (g.c < 2) g.c = 2;
    if (g.c > 22) g.c = 22;
    g.nwin = (255 + g.c - 1) / g.c;
    // without the window table every window has its own bucket set: keep W x 2^(c-1) within the 24 key bits of one sort pass
    // (bucket_sort.cuh) — only c = 22 (12 x 2^21) exceeds them, at sizes where c = 21 costs < 4 % more additions
    while (!table && g.c > 2 && ((uint64_t)g.nwin << (g.c - 1)) > (1ull << 24)) { --g.c; g.nwin = (255 + g.c - 1) / g.c; }
    g.bucket_sets = table ? 1 : g.nwin;
    g.total_sets = g.bucket_sets * g.ncols;
    g.nbuckets = g.total_sets << (g.c - 1);
    g.invalid_key = g.nbuckets;
    g.key_bits = 1;
    while ((1ull << g.key_bits) <= g.invalid_key) ++g.key_bits;
    // level-0 chunk: 128 entries per thread when that still yields >= ~2 waves of threads (148 SMs x 512), shorter chunks
    // (shorter dependent chains, more threads) for small MSMs, which are otherwise latency bound
    uint32_t ch = 128;
    const uint64_t total = (uint64_t)g.nwin * (n ? n : 1) * g.ncols;
    while (ch > 16 && total / ch < 147456) ch >>= 1;
    g.chunk0 = chunk_override ? chunk_override : ch;
    // bucket reduction: ~2^15 segment threads keep the SMs busy while the per-thread chain (2m adds + the (c-1)-bit offset
    // multiplication) stays short; measured on B200 (profiles/r1_tuning.txt): 2^19 buckets -> m = 16, 2^21 -> m = 64.
    uint32_t lb = g.c - 1;
    uint32_t total_log = lb;  // log2 of all buckets of all sets (rounded down)
    while ((1ull << (total_log + 1)) <= ((uint64_t)g.total_sets << lb)) ++total_log;
    g.log_m = total_log > 18 ? (total_log - 15 > 7 ? 7 : total_log - 15) : (lb > 6 ? (total_log > 16 ? 3 : 2) : 0);
    if (g.log_m > lb) g.log_m = lb;
    if (const char* e = getenv("ZKB_MSM_LOG_M")) { uint32_t v = (uint32_t)atoi(e); if (v <= lb) g.log_m = v; }
    return g;
}

}  // namespace zkb
