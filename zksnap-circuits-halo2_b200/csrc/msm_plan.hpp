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
    uint32_t chunk_up;     // entries per thread, levels >= 1
    uint32_t last_max;     // a level with <= last_max entries is finished by one thread
    uint32_t log_m;        // bucket-reduction segment length 2^log_m
    uint32_t sum_group;    // tree fan-in of the segment sum
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
    return best_c;
}

// entries per thread at partial levels >= 1: long chunks while the level is large enough to be throughput bound (fewer
// levels, less total work), short ones when only a few thousand partials remain and the dependent chain of full additions
// (~4 us each) is what the MSM waits for — this tail was ~0.5 ms of a 1.3 ms 2^16-point MSM
inline uint32_t msm_chunk_up(uint64_t count, uint32_t chunk_up_large) { return count >= (1u << 20) ? chunk_up_large : 8; }

// n = points per column.  For a batch of ncols columns the bucket array holds ncols * bucket_sets sets.
inline MsmGeometry msm_geometry(uint64_t n, uint32_t c_override = 0, uint32_t chunk_override = 0, bool table = false,
                                uint32_t ncols = 1) {
    MsmGeometry g{};
    g.ncols = ncols ? ncols : 1;
    g.c = c_override ? c_override : (table ? msm_pick_window_table(n ? n : 1) : msm_pick_window(n ? n : 1));
    if (g.c < 2) g.c = 2;
    if (g.c > 22) g.c = 22;
    g.nwin = (255 + g.c - 1) / g.c;
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
    g.chunk_up = 32;
    g.last_max = 16;
    // bucket reduction: ~2^15 segment threads keep the SMs busy while the per-thread chain (2m adds + the (c-1)-bit offset
    // multiplication) stays short; measured on B200 (profiles/r1_tuning.txt): 2^19 buckets -> m = 16, 2^21 -> m = 64.
    uint32_t lb = g.c - 1;
    uint32_t total_log = lb;  // log2 of all buckets of all sets (rounded down)
    while ((1ull << (total_log + 1)) <= ((uint64_t)g.total_sets << lb)) ++total_log;
    g.log_m = total_log > 18 ? (total_log - 15 > 7 ? 7 : total_log - 15) : (lb > 6 ? 3 : 0);
    if (g.log_m > lb) g.log_m = lb;
    g.sum_group = 8;
    if (const char* e = getenv("ZKB_MSM_LOG_M")) { uint32_t v = (uint32_t)atoi(e); if (v <= lb) g.log_m = v; }
    if (const char* e = getenv("ZKB_MSM_SUM_GROUP")) { uint32_t v = (uint32_t)atoi(e); if (v >= 2) g.sum_group = v; }
    return g;
}

}  // namespace zkb
