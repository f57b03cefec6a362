// bucket_sort.cuh — the MSM's own sort of (bucket key, point index) pairs: a most-significant-digit-first radix PARTITION.
//
// What the accumulation needs (msm.cuh step 3) is weaker than a stable sort: all entries of one bucket contiguous, buckets in
// ascending order; the order INSIDE a bucket is irrelevant, because a bucket is a sum of group elements and the commitment is
// returned in its canonical encoding.  That allows a partition that never ranks entries against each other:
//
//   level 1  digit = the top b1 bits of the key.  The entry array is cut into tiles; a CTA counts its tile's digits in a
//            shared-memory histogram (`count`), one CTA turns the global histogram into segment offsets (`scan`), and each
//            CTA then reserves, per digit, ONE run in that digit's segment with a single global atomicAdd and moves its
//            entries there (`scatter`): positions inside a (tile, digit) run come from shared-memory atomics, the tile is
//            reordered in shared memory first so that consecutive lanes write consecutive addresses.
//   level 2  the same kernels on the remaining b2 = key_bits - b1 bits, with the tiles confined to one level-1 segment each
//            (`tiles`: tile -> segment by binary search in the scanned tile counts, one thread per tile); a segment's sub-bins
//            are laid out in order inside the segment, so after level 2 the array is sorted by the whole key.
//
// Keys of up to 12 bits take one level, up to 24 bits two: every commit of this library (with the window table 2^(c-1) <= 2^21 buckets
// per column; without it c is capped so that W x 2^(c-1) <= 2^24; batches are cut into groups that fit).  Any digit distribution is
// handled by construction: tiles are equal-sized pieces of the INPUT, a heavy bucket is just a long run that many tiles
// append to (witness columns put most entries into a few buckets; the all-equal column puts everything into W of them).
// Two passes over the pairs instead of the three of an 8-bit least-significant-digit sort, no ranking by warp-wide matching, and
// the entry count is read on the device (`bsort_init_thread`): the launches' geometry only needs an upper bound.
//
// Every kernel is a sequence of per-thread phase functions with a barrier in between, so the CPU emulator (hostemu.cu) runs
// exactly this code.
#pragma once
#include "field.cuh"

namespace zkb {

// CTA geometry (overridable for A/B builds).  Measured at 2^24 (profiles/r2_sort_ab.txt): 1024 threads x 8 entries — 32 registers, two
// CTAs = 64 warps per SM — 3.35 ms; 512 x 16 (64 registers, 32 warps per SM) 3.77 ms: the scatter waits on global loads, warps hide it.
#ifndef ZKB_BSORT_THREADS
#define ZKB_BSORT_THREADS 1024
#endif
#ifndef ZKB_BSORT_ITEMS
#define ZKB_BSORT_ITEMS 8
#endif
constexpr uint32_t BSORT_THREADS = ZKB_BSORT_THREADS;
constexpr uint32_t BSORT_ITEMS = ZKB_BSORT_ITEMS;                  // entries a scatter thread holds in registers
constexpr uint32_t BSORT_MAX_TILE = BSORT_THREADS * BSORT_ITEMS;   // entries per tile (8192)
constexpr uint32_t BSORT_MAX_BITS = 12;                            // digit width per level (<= 11: 2 bins per thread, 12: 4)
constexpr uint32_t BSORT_GROUP = 16;                               // per-thread partial sums scanned serially by one thread
constexpr uint32_t BSORT_GROUPS = BSORT_THREADS / BSORT_GROUP;     // 64

struct BsortArgs {
    const uint32_t* keys_in;
    const uint32_t* vals_in;
    uint32_t* keys_out;
    uint32_t* vals_out;
    const uint32_t* seg_off;      // [nseg + 1] first entry of every input segment (level 1: {0, count})
    const uint32_t* tile_start;   // [nseg + 1] first tile of every input segment
    uint4* tile_info;             // [max tiles] (segment, first entry, length, -) of every tile, written by `tiles`
    uint32_t max_tiles;
    uint32_t nseg;
    uint32_t shift, bits;         // digit = (key >> shift) & (2^bits - 1)
    uint32_t tile;                // entries per tile, <= BSORT_MAX_TILE
    uint32_t* cnt;                // [nseg << bits] digit counts; `scan` turns them into output cursors
    uint32_t* next_seg_off;       // scan, nseg == 1: [2^bits + 1] segment offsets of the next level (may be null)
    uint32_t* next_tile_start;    //                   [2^bits + 1] first tile of every next-level segment
    uint32_t next_tile;
};

ZKB_HD uint32_t bsort_atomic_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}

ZKB_HD uint32_t bsort_digit(const BsortArgs& a, uint32_t key) { return (key >> a.shift) & ((1u << a.bits) - 1u); }

// level-1 bookkeeping from the device-side entry count: one segment, ceil(count / tile) tiles
ZKB_HD void bsort_init_thread(const unsigned long long* count, uint32_t tile, uint32_t* seg_off, uint32_t* tile_start) {
    const uint32_t n = (uint32_t)*count;
    seg_off[0] = 0;
    seg_off[1] = n;
    tile_start[0] = 0;
    tile_start[1] = (n + tile - 1) / tile;
}

// tile t -> (segment, first entry, length), length 0 when t is past the last tile.  One thread per tile (`tiles` kernel): the
// binary search is a chain of dependent loads that the tile's CTA would otherwise wait for with every thread idle.
ZKB_HD void bsort_tiles_thread(const BsortArgs& a, uint32_t t) {
    if (t >= a.max_tiles) return;
    uint32_t info[3] = {0, 0, 0};
    if (t >= a.tile_start[a.nseg]) { a.tile_info[t] = make_uint4(0, 0, 0, 0); return; }
    uint32_t lo = 0, hi = a.nseg;   // tile_start[lo] <= t < tile_start[hi]; empty segments (equal starts) are skipped
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.tile_start[mid] <= t) lo = mid;
        else hi = mid;
    }
    const uint32_t begin = a.seg_off[lo] + (t - a.tile_start[lo]) * a.tile;
    const uint32_t left = a.seg_off[lo + 1] - begin;
    info[0] = lo;
    info[1] = begin;
    info[2] = left < a.tile ? left : a.tile;
    a.tile_info[t] = make_uint4(info[0], info[1], info[2], 0);
}

// ---- shared phases: clear the histogram, look the tile up -----------------------------------------------------------------------
ZKB_HD void bsort_phase_begin(const BsortArgs& a, uint32_t t, uint32_t tid, uint32_t* hist, uint32_t* info) {
    for (uint32_t b = tid; b < (1u << a.bits); b += BSORT_THREADS) hist[b] = 0;
    if (tid == 0) {
        const uint4 ti = a.tile_info[t];
        info[0] = ti.x; info[1] = ti.y; info[2] = ti.z;
    }
}

// ---- count ----------------------------------------------------------------------------------------------------------------------
// All loads of a batch are issued before the first atomic: an atomic orders the memory operations around it, so a loop of
// load -> atomic pays the global-memory latency once per ENTRY (measured: long-scoreboard stalls dominated both kernels).
// Out-of-range slots re-read the tile's last entry (always valid: len >= 1) so the loads need no branch.
constexpr uint32_t BSORT_COUNT_BATCH = 8;
ZKB_HD void bsort_count_phase_hist(const BsortArgs& a, uint32_t tid, uint32_t* hist, const uint32_t* info) {
    const uint32_t lo = info[1], len = info[2];
    if (!len) return;
    for (uint32_t first = 0; first < len; first += BSORT_COUNT_BATCH * BSORT_THREADS) {
        uint32_t k[BSORT_COUNT_BATCH];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t it = 0; it < BSORT_COUNT_BATCH; ++it) {
            const uint32_t i = first + it * BSORT_THREADS + tid;
            k[it] = a.keys_in[lo + (i < len ? i : len - 1)];
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t it = 0; it < BSORT_COUNT_BATCH; ++it)
            if (first + it * BSORT_THREADS + tid < len) bsort_atomic_add(hist + bsort_digit(a, k[it]), 1);
    }
}
ZKB_HD void bsort_count_phase_flush(const BsortArgs& a, uint32_t tid, const uint32_t* hist, const uint32_t* info) {
    if (!info[2]) return;
    uint32_t* cnt = a.cnt + ((size_t)info[0] << a.bits);
    for (uint32_t b = tid; b < (1u << a.bits); b += BSORT_THREADS)
        if (hist[b]) bsort_atomic_add(cnt + b, hist[b]);
}

// ---- exclusive scan of 2^bits values by one CTA: thread tid owns the bins [tid K, (tid + 1) K), K = ceil(2^bits / threads) ---------
// phase A: per-thread sums; phase B: the first BSORT_GROUPS threads scan their group of sums; phase C: thread 0 scans the group
// totals; then the caller walks its bins from bsort_scan_base().  `part` / `gsum`: BSORT_THREADS / BSORT_GROUPS words (x2 with tiles).
ZKB_HD uint32_t bsort_bins_per_thread(uint32_t bits) { return ((1u << bits) + BSORT_THREADS - 1) / BSORT_THREADS; }
ZKB_HD void bsort_scan_phase_groups(uint32_t tid, uint32_t* part, uint32_t* gsum) {
    if (tid >= BSORT_GROUPS) return;
    uint32_t run = 0;
    for (uint32_t j = 0; j < BSORT_GROUP; ++j) {
        const uint32_t v = part[tid * BSORT_GROUP + j];
        part[tid * BSORT_GROUP + j] = run;
        run += v;
    }
    gsum[tid] = run;
}
ZKB_HD void bsort_scan_phase_top(uint32_t tid, uint32_t* gsum) {
    if (tid != 0) return;
    uint32_t run = 0;
    for (uint32_t j = 0; j < BSORT_GROUPS; ++j) {
        const uint32_t v = gsum[j];
        gsum[j] = run;
        run += v;
    }
}
ZKB_HD uint32_t bsort_scan_base(uint32_t tid, const uint32_t* part, const uint32_t* gsum) { return gsum[tid / BSORT_GROUP] + part[tid]; }

// ---- scan kernel: CTA `seg` turns the counts of its segment into absolute output cursors (in place) -------------------------------
// smem: part[2 x BSORT_THREADS], gsum[2 x BSORT_GROUPS] (second halves: tiles of the next level)
ZKB_HD void bsort_scan_phase_sum(const BsortArgs& a, uint32_t seg, uint32_t tid, uint32_t* part) {
    const uint32_t bins = 1u << a.bits, K = bsort_bins_per_thread(a.bits);
    const uint32_t* cnt = a.cnt + ((size_t)seg << a.bits);
    uint32_t s = 0, tiles = 0;
    for (uint32_t b = tid * K; b < (tid + 1) * K && b < bins; ++b) {
        s += cnt[b];
        if (a.next_tile_start) tiles += (cnt[b] + a.next_tile - 1) / a.next_tile;
    }
    part[tid] = s;
    part[BSORT_THREADS + tid] = tiles;
}
// `base` / `tbase`: exclusive prefix over the threads of the sums phase_sum left in part[tid] / part[BSORT_THREADS + tid] (the kernel
// gets them from a warp-shuffle scan, the emulator from the serial phases above — same values)
ZKB_HD void bsort_scan_phase_write(const BsortArgs& a, uint32_t seg, uint32_t tid, uint32_t base, uint32_t tbase) {
    const uint32_t bins = 1u << a.bits, K = bsort_bins_per_thread(a.bits);
    uint32_t* cnt = a.cnt + ((size_t)seg << a.bits);
    uint32_t run = a.seg_off[seg] + base;
    uint32_t trun = tbase;
    for (uint32_t b = tid * K; b < (tid + 1) * K && b < bins; ++b) {
        const uint32_t c = cnt[b];
        cnt[b] = run;
        if (a.next_seg_off) a.next_seg_off[b] = run;
        if (a.next_tile_start) a.next_tile_start[b] = trun;
        run += c;
        if (a.next_tile_start) trun += (c + a.next_tile - 1) / a.next_tile;
        if (b + 1 == bins) {
            if (a.next_seg_off) a.next_seg_off[bins] = run;
            if (a.next_tile_start) a.next_tile_start[bins] = trun;
        }
    }
}

// ---- scatter --------------------------------------------------------------------------------------------------------------------
// smem: hist[2^bits], delta[2^bits], part[BSORT_THREADS], gsum[BSORT_GROUPS], info[4], skeys[tile], svals[tile]
// registers: rk / rv [BSORT_ITEMS] = the thread's entries; g [KMAX] = the reservations in flight
ZKB_HD void bsort_scatter_phase_rank(const BsortArgs& a, uint32_t tid, uint32_t* hist, const uint32_t* info, uint32_t* rk, uint32_t* rv) {
    const uint32_t lo = info[1], len = info[2];   // len >= 1 (the kernel returns on an empty tile)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t it = 0; it < BSORT_ITEMS; ++it) {   // every key load first (see bsort_count_phase_hist), branch-free
        const uint32_t i = it * BSORT_THREADS + tid;
        rk[it] = a.keys_in[lo + (i < len ? i : len - 1)];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t it = 0; it < BSORT_ITEMS; ++it)     // counts only: the position inside the run is handed out by the staging
        if (it * BSORT_THREADS + tid < len) bsort_atomic_add(hist + bsort_digit(a, rk[it]), 1);
    // the values are not needed before the staging: their loads are issued last and fly during the scan and the reservation
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t it = 0; it < BSORT_ITEMS; ++it) {
        const uint32_t i = it * BSORT_THREADS + tid;
        rv[it] = a.vals_in[lo + (i < len ? i : len - 1)];
    }
}
ZKB_HD void bsort_scatter_phase_sum(const BsortArgs& a, uint32_t tid, const uint32_t* hist, uint32_t* part) {
    const uint32_t bins = 1u << a.bits, K = bsort_bins_per_thread(a.bits);
    uint32_t s = 0;
    for (uint32_t b = tid * K; b < (tid + 1) * K && b < bins; ++b) s += hist[b];
    part[tid] = s;
}
// Local offsets replace the counts, and one global atomicAdd per non-empty digit reserves the tile's run in the digit's segment.
// The reservations are only ISSUED here (g stays in the thread's registers): staging does not need them, so their round trip
// to L2 hides behind it; delta[] starts as minus the local offset and bsort_scatter_phase_delta adds the reserved start afterwards.
// KMAX: bins per thread the instantiation covers (2 for digits of <= 11 bits — every table-mode commit — 4 for 12-bit digits)
constexpr uint32_t BSORT_KMAX = ((1u << BSORT_MAX_BITS) + BSORT_THREADS - 1) / BSORT_THREADS;
constexpr uint32_t BSORT_KMAX_NARROW = (BSORT_KMAX + 1) / 2;
template <uint32_t KMAX>
ZKB_HD void bsort_scatter_phase_reserve(const BsortArgs& a, uint32_t tid, uint32_t* hist, uint32_t* delta, uint32_t base, const uint32_t* info,
                                        uint32_t* g) {
    const uint32_t bins = 1u << a.bits, K = bsort_bins_per_thread(a.bits);
    uint32_t* cnt = a.cnt + ((size_t)info[0] << a.bits);
    uint32_t c[KMAX];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < KMAX; ++j) { const uint32_t b = tid * K + j; c[j] = (j < K && b < bins) ? hist[b] : 0; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < KMAX; ++j) g[j] = c[j] ? bsort_atomic_add(cnt + tid * K + j, c[j]) : 0;
    uint32_t run = base;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < KMAX; ++j) {
        const uint32_t b = tid * K + j;
        if (j < K && b < bins) { hist[b] = run; delta[b] = 0u - run; run += c[j]; }
    }
}
template <uint32_t KMAX>
ZKB_HD void bsort_scatter_phase_delta(const BsortArgs& a, uint32_t tid, uint32_t* delta, const uint32_t* g) {
    const uint32_t bins = 1u << a.bits, K = bsort_bins_per_thread(a.bits);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < KMAX; ++j) {
        const uint32_t b = tid * K + j;
        if (j < K && b < bins) delta[b] += g[j];   // reserved start - local offset (empty digits: never looked up)
    }
}
// hist holds the local offset of every digit's run: a second shared-memory atomic hands out the positions inside the run (no
// ranks kept in registers; the order inside a run is free)
ZKB_HD void bsort_scatter_phase_stage(const BsortArgs& a, uint32_t tid, uint32_t* hist, const uint32_t* info, const uint32_t* rk,
                                      const uint32_t* rv, uint32_t* skeys, uint32_t* svals) {
    const uint32_t len = info[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t it = 0; it < BSORT_ITEMS; ++it) {
        const uint32_t i = it * BSORT_THREADS + tid;
        if (i < len) {
            const uint32_t pos = bsort_atomic_add(hist + bsort_digit(a, rk[it]), 1);
            skeys[pos] = rk[it];
            svals[pos] = rv[it];
        }
    }
}
// (measured and dropped: (key, value) staged as one 8-byte word, atomics batched ahead of their stores — 2^24: 3.13 vs 3.04 ms)
ZKB_HD void bsort_scatter_phase_write(const BsortArgs& a, uint32_t tid, const uint32_t* delta, const uint32_t* info, const uint32_t* skeys,
                                      const uint32_t* svals) {
    const uint32_t len = info[2];
    for (uint32_t j = tid; j < len; j += BSORT_THREADS) {
        const uint32_t key = skeys[j];
        const uint32_t dst = j + delta[bsort_digit(a, key)];
        a.keys_out[dst] = key;
        a.vals_out[dst] = svals[j];
    }
}

// ---- host-side plan (shared with the emulator) -------------------------------------------------------------------------------------
struct BsortPlan {
    uint32_t levels;        // 1 or 2 (0: key too wide for this sort)
    uint32_t bits[2];       // digit widths, most significant level first
    uint32_t tile;
    // temp layout in 32-bit words
    size_t off_seg1, off_tile1, off_cnt1, off_seg2, off_tile2, off_cnt2, off_info, words;
    size_t zero_from, zero_words;   // the count arrays (contiguous) are cleared before every sort
};
inline uint64_t bsort_max_tiles(const BsortPlan& p, uint32_t level, uint64_t entries);
inline BsortPlan bsort_plan(uint64_t entries, uint32_t key_bits, uint32_t tile_override = 0, uint32_t b1_override = 0) {
    BsortPlan p{};
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 2 * BSORT_MAX_BITS) return p;
    if (key_bits <= BSORT_MAX_BITS) { p.levels = 1; p.bits[0] = key_bits; }
    else {
        p.levels = 2; p.bits[0] = key_bits / 2; p.bits[1] = key_bits - p.bits[0];   // measured: the narrower digit first (2^24: 3.83 vs 4.01 ms)
        if (b1_override && b1_override <= BSORT_MAX_BITS && b1_override < key_bits && key_bits - b1_override <= BSORT_MAX_BITS) { p.bits[0] = b1_override; p.bits[1] = key_bits - b1_override; }
    }
    // tile: the full 8192 entries when that still gives every SM several tiles, shorter for small sorts (latency regime)
    uint64_t t = (entries / 592 + BSORT_THREADS - 1) / BSORT_THREADS * BSORT_THREADS;
    if (t < BSORT_THREADS) t = BSORT_THREADS;
    if (t > BSORT_MAX_TILE) t = BSORT_MAX_TILE;
    p.tile = tile_override ? tile_override : (uint32_t)t;
    size_t o = 0;
    p.off_seg1 = o; o += 2;
    p.off_tile1 = o; o += 2;
    const size_t n1 = (size_t)1 << p.bits[0];
    p.off_seg2 = o; o += n1 + 1;
    p.off_tile2 = o; o += n1 + 1;
    o = (o + 3) & ~(size_t)3;
    p.zero_from = o;
    p.off_cnt1 = o; o += n1;
    p.off_cnt2 = o; o += p.levels == 2 ? (size_t)1 << key_bits : 0;
    p.zero_words = o - p.zero_from;
    o = (o + 3) & ~(size_t)3;
    p.off_info = o;   // uint4 per tile, 16-byte aligned
    o += 4 * (size_t)bsort_max_tiles(p, p.levels - 1, entries);
    p.words = o;
    return p;
}
// upper bound of the tiles of a level when at most `entries` entries are sorted
inline uint64_t bsort_max_tiles(const BsortPlan& p, uint32_t level, uint64_t entries) {
    const uint64_t t = (entries + p.tile - 1) / p.tile;
    return level == 0 ? (t ? t : 1) : t + ((uint64_t)1 << p.bits[0]);
}
inline size_t bsort_scatter_smem(const BsortPlan& p, uint32_t level) {
    return (((size_t)2 << p.bits[level]) + BSORT_THREADS + BSORT_GROUPS + 4 + 2 * (size_t)p.tile) * 4;
}
inline size_t bsort_count_smem(const BsortPlan& p, uint32_t level) { return (((size_t)1 << p.bits[level]) + 4) * 4; }

}  // namespace zkb
