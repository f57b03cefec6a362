// msm.cu — kernels and host driver of the G1 MSM (see msm.cuh for the algorithm).
#include <cstdlib>
#include <cstring>

#include "bucket_sort.cuh"
#include "host_fq64.hpp"
#include "msm_host.hpp"

namespace zkb {

// One scalar per thread.  The entries of a warp are written contiguously: warp-level exclusive scan of the per-thread
// counts, one atomicAdd per warp on the global entry counter.  STAGED: the warp's entries go through shared memory so the
// global stores are coalesced (nwin <= 24, i.e. every MSM of more than a few thousand points); otherwise each thread writes
// its own short run directly.
constexpr uint32_t MSM_DIGIT_STAGE_WINDOWS = 24;
template <bool STAGED, uint32_t CT_C>
__global__ void __launch_bounds__(256) msm_digits_kernel(const MsmDigitArgs a) {
    extern __shared__ uint32_t digit_smem[];  // STAGED: per warp nwin * 32 keys, then nwin * 32 values
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < a.n * a.ncols;
    Fr s = Fr::zero();
    if (live) s = fp_from_mont(fr_load2(a.scalars, t));
    uint32_t cnt = 0;
    if (live) msm_digits_foreach<CT_C>(a, t, s, [&](uint32_t, uint32_t) { ++cnt; });
    const uint32_t lane = threadIdx.x & 31;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && warp_total) base = atomicAdd(a.counter, (unsigned long long)warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint32_t off = incl - cnt;
    if (STAGED) {
        const uint32_t per_warp = a.nwin * 32;
        uint32_t* sk = digit_smem + (threadIdx.x >> 5) * 2 * per_warp;
        uint32_t* sv = sk + per_warp;
        if (live && cnt) msm_digits_foreach<CT_C>(a, t, s, [&](uint32_t k, uint32_t v) { sk[off] = k; sv[off] = v; ++off; });
        __syncwarp();
        for (uint32_t j = lane; j < warp_total; j += 32) {
            a.keys[base + j] = sk[j];
            a.vals[base + j] = sv[j];
        }
    } else {
        if (live && cnt) msm_digits_foreach<CT_C>(a, t, s, [&](uint32_t k, uint32_t v) { a.keys[base + off] = k; a.vals[base + off] = v; ++off; });
    }
}

// the tree phase is kept out of line: its full additions must not raise the register count of the level-0 chunk loop
__device__ __noinline__ void msm_acc_tree(const MsmAccArgs& a, uint32_t* skey, uint4* sval, uint4* scr) {
    for (uint32_t d = 1; d < MSM_ACC_CTA; d <<= 1) {
        const uint32_t pairs = MSM_ACC_CTA / (2 * d);
        for (uint32_t base = 0; base < pairs; base += COOP_GROUPS) {
            __syncthreads();   // the previous round's results (other warps) are in place
            for (uint32_t phase = 0; phase < COOP_PHASES; ++phase) {
                msm_acc_phase_combine(a, threadIdx.x, d, base, phase, MSM_ACC_CTA, skey, sval, scr);
                __syncwarp();  // the four lanes of an addition sit in one warp
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) msm_acc_phase_emit(a, blockIdx.x, skey, sval);
}

template <bool LEVEL0>
__global__ void __launch_bounds__(MSM_ACC_CTA, 4) msm_accumulate_kernel(const MsmAccArgs a0) {
    MsmAccArgs a = a0;
    if (LEVEL0 && a0.count_dev) a.count = *a0.count_dev;   // small commits: the digits kernel's count never visits the host
    __shared__ uint32_t skey[2 * MSM_ACC_CTA];
    __shared__ uint4 sval[2 * MSM_ACC_CTA * 8];
    __shared__ uint4 scr[COOP_GROUPS * COOP_SCRATCH_FQ * 2];
    msm_acc_phase_chunk<LEVEL0>(a, (uint64_t)blockIdx.x * MSM_ACC_CTA + threadIdx.x, threadIdx.x, skey, sval);
    msm_acc_tree(a, skey, sval, scr);
}

template <bool LEVEL0>
__global__ void __launch_bounds__(MSM_ACC_CTA) msm_accumulate_direct_kernel(const MsmAccArgs a0) {
    MsmAccArgs a = a0;
    if (LEVEL0 && a0.count_dev) a.count = *a0.count_dev;
    const uint64_t t = (uint64_t)blockIdx.x * MSM_ACC_CTA + threadIdx.x;
    if (t * a.chunk < a.count) msm_acc_thread_direct<LEVEL0>(a, t);
    else { a.pkeys_out[2 * t] = MSM_INVALID_KEY; a.pkeys_out[2 * t + 1] = MSM_INVALID_KEY; }
}

__global__ void __launch_bounds__(MSM_ACC_CTA) msm_sum_tree_kernel(const MsmSumTreeArgs a) {
    __shared__ uint4 sval[MSM_ACC_CTA * 8];
    __shared__ uint4 scr[COOP_GROUPS * COOP_SCRATCH_FQ * 2];
    msm_sum_tree_phase_load(a, blockIdx.x, threadIdx.x, sval);
    for (uint32_t d = MSM_ACC_CTA / 2; d >= 1; d >>= 1)
        for (uint32_t base = 0; base < d; base += COOP_GROUPS) {
            __syncthreads();
            for (uint32_t phase = 0; phase < COOP_PHASES; ++phase) {
                msm_sum_tree_phase_step(threadIdx.x, d, base, phase, sval, scr);
                __syncwarp();
            }
        }
    __syncthreads();
    if (threadIdx.x == 0) msm_sum_tree_phase_store(a, blockIdx.x, sval);
}

__global__ void __launch_bounds__(128) msm_merge_kernel(const MsmMergeArgs a) {
    msm_merge_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(128) msm_reduce_segment_kernel(const MsmReduceArgs a) {
    msm_reduce_segment_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// four lanes per segment, lock-step phases (msm.cuh): the latency-regime form of the segment sum
__global__ void __launch_bounds__(MSM_ACC_CTA) msm_reduce_segment_coop_kernel(const MsmReduceArgs a) {
    __shared__ uint4 state[COOP_GROUPS * 3 * 8];
    __shared__ uint4 scr[COOP_GROUPS * COOP_SCRATCH_FQ * 2];
    const uint32_t g = threadIdx.x >> 2, role = threadIdx.x & 3;
    const uint64_t t = (uint64_t)blockIdx.x * COOP_GROUPS + g;
    uint4* st = state + 24 * g;
    uint4* sc = scr + 2 * COOP_SCRATCH_FQ * g;
    msm_reduce_coop_init(role, st);
    __syncwarp();
    const uint32_t steps = msm_reduce_coop_steps(a);
    for (uint32_t s = 0; s < steps; ++s) {
        msm_reduce_coop_step(a, t, role, s, st, sc);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) msm_finalize_kernel(const MsmFinalArgs a) {
    msm_finalize_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(128) srs_table_kernel(const SrsTableArgs a) {
    srs_table_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(128) g1_fixed_base_mul_kernel(const FixedBaseArgs a) {
    g1_fixed_base_mul_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// ---- bucket sort (bucket_sort.cuh): phases separated by CTA barriers -------------------------------------------------------------
__global__ void bsort_init_kernel(const unsigned long long* count, uint32_t tile, uint32_t* seg_off, uint32_t* tile_start) {
    bsort_init_thread(count, tile, seg_off, tile_start);
}

__global__ void __launch_bounds__(128) bsort_tiles_kernel(const BsortArgs a) {
    bsort_tiles_thread(a, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(BSORT_THREADS) bsort_count_kernel(const BsortArgs a) {
    extern __shared__ uint32_t bsort_smem[];
    uint32_t* hist = bsort_smem;
    uint32_t* info = hist + (1u << a.bits);
    bsort_phase_begin(a, blockIdx.x, threadIdx.x, hist, info);
    __syncthreads();
    if (!info[2]) return;
    bsort_count_phase_hist(a, threadIdx.x, hist, info);
    __syncthreads();
    bsort_count_phase_flush(a, threadIdx.x, hist, info);
}

// exclusive prefix of v over the threads of the CTA (warp shuffles + one word per warp); two barriers
__device__ __forceinline__ uint32_t bsort_cta_scan(uint32_t v, uint32_t* wsum) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < BSORT_THREADS / 32 ? wsum[lane] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wi, d);
            if ((int)lane >= d) wi += u;
        }
        if (lane < BSORT_THREADS / 32) wsum[lane] = wi - w;
    }
    __syncthreads();
    return wsum[warp] + incl - v;
}

__global__ void __launch_bounds__(BSORT_THREADS) bsort_scan_kernel(const BsortArgs a) {
    __shared__ uint32_t part[2 * BSORT_THREADS];
    __shared__ uint32_t wsum[2][32];
    bsort_scan_phase_sum(a, blockIdx.x, threadIdx.x, part);
    const uint32_t base = bsort_cta_scan(part[threadIdx.x], wsum[0]);
    const uint32_t tbase = bsort_cta_scan(part[BSORT_THREADS + threadIdx.x], wsum[1]);
    bsort_scan_phase_write(a, blockIdx.x, threadIdx.x, base, tbase);
}

template <uint32_t KMAX>
__global__ void __launch_bounds__(BSORT_THREADS, (BSORT_ITEMS <= 8 ? 2048 : 1024) / BSORT_THREADS) bsort_scatter_kernel(const BsortArgs a) {
    extern __shared__ uint32_t bsort_smem[];
    const uint32_t bins = 1u << a.bits;
    uint32_t* hist = bsort_smem;
    uint32_t* delta = hist + bins;
    uint32_t* part = delta + bins;
    uint32_t* wsum = part + BSORT_THREADS;
    uint32_t* info = wsum + BSORT_GROUPS;
    uint32_t* skeys = info + 4;
    uint32_t* svals = skeys + a.tile;
    uint32_t rk[BSORT_ITEMS], rv[BSORT_ITEMS], g[KMAX];
    bsort_phase_begin(a, blockIdx.x, threadIdx.x, hist, info);
    __syncthreads();
    if (!info[2]) return;
    bsort_scatter_phase_rank(a, threadIdx.x, hist, info, rk, rv);
    __syncthreads();
    bsort_scatter_phase_sum(a, threadIdx.x, hist, part);
    const uint32_t base = bsort_cta_scan(part[threadIdx.x], wsum);
    bsort_scatter_phase_reserve<KMAX>(a, threadIdx.x, hist, delta, base, info, g);
    __syncthreads();
    bsort_scatter_phase_stage(a, threadIdx.x, hist, info, rk, rv, skeys, svals);
    bsort_scatter_phase_delta<KMAX>(a, threadIdx.x, delta, g);
    __syncthreads();
    bsort_scatter_phase_write(a, threadIdx.x, delta, info, skeys, svals);
}

struct BsortAttr { bool set = false; };
// Sorts the first *d_count (== valid, already known to the host) entries of (keys[0], vals[0]) by their low key_bits bits; *cur = the
// buffer pair that holds the result.
static int bsort_run(DevBuf* keys, DevBuf* vals, DevBuf& tmp, const unsigned long long* d_count, uint64_t valid, const BsortPlan& p,
                     cudaStream_t s, int* cur) {
    bool& attr = per_device<BsortAttr>().set;
    if (!attr) {
        BsortPlan big{};
        big.bits[0] = BSORT_MAX_BITS; big.tile = BSORT_MAX_TILE;
        ZKB_CUDA_TRY(cudaFuncSetAttribute(bsort_scatter_kernel<BSORT_KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsort_scatter_smem(big, 0)));
        ZKB_CUDA_TRY(cudaFuncSetAttribute(bsort_scatter_kernel<BSORT_KMAX_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsort_scatter_smem(big, 0)));
        attr = true;
    }
    uint32_t* t = tmp.as<uint32_t>();
    ZKB_CUDA_TRY(cudaMemsetAsync(t + p.zero_from, 0, p.zero_words * 4, s));
    bsort_init_kernel<<<1, 1, 0, s>>>(d_count, p.tile, t + p.off_seg1, t + p.off_tile1);
    count_launch();
    int in = 0;
    uint32_t shift = p.levels == 2 ? p.bits[1] : 0;
    for (uint32_t level = 0; level < p.levels; ++level) {
        BsortArgs a{};
        a.keys_in = keys[in].as<uint32_t>(); a.vals_in = vals[in].as<uint32_t>();
        a.keys_out = keys[in ^ 1].as<uint32_t>(); a.vals_out = vals[in ^ 1].as<uint32_t>();
        a.seg_off = t + (level == 0 ? p.off_seg1 : p.off_seg2);
        a.tile_start = t + (level == 0 ? p.off_tile1 : p.off_tile2);
        a.nseg = level == 0 ? 1u : 1u << p.bits[0];
        a.shift = shift; a.bits = p.bits[level]; a.tile = p.tile;
        a.cnt = t + (level == 0 ? p.off_cnt1 : p.off_cnt2);
        if (level == 0 && p.levels == 2) { a.next_seg_off = t + p.off_seg2; a.next_tile_start = t + p.off_tile2; a.next_tile = p.tile; }
        const unsigned tiles = (unsigned)bsort_max_tiles(p, level, valid);
        a.tile_info = reinterpret_cast<uint4*>(t + p.off_info);
        a.max_tiles = tiles;
        bsort_tiles_kernel<<<(tiles + 127) / 128, 128, 0, s>>>(a);
        bsort_count_kernel<<<tiles, BSORT_THREADS, bsort_count_smem(p, level), s>>>(a);
        bsort_scan_kernel<<<a.nseg, BSORT_THREADS, 0, s>>>(a);
        if (bsort_bins_per_thread(a.bits) <= BSORT_KMAX_NARROW) bsort_scatter_kernel<BSORT_KMAX_NARROW><<<tiles, BSORT_THREADS, bsort_scatter_smem(p, level), s>>>(a);
        else bsort_scatter_kernel<BSORT_KMAX><<<tiles, BSORT_THREADS, bsort_scatter_smem(p, level), s>>>(a);
        count_launch(4);
        ZKB_CUDA_TRY(cudaGetLastError());
        in ^= 1;
        shift = 0;
    }
    *cur = in;
    return ZKB_OK;
}

// Integer-pipe peak probe: 8 independent chains per thread of a_j = a_j * y + x (IMAD, multiplicand is the running
// value so ptxas cannot hoist the product).  A plain IMAD.WIDE.U32 issues at the same 2 cycles per warp instruction
// (tools/ubench.cu, profiles/r1_ubench.txt); the carry-chained IMAD.WIDE.U32.X costs 4.
__global__ void __launch_bounds__(256) imad_peak_kernel(uint64_t* out, uint32_t x, uint32_t y, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
    y |= 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(y), "r"(x));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;  // keeps the chains alive
}

int measure_imad_peak(double* macs_per_s) {
    Ctx& c = ctx();
    uint64_t* d = nullptr;
    ZKB_CUDA_TRY(cudaMalloc(&d, 8));
    const int iters = 4096, blocks = c.sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c.stream);
        imad_peak_kernel<<<blocks, threads, 0, c.stream>>>(d, 0x9e3779b9u, 0x85ebca6bu, iters);
        cudaEventRecord(e1, c.stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    count_launch(5);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    ZKB_CUDA_TRY(cudaGetLastError());
    *macs_per_s = (double)blocks * threads * iters * 32.0 / (best * 1e-3);
    return ZKB_OK;
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

int g1_fixed_base_mul_dev(const uint4* d_scalars, uint64_t n, uint4* d_out, cudaStream_t s) {
    if (n == 0) return ZKB_OK;
    FixedBaseArgs a{d_scalars, n, d_out};
    g1_fixed_base_mul_kernel<<<blocks_for(n, 128), 128, 0, s>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// rows 1..W-1 of the SRS window table: row w = 2^c * row (w-1); row 0 is a copy of the bases.  The running row is kept in
// XYZZ (scratch, n x 128 B) and each finished row is normalised to affine with one inversion per 16 points.
int srs_table_build(const uint4* d_bases, uint64_t n, uint32_t c, uint32_t nwin, uint4* d_table, cudaStream_t s) {
    ZKB_CUDA_TRY(cudaMemcpyAsync(d_table, d_bases, n * 64, cudaMemcpyDeviceToDevice, s));
    if (nwin < 2) return ZKB_OK;
    DevBuf scratch;
    ZKB_TRY(scratch.reserve(n * 128));
    int rc = ZKB_OK;
    for (uint32_t w = 1; w < nwin && rc == ZKB_OK; ++w) {
        SrsTableArgs a{d_bases, scratch.as<uint4>(), n, c, w == 1 ? 1u : 0u};
        srs_table_kernel<<<blocks_for(n, 128), 128, 0, s>>>(a);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { rc = ZKB_ERR_CUDA; break; }
        rc = g1_batch_to_affine_dev(scratch.as<uint4>(), n, d_table + 4 * n * (uint64_t)w, false, s);
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == ZKB_OK) rc = ZKB_ERR_CUDA;
    scratch.release();
    if (rc != ZKB_OK) { cudaGetLastError(); set_error("SRS window table build failed"); }
    return rc;
}

MsmWorkspace& msm_workspace() {
    MsmWorkspace& w = per_device<MsmWorkspace>();
    if (!w.env_read) {
        w.env_read = true;
        if (const char* e = getenv("ZKB_MSM_OVERLAP")) w.overlap = atoi(e);
    }
    return w;
}

void msm_release_workspace() {
    MsmWorkspace& w = msm_workspace();
    for (auto& b : w.keys) b.release();
    for (auto& b : w.vals) b.release();
    for (auto& b : w.pk) b.release();
    for (auto& b : w.pv) b.release();
    for (auto& b : w.seg) b.release();
    for (auto& b : w.sort_tmp) b.release();
    for (auto& b : w.counter) b.release();
    w.buckets.release();
    if (w.h_sums) { cudaFreeHost(w.h_sums); w.h_sums = nullptr; w.h_sums_cap = 0; }
    if (w.h_cnt) { cudaFreeHost(w.h_cnt); w.h_cnt = nullptr; }
    if (w.aux) { cudaStreamDestroy(w.aux); w.aux = nullptr; }
    for (int i = 0; i < 2; ++i) {
        if (w.ev_sorted[i]) { cudaEventDestroy(w.ev_sorted[i]); w.ev_sorted[i] = nullptr; }
        if (w.ev_acc[i]) { cudaEventDestroy(w.ev_acc[i]); w.ev_acc[i] = nullptr; }
        w.acc_pending[i] = false;
    }
}

// normalisation on the host: 64-bit arithmetic (host_fq64.hpp: ~8 us; the portable twins of the device chains take ~130 us)
static void xyzz_to_out(const XYZZ& p, uint64_t out[12]) { host64::to_out(host64::from_xyzz(p), out); }

void msm_identity_out(uint64_t out[12]) { xyzz_to_out(XYZZ::identity(), out); }

int msm_run(const uint4* d_scalars, const uint4* d_bases, uint64_t n, cudaStream_t s, uint64_t* out_jac,
            const MsmTable* table, uint32_t ncols, uint32_t phase, cudaEvent_t input_ready) {
    if (ncols == 0) return ZKB_OK;
    if (phase != MSM_WHOLE && (!table || ncols != 1 || n == 0)) { set_error("sliced MSM needs the SRS window table"); return ZKB_ERR_ARG; }
    if (n == 0) { for (uint32_t k = 0; k < ncols; ++k) msm_identity_out(out_jac + 12 * k); return ZKB_OK; }
    Ctx& c = ctx();
    MsmWorkspace& w = msm_workspace();
    const MsmGeometry g = table ? msm_geometry(n, table->c, c.msm_chunk_override, true, ncols)
                                : msm_geometry(n, c.msm_c_override, c.msm_chunk_override, false, ncols);
    const uint64_t total = (uint64_t)g.nwin * n * ncols;
    if (total >= (1ull << 31) || (uint64_t)g.total_sets << (g.c - 1) >= (1ull << 31)) {
        set_error("MSM batch too large for 32-bit sort keys: n=%zu cols=%u windows=%u", (size_t)n, ncols, g.nwin);
        return ZKB_ERR_ARG;
    }

    // ---- slices: digits + sort of this slice run on the aux stream (buffer set = slice parity) so that they overlap the previous
    // slice's accumulation on `s`; the accumulation waits for the sort through an event, and a buffer set is not rewritten before
    // the accumulation that read it (two slices ago) has finished.
    if (!w.h_cnt) ZKB_CUDA_TRY(cudaMallocHost(&w.h_cnt, 16));
    if (phase & MSM_FIRST) { w.slice_ix = 0; w.acc_pending[0] = w.acc_pending[1] = false; }
    const bool overlap = phase != MSM_WHOLE && w.overlap && input_ready;
    if (overlap && !w.aux) {
        ZKB_CUDA_TRY(cudaStreamCreateWithFlags(&w.aux, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            ZKB_CUDA_TRY(cudaEventCreateWithFlags(&w.ev_sorted[i], cudaEventDisableTiming));
            ZKB_CUDA_TRY(cudaEventCreateWithFlags(&w.ev_acc[i], cudaEventDisableTiming));
        }
    }
    const int set = overlap ? (int)(w.slice_ix++ & 1) : 0;
    cudaStream_t sp = overlap ? w.aux : s;   // the prepare stream
    DevBuf* const keys = &w.keys[2 * set];
    DevBuf* const vals = &w.vals[2 * set];
    // ---- workspace
    for (int i = 0; i < 2; ++i) {
        ZKB_TRY(keys[i].reserve(total * 4));
        ZKB_TRY(vals[i].reserve(total * 4));
    }
    // the library's own bucket sort (bucket_sort.cuh): keys of up to 24 bits — every single-column commit (W x 2^(c-1) <= 13 x 2^19) and
    // every group of a batch (the caller bounds its groups accordingly)
    uint32_t sort_kb = 1;  // keys are < nbuckets (zero digits are never emitted, so there is no "invalid" key to sort)
    while ((1ull << sort_kb) < g.nbuckets) ++sort_kb;
    const BsortPlan bp0 = bsort_plan(total, sort_kb);
    if (!bp0.levels) { set_error("MSM batch too wide for one sort pass: %u bucket-key bits (at most %u)", sort_kb, 2 * BSORT_MAX_BITS); return ZKB_ERR_ARG; }
    ZKB_TRY(w.sort_tmp[set].reserve(bp0.words * 4));
    ZKB_TRY(w.counter[set].reserve(8));
    if (overlap) {
        ZKB_CUDA_TRY(cudaStreamWaitEvent(sp, input_ready, 0));
        if (w.acc_pending[set]) ZKB_CUDA_TRY(cudaStreamWaitEvent(sp, w.ev_acc[set], 0));
    }
    const size_t bucket_bytes = (size_t)g.nbuckets * 128;
    ZKB_TRY(w.buckets.reserve(phase == MSM_WHOLE ? bucket_bytes : 2 * bucket_bytes));  // sliced: main + scratch array
    uint4* const bucket_main = w.buckets.as<uint4>();
    uint4* const bucket_acc = (phase & MSM_FIRST) ? bucket_main : bucket_main + 8 * (size_t)g.nbuckets;
    const uint64_t J0 = 1ull << (g.c - 1 - g.log_m);
    for (int i = 0; i < 2; ++i) ZKB_TRY(w.seg[i].reserve((size_t)g.total_sets * J0 * 128));
    const size_t host_bytes = (size_t)g.total_sets * 128 > (size_t)ncols * 96 ? (size_t)g.total_sets * 128 : (size_t)ncols * 96;
    if (w.h_sums_cap < host_bytes) {
        if (w.h_sums) cudaFreeHost(w.h_sums);
        w.h_sums = nullptr;
        w.h_sums_cap = 0;
        ZKB_CUDA_TRY(cudaMallocHost(&w.h_sums, host_bytes));
        w.h_sums_cap = host_bytes;
    }

    // ---- 1. digits
    {
        ProfScope prof("msm_digits", sp);
        MsmDigitArgs a{};
        a.scalars = d_scalars; a.n = n; a.ncols = ncols; a.sets_per_col = g.bucket_sets; a.c = g.c; a.nwin = g.nwin;
        a.keys = keys[0].as<uint32_t>(); a.vals = vals[0].as<uint32_t>();
        a.counter = reinterpret_cast<unsigned long long*>(w.counter[set].p);
        a.invalid_key = g.invalid_key;
        a.table_mode = table ? 1 : 0;
        a.row_stride = table ? table->row_stride : 0;
        ZKB_CUDA_TRY(cudaMemsetAsync(w.counter[set].p, 0, 8, sp));
        const unsigned nb = blocks_for(n * ncols, 256);
        const size_t sm = (size_t)g.nwin * 32 * 2 * 4 * 8;
        switch (g.nwin <= MSM_DIGIT_STAGE_WINDOWS ? g.c : 0u) {  // the window widths large MSMs use are compiled in
#define ZKB_DIGITS_CASE(C) case C: msm_digits_kernel<true, C><<<nb, 256, sm, sp>>>(a); break;
            ZKB_DIGITS_CASE(16) ZKB_DIGITS_CASE(17) ZKB_DIGITS_CASE(18) ZKB_DIGITS_CASE(19)
            ZKB_DIGITS_CASE(20) ZKB_DIGITS_CASE(21) ZKB_DIGITS_CASE(22)
#undef ZKB_DIGITS_CASE
            case 0: msm_digits_kernel<false, 0><<<nb, 256, 0, sp>>>(a); break;
            default: msm_digits_kernel<true, 0><<<nb, 256, sm, sp>>>(a); break;
        }
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
    }
    // The number of non-zero digits decides the size of everything downstream (a witness column has few).  Large commits read it
    // back: the sort's and the accumulation's launch plans are fitted to it (whole waves of equal chunks; a witness-like 2^24 column
    // has 29 M entries of 201 M possible).  Small commits (latency regime, <= 2^21 possible entries) never wait for it: every launch is
    // sized for the upper bound W n, the kernels read the count on the device, and it comes back with the result.
    const bool count_on_device = phase == MSM_WHOLE && total <= (1ull << 21);
    unsigned long long* const h_cnt = reinterpret_cast<unsigned long long*>(w.h_cnt) + set;
    ZKB_CUDA_TRY(cudaMemcpyAsync(h_cnt, w.counter[set].p, 8, cudaMemcpyDeviceToHost, sp));
    uint64_t valid = total;
    if (!count_on_device) {
        ZKB_CUDA_TRY(cudaStreamSynchronize(sp));   // overlap: waits for this slice's digits only, not for the accumulation on `s`
        valid = *h_cnt;
        w.last_entries = (phase & MSM_FIRST) ? valid : w.last_entries + valid;
    }
    // ---- 2. sort
    const uint32_t* sk;
    const uint32_t* sv;
    {
        ProfScope prof("msm_sort", sp);
        {
            static const uint32_t tile_override = getenv("ZKB_MSM_SORT_TILE") ? (uint32_t)atoi(getenv("ZKB_MSM_SORT_TILE")) : 0;
            static const uint32_t b1_override = getenv("ZKB_MSM_SORT_B1") ? (uint32_t)atoi(getenv("ZKB_MSM_SORT_B1")) : 0;   // experiment hooks
            const BsortPlan bp = bsort_plan(valid, sort_kb, tile_override, b1_override);
            ZKB_TRY(w.sort_tmp[set].reserve(bp.words * 4));   // a shorter tile can mean more tile descriptors than the first estimate
            int cur = 0;
            if (valid) ZKB_TRY(bsort_run(keys, vals, w.sort_tmp[set], reinterpret_cast<const unsigned long long*>(w.counter[set].p), valid, bp, sp, &cur));
            sk = keys[cur].as<uint32_t>();
            sv = vals[cur].as<uint32_t>();
        }
    }
    if (overlap) {
        ZKB_CUDA_TRY(cudaEventRecord(w.ev_sorted[set], sp));
        ZKB_CUDA_TRY(cudaStreamWaitEvent(s, w.ev_sorted[set], 0));
    }
    // level-0 chunk: all threads do the same amount of work, so the grid is cut to a whole number of waves (4 CTAs of 128
    // threads per SM) — a last wave that is 25 % full costs as much as a full one (2^19 points: 3.2 waves of chunk 32 -> 1 wave
    // of chunk 104)
    const uint64_t wave_threads = (uint64_t)c.sm_count * 4 * MSM_ACC_CTA;
    uint64_t waves = (valid + wave_threads * 128 - 1) / (wave_threads * 128);
    if (waves < 1) waves = 1;
    uint32_t chunk0 = c.msm_chunk_override ? c.msm_chunk_override : (uint32_t)((valid + waves * wave_threads - 1) / (waves * wave_threads));
    // shortest chains: 16 entries per thread, 8 for the smallest commits (2^13 points: 0.41 -> 0.39 ms; at 2^15 the extra partial entries already cost more)
    const uint32_t chunk_min = valid <= 200000 ? 8u : 16u;
    if (!c.msm_chunk_override) chunk0 = chunk0 < chunk_min ? chunk_min : (chunk0 > 128 ? 128 : chunk0);
    static const int tree_max_waves = getenv("ZKB_MSM_TREE_WAVES") ? atoi(getenv("ZKB_MSM_TREE_WAVES")) : 1;
    const MsmAccPlan plan = msm_acc_plan(valid, chunk0, waves > (uint64_t)tree_max_waves);
    for (int i = 0; i < 2; ++i) {
        ZKB_TRY(w.pk[i].reserve(plan.partials0 * 4 + 64));
        ZKB_TRY(w.pv[i].reserve(plan.partials0 * 128 + 128));
    }
    // ---- 3. accumulate (levels): every CTA leaves two partial entries, so the list shrinks by chunk x 64 per level
    {
        ProfScope prof("msm_accumulate", s);
        ZKB_CUDA_TRY(cudaMemsetAsync(bucket_acc, 0, bucket_bytes, s));
        uint64_t count = valid;
        int pp = 0;
        for (uint32_t level = 0; level < plan.levels; ++level) {
            MsmAccArgs a{};
            a.keys = level == 0 ? sk : w.pk[pp ^ 1].as<uint32_t>();
            a.vals = sv;
            a.bases = table ? table->rows : d_bases;
            a.pin = w.pv[pp ^ 1].as<uint4>();
            a.count = count;
            a.count_dev = (level == 0 && count_on_device) ? reinterpret_cast<const unsigned long long*>(w.counter[set].p) : nullptr;
            a.chunk = plan.chunk[level];
            a.invalid_key = g.invalid_key;
            a.last_level = level + 1 == plan.levels ? 1 : 0;
            a.buckets = bucket_acc;
            a.pkeys_out = w.pk[pp].as<uint32_t>();
            a.pvals_out = w.pv[pp].as<uint4>();
            if (level == 0 && plan.direct0) msm_accumulate_direct_kernel<true><<<(unsigned)plan.ctas[0], MSM_ACC_CTA, 0, s>>>(a);
            else if (level == 0) msm_accumulate_kernel<true><<<(unsigned)plan.ctas[level], MSM_ACC_CTA, 0, s>>>(a);
            else msm_accumulate_kernel<false><<<(unsigned)plan.ctas[level], MSM_ACC_CTA, 0, s>>>(a);
            count_launch();
            ZKB_CUDA_TRY(cudaGetLastError());
            count = (level == 0 && plan.direct0) ? plan.partials0 : 2 * plan.ctas[level];
            pp ^= 1;
        }
    }
    if (!(phase & MSM_FIRST)) {  // fold this slice's buckets into the main array
        ProfScope prof("msm_accumulate", s);
        MsmMergeArgs m{bucket_main, bucket_acc, g.nbuckets};
        msm_merge_kernel<<<blocks_for(g.nbuckets, 128), 128, 0, s>>>(m);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
    }
    if (overlap) {
        ZKB_CUDA_TRY(cudaEventRecord(w.ev_acc[set], s));
        w.acc_pending[set] = true;
    }
    if (!(phase & MSM_LAST)) return ZKB_OK;
    // ---- 4. bucket reduction: one weighted sum per bucket set
    const uint4* sums;
    int cur = 0;
    {
        ProfScope prof("msm_reduce", s);
        MsmReduceArgs r{};
        r.buckets = bucket_main; r.c = g.c; r.nwin = g.total_sets; r.log_m = g.log_m;
        r.seg_out = w.seg[0].as<uint4>();
        uint64_t J = J0;
        // few segments (latency regime): four lanes share a segment; many: one thread per segment (throughput)
        static const uint64_t coop_max = getenv("ZKB_MSM_REDUCE_COOP_MAX") ? strtoull(getenv("ZKB_MSM_REDUCE_COOP_MAX"), nullptr, 10) : (1ull << 14);   // 4 lanes x 2^14 = one wave of threads
        if ((uint64_t)g.total_sets * J <= coop_max)
            msm_reduce_segment_coop_kernel<<<blocks_for((uint64_t)g.total_sets * J, COOP_GROUPS), MSM_ACC_CTA, 0, s>>>(r);
        else
            msm_reduce_segment_kernel<<<blocks_for((uint64_t)g.total_sets * J, 128), 128, 0, s>>>(r);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        while (J > 1) {   // CTA-cooperative sums: one or two launches
            const uint32_t grp = msm_sum_tree_group(J);
            const uint64_t Jn = (J + (uint64_t)MSM_ACC_CTA * grp - 1) / ((uint64_t)MSM_ACC_CTA * grp);
            MsmSumTreeArgs sa{w.seg[cur].as<uint4>(), w.seg[cur ^ 1].as<uint4>(), J, Jn, grp};
            msm_sum_tree_kernel<<<(unsigned)((uint64_t)g.total_sets * Jn), MSM_ACC_CTA, 0, s>>>(sa);
            count_launch();
            ZKB_CUDA_TRY(cudaGetLastError());
            cur ^= 1;
            J = Jn;
        }
        sums = w.seg[cur].as<uint4>();
    }
    // ---- 5. fold the per-set sums and normalise: on the host for a few columns (north_star: "tiny bucket sums combined
    // on the host"), on the device for wide batches where hundreds of host inversions would dominate
    if (ncols >= 8) {
        ProfScope prof("msm_reduce", s);
        MsmFinalArgs f{sums, ncols, g.bucket_sets, g.c, w.seg[cur ^ 1].as<uint4>()};
        msm_finalize_kernel<<<blocks_for(ncols, 32), 32, 0, s>>>(f);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        ZKB_CUDA_TRY(cudaMemcpyAsync(w.h_sums, f.out, (size_t)ncols * 96, cudaMemcpyDeviceToHost, s));
        ZKB_CUDA_TRY(cudaStreamSynchronize(s));
        if (count_on_device) w.last_entries = *h_cnt;
        memcpy(out_jac, w.h_sums, (size_t)ncols * 96);
        return ZKB_OK;
    }
    ZKB_CUDA_TRY(cudaMemcpyAsync(w.h_sums, sums, (size_t)g.total_sets * 128, cudaMemcpyDeviceToHost, s));
    ZKB_CUDA_TRY(cudaStreamSynchronize(s));
    if (count_on_device) w.last_entries = *h_cnt;
    for (uint32_t col = 0; col < ncols; ++col) {
        XYZZ hs[MSM_MAX_WINDOWS];  // c >= 2 gives at most 128 window sums per column
        for (uint32_t i = 0; i < g.bucket_sets; ++i)
            hs[i] = XYZZ::load(reinterpret_cast<const uint4*>(w.h_sums) + 8 * ((size_t)col * g.bucket_sets + i));
        host64::to_out(host64::combine_windows(hs, g.bucket_sets, g.c), out_jac + 12 * col);
    }
    return ZKB_OK;
}

int g1_sum_host(const uint64_t* pts, size_t count, uint64_t out[12]) {
    host64::P acc = host64::identity();
    for (size_t i = 0; i < count; ++i) acc = host64::added(acc, host64::from_jacobian(pts + 12 * i));
    host64::to_out(acc, out);
    return ZKB_OK;
}

}  // namespace zkb

// Self-test entry of the bucket sort: n (key, value) pairs from host memory, sorted by the low key_bits bits of the key on the
// device and copied back (the order inside one key is unspecified).  Counts as test infrastructure like zkb_field_vec_op.
int zkb_bucket_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, uint32_t key_bits, uint32_t tile) {
    using namespace zkb;
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (n == 0) return ZKB_OK;
    if (!keys || !vals) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    if (n >= (1ull << 31)) { set_error("too many pairs"); return ZKB_ERR_ARG; }
    if (tile && (tile > BSORT_MAX_TILE || tile % BSORT_THREADS)) { set_error("tile must be a multiple of %u and at most %u", BSORT_THREADS, BSORT_MAX_TILE); return ZKB_ERR_ARG; }
    const BsortPlan p = bsort_plan(n, key_bits, tile);
    if (!p.levels) { set_error("keys of %u bits are too wide for the two-level bucket sort (at most %u)", key_bits, 2 * BSORT_MAX_BITS); return ZKB_ERR_ARG; }
    DevBuf k[2], v[2], tmp, cnt;
    auto done = [&](int rc) { for (auto& b : k) b.release(); for (auto& b : v) b.release(); tmp.release(); cnt.release(); return rc; };
    for (int i = 0; i < 2; ++i) {
        int rc = k[i].reserve(n * 4);
        if (rc == ZKB_OK) rc = v[i].reserve(n * 4);
        if (rc != ZKB_OK) return done(rc);
    }
    int rc = tmp.reserve(p.words * 4);
    if (rc == ZKB_OK) rc = cnt.reserve(8);
    if (rc != ZKB_OK) return done(rc);
    cudaStream_t s = ctx().stream;
    const unsigned long long n64 = n;
    int cur = 0;
    if (cudaMemcpyAsync(k[0].p, keys, n * 4, cudaMemcpyHostToDevice, s) != cudaSuccess || cudaMemcpyAsync(v[0].p, vals, n * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(cnt.p, &n64, 8, cudaMemcpyHostToDevice, s) != cudaSuccess) { set_error("upload failed"); return done(ZKB_ERR_CUDA); }
    rc = bsort_run(k, v, tmp, reinterpret_cast<const unsigned long long*>(cnt.p), n, p, s, &cur);
    if (rc != ZKB_OK) return done(rc);
    if (cudaMemcpyAsync(keys, k[cur].p, n * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaMemcpyAsync(vals, v[cur].p, n * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_error("bucket sort failed: %s", cudaGetErrorString(cudaGetLastError())); return done(ZKB_ERR_CUDA); }
    return done(ZKB_OK);
}
