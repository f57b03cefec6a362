// poly.cu — polynomials resident in HBM (handles) and the Fr vector kernels of poly.cuh.
//
// The prover's polynomials (advice columns, permutation / lookup products, the quotient pieces) are committed, moved
// between bases and evaluated several times each.  Keeping them in HBM under a handle turns
//   commit_lagrange -> lagrange_to_coeff -> coeff_to_extended -> eval_polynomial -> kate_division -> commit
// into device-only steps: one upload per polynomial, 96-byte commitments and 32-byte evaluations coming back
// (coset-resident pipeline, eval / division helpers, proving-key residency).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msm_host.hpp"
#include "ntt_host.hpp"

#include "graph.cuh"
#include "lookup.cuh"
#include "poly.cuh"

namespace zkb {

__global__ void __launch_bounds__(128) poly_eval_chunk_kernel(const PolyEvalArgs a) {
    poly_eval_chunk_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) kate_expand_kernel(const KateExpandArgs a) {
    kate_expand_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) fr_batch_invert_kernel(const BatchInvertArgs a) {
    fr_batch_invert_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// element-wise field op on the real PTX carry chains: out[i] = a[i] (op) b[i]; field 0 Fr / 1 Fq; op 0 mul, 1 add, 2 sub, 3 sqr,
// 4 neg, 5 double, 6 from_mont(a), 7-9 canon(mul/add/sub lazy), 10 raw lazy product.  A self-test hook (the CPU emulator runs the portable twins of the chains, not the PTX).
template <class P>
__device__ __forceinline__ Fp<P> field_op(int op, const Fp<P>& x, const Fp<P>& y) {
    switch (op) {
        case 0: return fp_mul(x, y);
        case 1: return fp_add(x, y);
        case 2: return fp_sub(x, y);
        case 3: return fp_sqr(x);
        case 4: return fp_neg(x);
        case 5: return fp_dbl(x);
        case 6: return fp_from_mont(x);
        case 7: return fp_canon(fp_mul_lazy(x, y));   // lazy forms: inputs < 2M, result made canonical for comparison
        case 8: return fp_canon(fp_add_lazy(x, y));
        case 9: return fp_canon(fp_sub_lazy(x, y));
        default: return fp_mul_lazy(x, y);            // raw lazy product (must be < 2M)
    }
}
__global__ void __launch_bounds__(128) field_vec_op_kernel(int field, int op, const uint4* a, const uint4* b, uint4* out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (field == 0) field_op<FrParams>(op, Fr::load(a + 2 * i), Fr::load(b + 2 * i)).store(out + 2 * i);
    else field_op<FqParams>(op, Fq::load(a + 2 * i), Fq::load(b + 2 * i)).store(out + 2 * i);
}

__global__ void __launch_bounds__(128) poly_mul_kernel(const PolyMulArgs a) {
    poly_mul_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) poly_mul_periodic_kernel(const PolyMulPeriodicArgs a) {
    poly_mul_periodic_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) poly_scale_add_kernel(const PolyScaleAddArgs a) {
    poly_scale_add_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void poly_add_const_kernel(uint4* a, Fr c) {  // a[0] += c
    fr_store2(a, 0, fp_add(fr_load2(a, 0), c));
}
__global__ void __launch_bounds__(128) scan_chunk_product_kernel(const ScanChunkArgs a) {
    scan_chunk_product_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) scan_expand_kernel(const ScanExpandArgs a) {
    scan_expand_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// one thread per row of the extended domain; dynamic shared memory: slot planes (nslots x 2 x blockDim uint4), then the program
__global__ void __launch_bounds__(128) graph_evaluate_kernel(const GraphArgs g, uint32_t nslots, uint32_t stage_prog) {
    extern __shared__ uint4 graph_smem[];
    uint4* slots = graph_smem;
    const uint4* prog = g.prog;
    if (stage_prog) {
        uint4* sp = graph_smem + (size_t)2 * nslots * blockDim.x;
        for (uint32_t i = threadIdx.x; i < g.ninstr; i += blockDim.x) sp[i] = g.prog[i];
        __syncthreads();
        prog = sp;
    }
    graph_eval_thread(g, prog, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, slots, blockDim.x, threadIdx.x);
}

__global__ void __launch_bounds__(128) lookup_canon_kernel(const LookupArgs a) { lookup_canon_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(128) lookup_sort_init_kernel(const LookupSortArgs a) { lookup_sort_init_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
// the bitonic network up to k = block size on one block in shared memory: every block ends sorted ascending
__global__ void __launch_bounds__(LOOKUP_SORT_THREADS) lookup_sort_block_kernel(const LookupSortArgs a) {
    extern __shared__ uint4 lookup_sort_smem[];
    uint4* slo = lookup_sort_smem;
    uint4* shi = slo + LOOKUP_SORT_BLOCK;
    uint32_t* srow = reinterpret_cast<uint32_t*>(shi + LOOKUP_SORT_BLOCK);
    lookup_sort_block_load(a, blockIdx.x, threadIdx.x, slo, shi, srow);
    for (uint32_t k = 2; k <= LOOKUP_SORT_BLOCK; k <<= 1)
        for (uint32_t j = k / 2; j >= 1; j >>= 1) {
            __syncthreads();
            lookup_sort_block_step(threadIdx.x, k, j, slo, shi, srow);
        }
    __syncthreads();
    lookup_sort_block_store(a, blockIdx.x, threadIdx.x, slo, shi, srow);
}
__global__ void __launch_bounds__(128) lookup_merge_partition_kernel(const LookupMergeArgs a) {
    lookup_merge_partition_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(LOOKUP_MERGE_THREADS) lookup_merge_tile_kernel(const LookupMergeArgs a) {
    __shared__ uint4 slo[LOOKUP_MERGE_TILE];
    __shared__ uint4 shi[LOOKUP_MERGE_TILE];
    __shared__ uint32_t srow[LOOKUP_MERGE_TILE];
    lookup_merge_tile_load(a, blockIdx.x, threadIdx.x, slo, shi, srow);
    __syncthreads();
    lookup_merge_tile_merge(a, blockIdx.x, threadIdx.x, slo, shi, srow);
}
__global__ void __launch_bounds__(128) lookup_take_rows_kernel(const uint32_t* row, uint32_t* idx, uint64_t u) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < u) idx[i] = row[i];
}
// ---- exclusive scan of 32-bit counts (the ranks of the lookup argument): tiles of 4096, tile sums scanned by one CTA (u <= 2^32 / ...) ---
constexpr uint32_t SCAN_THREADS = 1024, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
__device__ __forceinline__ uint32_t scan_cta_exclusive(uint32_t v, uint32_t* wsum, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += x;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wsum[lane];   // SCAN_THREADS / 32 == 32 warps
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, wi, d);
            if ((int)lane >= d) wi += x;
        }
        wsum[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return wsum[warp] + incl - v;
}
// phase 0: out = exclusive scan inside every tile, sums[tile] = tile total; phase 1 (one CTA): sums -> exclusive scan of sums, tile by
// tile; phase 2: out += sums[tile]
__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_kernel(const uint32_t* in, uint32_t* out, uint32_t* sums, uint64_t n, int phase) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t total;
    if (phase == 0) {
        const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (uint32_t i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
        uint32_t run = scan_cta_exclusive(s, wsum, &total);
#pragma unroll
        for (uint32_t i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
        if (threadIdx.x == 0) sums[blockIdx.x] = total;
    } else if (phase == 1) {
        uint32_t carry = 0;   // n = number of tiles here
        for (uint64_t first = 0; first < n; first += SCAN_THREADS) {
            const uint64_t i = first + threadIdx.x;
            const uint32_t v = i < n ? sums[i] : 0;
            const uint32_t ex = scan_cta_exclusive(v, wsum, &total);
            if (i < n) sums[i] = carry + ex;
            carry += total;
            __syncthreads();
        }
    } else {
        const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
        const uint32_t add = sums[blockIdx.x];
#pragma unroll
        for (uint32_t i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n) out[base + i] += add;
    }
}
static int scan_u32_exclusive_dev(const uint32_t* in, uint32_t* out, uint64_t n, DevBuf& tmp, cudaStream_t s) {
    const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    ZKB_TRY(tmp.reserve(tiles * 4 + 16));
    uint32_t* sums = tmp.as<uint32_t>();
    scan_u32_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, sums, n, 0);
    if (tiles > 1) {
        scan_u32_kernel<<<1, SCAN_THREADS, 0, s>>>(nullptr, nullptr, sums, tiles, 1);
        scan_u32_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(nullptr, out, sums, n, 2);
        count_launch(2);
    }
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}
__global__ void __launch_bounds__(128) lookup_first_kernel(const LookupArgs a) { lookup_first_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(128) lookup_match_kernel(const LookupArgs a) { lookup_match_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(128) lookup_rep_rows_kernel(const LookupArgs a) { lookup_rep_rows_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(128) lookup_leftover_kernel(const LookupArgs a) { lookup_leftover_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(128) lookup_not_kernel(const uint32_t* in, uint32_t* out, uint64_t u) {   // out = 1 - in
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < u) out[i] = 1u - in[i];
}

static inline unsigned nblk(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
static inline uint64_t chunks(uint64_t n) { return (n + POLY_CHUNK - 1) / POLY_CHUNK; }

static Fr fr_of(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}
static void words_of(const Fr& v, uint32_t (&o)[8]) {
    for (int i = 0; i < 8; ++i) o[i] = v.l[i];
}
static Fr pow_chunk(Fr x) {  // x^POLY_CHUNK, POLY_CHUNK = 2^6
    for (uint32_t c = POLY_CHUNK; c > 1; c >>= 1) x = fp_sqr(x);
    return x;
}

struct PolyWs {
    DevBuf tmp;
    void* h_out = nullptr;  // pinned 32 B
};
struct GraphWsBuf : DevBuf {};
static DevBuf& graph_ws() { return per_device<GraphWsBuf>(); }
struct GraphRing {   // upload buffers of zkb_graph_evaluate_dev
    DevBuf buf[4];
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned next = 0;
};
static GraphRing& graph_ring() { return per_device<GraphRing>(); }
static PolyWs& poly_ws() { return per_device<PolyWs>(); }

// sum_i a[i] x^i -> d_out (one Fr on the device)
static int poly_eval_dev(const uint4* d_a, uint64_t n, Fr x, uint4* d_tmp, uint4** d_result, cudaStream_t s) {
    const uint4* cur = d_a;
    uint4* out = d_tmp;
    for (;;) {
        const uint64_t t = chunks(n);
        PolyEvalArgs a{};
        a.a = cur; a.n = n; a.out = out;
        words_of(x, a.x);
        poly_eval_chunk_kernel<<<nblk(t, 128), 128, 0, s>>>(a);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        if (t == 1) { *d_result = out; return ZKB_OK; }
        cur = out;
        out += 2 * t;
        n = t;
        x = pow_chunk(x);
    }
}

// Q_j = sum_{i>j} a_i b^(i-j-1) for j = 0..n-1 (Q_{n-1} = 0) -> d_q; d_tmp needs 2 * (chunks(n) + chunks(chunks(n)) + ...) Fr
static int kate_q_dev(const uint4* d_a, uint64_t n, Fr b, uint4* d_q, uint4* d_tmp, cudaStream_t s) {
    const uint64_t t = chunks(n);
    const uint4* carry = nullptr;
    if (t > 1) {
        uint4* S = d_tmp;
        uint4* K = d_tmp + 2 * t;
        PolyEvalArgs e{};
        e.a = d_a; e.n = n; e.out = S;
        words_of(b, e.x);
        poly_eval_chunk_kernel<<<nblk(t, 128), 128, 0, s>>>(e);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        ZKB_TRY(kate_q_dev(S, t, pow_chunk(b), K, d_tmp + 4 * t, s));
        carry = K;
    }
    KateExpandArgs x{};
    x.a = d_a; x.n = n; x.carry = carry; x.q = d_q;
    words_of(b, x.b);
    kate_expand_kernel<<<nblk(t, 128), 128, 0, s>>>(x);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// z[i] = prod_{j < i} v[j]; d_z may alias d_v; d_tmp as for kate_q_dev
static int prefix_product_dev(const uint4* d_v, uint64_t n, uint4* d_z, uint4* d_tmp, cudaStream_t s) {
    const uint64_t t = chunks(n);
    const uint4* carry = nullptr;
    if (t > 1) {
        uint4* P = d_tmp;          // chunk products
        uint4* C = d_tmp + 2 * t;  // their exclusive prefix products
        ScanChunkArgs c{d_v, n, P};
        scan_chunk_product_kernel<<<nblk(t, 128), 128, 0, s>>>(c);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        ZKB_TRY(prefix_product_dev(P, t, C, d_tmp + 4 * t, s));
        carry = C;
    }
    ScanExpandArgs x{d_v, n, carry, d_z};
    scan_expand_kernel<<<nblk(t, 128), 128, 0, s>>>(x);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

static size_t poly_tmp_bytes(uint64_t n) {  // generous: 4 Fr per chunk on every level
    size_t total = 0;
    for (uint64_t t = chunks(n); ; t = chunks(t)) {
        total += 4 * t * 32;
        if (t <= 1) break;
    }
    return total + 256;
}

// ---- registry ----------------------------------------------------------------------------------------------------------------
struct Poly {
    DevBuf buf;
    uint64_t n = 0;
    int slot = 0;   // device slot whose HBM holds it (the slot of the thread that created it)
};
static std::map<uint64_t, Poly*>& poly_map() {
    static std::map<uint64_t, Poly*> m;
    return m;
}
static uint64_t g_next_poly = 1;
static std::mutex g_poly_mu;   // the registry is global (handles are unique across devices); entry points lock only their own device

// Buffer pool.  The prover allocates and frees polynomials of two or three sizes (2^k, 2^extended_k) all the time, and
// cudaMalloc / cudaFree of half a gigabyte cost a device synchronisation each.  Freed buffers are kept (up to a quarter of the
// device memory, ZKB_POLY_POOL_MB overrides, 0 disables) and handed out again; every polynomial op runs on the library stream,
// so reuse is ordered after the last use.  Any failed allocation in the library empties the pool and retries (DevBuf::reserve).
struct PolyPool {
    std::multimap<size_t, void*> free;  // capacity -> buffer
    size_t bytes = 0;
    size_t limit = ~(size_t)0;          // resolved on first use
};
static PolyPool& poly_pool() { return per_device<PolyPool>(); }
void poly_pool_flush() {
    PolyPool& pool = poly_pool();
    for (auto& kv : pool.free) cudaFree(kv.second);
    pool.free.clear();
    pool.bytes = 0;
}
static size_t poly_pool_limit() {
    PolyPool& pool = poly_pool();
    if (pool.limit == ~(size_t)0) {
        const char* e = getenv("ZKB_POLY_POOL_MB");
        size_t fr = 0, total = 0;
        if (e) pool.limit = (size_t)strtoull(e, nullptr, 10) << 20;
        else pool.limit = cudaMemGetInfo(&fr, &total) == cudaSuccess ? total / 4 : 0;
    }
    return pool.limit;
}
static int pool_alloc(DevBuf& b, size_t bytes) {
    PolyPool& pool = poly_pool();
    auto it = pool.free.lower_bound(bytes);
    if (it != pool.free.end() && it->first <= bytes + bytes / 4 + 4096) {
        b.p = it->second;
        b.cap = it->first;
        pool.bytes -= it->first;
        pool.free.erase(it);
        return ZKB_OK;
    }
    return b.reserve(bytes);
}
static void pool_release(DevBuf& b) {
    PolyPool& pool = poly_pool();
    if (b.p && pool.bytes + b.cap <= poly_pool_limit()) {
        pool.free.emplace(b.cap, b.p);
        pool.bytes += b.cap;
        b.p = nullptr;
        b.cap = 0;
        return;
    }
    b.release();
}

static void poly_erase(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_poly_mu);
    poly_map().erase(h);
}
static int find_poly(uint64_t h, Poly** out) {
    std::lock_guard<std::mutex> lk(g_poly_mu);
    auto it = poly_map().find(h);
    if (it == poly_map().end()) { set_error("unknown polynomial handle %llu", (unsigned long long)h); return ZKB_ERR_HANDLE; }
    if (it->second->slot != cur_slot()) {
        set_error("polynomial handle %llu lives on device slot %d; this thread acts on slot %d (zkb_thread_bind_device)", (unsigned long long)h, it->second->slot, cur_slot());
        return ZKB_ERR_HANDLE;
    }
    *out = it->second;
    return ZKB_OK;
}
static int new_poly(uint64_t n, Poly** out, uint64_t* handle) {
    Poly* p = new Poly();
    p->n = n;
    p->slot = cur_slot();
    int rc = pool_alloc(p->buf, n ? n * 32 : 32);
    if (rc != ZKB_OK) { delete p; return rc; }
    std::lock_guard<std::mutex> lk(g_poly_mu);
    *handle = g_next_poly++;
    poly_map()[*handle] = p;
    *out = p;
    return ZKB_OK;
}
void poly_release_all() {
    for (auto& kv : poly_map()) { kv.second->buf.release(); delete kv.second; }
    poly_map().clear();
    poly_pool_flush();
    PolyWs& w = poly_ws();
    w.tmp.release();
    graph_ws().release();
    for (int i = 0; i < 4; ++i) {
        graph_ring().buf[i].release();
        if (graph_ring().ev[i]) { cudaEventDestroy(graph_ring().ev[i]); graph_ring().ev[i] = nullptr; }
    }
    if (w.h_out) cudaFreeHost(w.h_out);
    w.h_out = nullptr;
}

int srs_msm_dev_by_handle(uint64_t srs_handle, const uint4* d_scalars, size_t n, cudaStream_t s, uint64_t* out);  // api.cu
int domain_dev_by_op(int op, const uint4* d_in, uint4* d_a, uint4* d_b, size_t ncols, uint32_t k, uint32_t ek, cudaStream_t s);  // api.cu

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_poly_upload(const uint64_t* values, size_t n, uint64_t* handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!handle || (n && !values)) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(new_poly(n, &p, handle));
    if (n) {
        cudaError_t e = cudaMemcpyAsync(p->buf.p, values, n * 32, cudaMemcpyHostToDevice, ctx().stream);
        count_h2d(n * 32);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);  // the caller may reuse `values` on return
        if (e != cudaSuccess) { set_error("polynomial upload failed: %s", cudaGetErrorString(e)); return ZKB_ERR_CUDA; }
    }
    return ZKB_OK;
}

// n raw Fr (32 B Montgomery limbs each: the element encoding of SerdeFormat::RawBytesUnchecked, which is how the reference
// stores its proving keys, /root/reference/aggregator/src/wrapper.rs:970-988) from `path` at byte `offset` straight into a
// polynomial handle: read(2) into one pinned buffer while the other is in flight to the device.
int zkb_poly_load_file(const char* path, uint64_t offset, size_t n, uint64_t* handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!handle || !path) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return ZKB_ERR_ARG; }
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    if (fseeko(f, 0, SEEK_END) != 0) { set_error("cannot seek in %s", path); return ZKB_ERR_ARG; }
    const uint64_t fsize = (uint64_t)ftello(f);
    if (offset > fsize || (uint64_t)n * 32 > fsize - offset) {
        set_error("%s holds %llu bytes, %zu field elements at offset %llu need %llu", path, (unsigned long long)fsize, n,
                  (unsigned long long)offset, (unsigned long long)(offset + (uint64_t)n * 32));
        return ZKB_ERR_ARG;
    }
    if (fseeko(f, (off_t)offset, SEEK_SET) != 0) { set_error("cannot seek in %s", path); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(new_poly(n, &p, handle));
    const size_t chunk = (size_t)16 << 20;
    void* pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = ctx().stream;
    auto cleanup = [&](int code) {
        cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i) {
            if (pin[i]) cudaFreeHost(pin[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (code != ZKB_OK) { pool_release(p->buf); delete p; poly_erase(*handle); *handle = 0; }
        return code;
    };
    const size_t total = n * 32;
    const size_t pin_bytes = total < chunk ? (total ? total : 32) : chunk;
    for (int i = 0; i < 2; ++i) {
        if (cudaMallocHost(&pin[i], pin_bytes) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            set_error("pinned staging allocation failed");
            return cleanup(ZKB_ERR_OOM);
        }
    }
    for (size_t done = 0, i = 0; done < total; ++i) {
        const size_t len = total - done < chunk ? total - done : chunk;
        if (i >= 2 && cudaEventSynchronize(ev[i & 1]) != cudaSuccess) { set_error("polynomial upload failed"); return cleanup(ZKB_ERR_CUDA); }
        if (fread(pin[i & 1], 1, len, f) != len) { set_error("short read from %s", path); return cleanup(ZKB_ERR_ARG); }
        cudaError_t e = cudaMemcpyAsync(p->buf.as<char>() + done, pin[i & 1], len, cudaMemcpyHostToDevice, st);
        count_h2d(len);
        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
        if (e != cudaSuccess) { set_error("polynomial upload failed: %s", cudaGetErrorString(e)); return cleanup(ZKB_ERR_CUDA); }
        done += len;
    }
    return cleanup(ZKB_OK);
}

int zkb_poly_alloc(size_t n, uint64_t* handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!handle) { set_error("handle is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(new_poly(n, &p, handle));
    ZKB_CUDA_TRY(cudaMemsetAsync(p->buf.p, 0, n ? n * 32 : 32, ctx().stream));
    return ZKB_OK;
}

// poly[offset .. offset + n) = values: the prover overwrites the last blinding_factors + 1 rows of a product column (computed on
// the device) with its random blinding values before committing
int zkb_poly_write(uint64_t handle, size_t offset, const uint64_t* values, size_t n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly* p;
    ZKB_TRY(find_poly(handle, &p));
    if (offset > p->n || n > p->n - offset) { set_error("write [%zu, %zu) into a polynomial of %zu elements", offset, offset + n, (size_t)p->n); return ZKB_ERR_ARG; }
    if (n == 0) return ZKB_OK;
    if (!values) { set_error("values is NULL"); return ZKB_ERR_ARG; }
    ZKB_CUDA_TRY(cudaMemcpyAsync(p->buf.as<char>() + offset * 32, values, n * 32, cudaMemcpyHostToDevice, ctx().stream));
    count_h2d(n * 32);
    ZKB_CUDA_TRY(cudaStreamSynchronize(ctx().stream));
    return ZKB_OK;
}

int zkb_poly_len(uint64_t handle, size_t* n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    Poly* p;
    ZKB_TRY(find_poly(handle, &p));
    if (n) *n = p->n;
    return ZKB_OK;
}

int zkb_poly_download(uint64_t handle, uint64_t* out, size_t n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly* p;
    ZKB_TRY(find_poly(handle, &p));
    if (n > p->n) { set_error("polynomial holds %zu elements, %zu requested", (size_t)p->n, n); return ZKB_ERR_ARG; }
    if (n == 0) return ZKB_OK;
    if (!out) { set_error("out is NULL"); return ZKB_ERR_ARG; }
    ZKB_CUDA_TRY(cudaMemcpyAsync(out, p->buf.p, n * 32, cudaMemcpyDeviceToHost, ctx().stream));
    ZKB_CUDA_TRY(cudaStreamSynchronize(ctx().stream));
    return ZKB_OK;
}

int zkb_poly_free(uint64_t handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    Poly* p;
    ZKB_TRY(find_poly(handle, &p));
    cudaSetDevice(ctx().device);
    if (poly_pool_limit() == 0) cudaStreamSynchronize(ctx().stream);
    pool_release(p->buf);  // stream-ordered reuse; a buffer that does not fit in the pool is cudaFree'd (which synchronises)
    delete p;
    poly_erase(handle);
    return ZKB_OK;
}

// new handle holding poly[offset .. offset + n): the pieces of h(X) after extended_to_coeff (committed and opened one by one)
int zkb_poly_slice(uint64_t poly, size_t offset, size_t n, uint64_t* out_handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!out_handle) { set_error("out_handle is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (offset > p->n || n > p->n - offset) { set_error("slice [%zu, %zu) of a polynomial with %zu elements", offset, offset + n, (size_t)p->n); return ZKB_ERR_ARG; }
    Poly* q;
    ZKB_TRY(new_poly(n, &q, out_handle));
    if (n) {
        cudaError_t e = cudaMemcpyAsync(q->buf.p, p->buf.as<char>() + offset * 32, n * 32, cudaMemcpyDeviceToDevice, ctx().stream);
        if (e != cudaSuccess) {
            set_error("slice copy failed: %s", cudaGetErrorString(e));
            pool_release(q->buf); delete q; poly_erase(*out_handle); *out_handle = 0;
            return ZKB_ERR_CUDA;
        }
    }
    return ZKB_OK;
}

int zkb_poly_commit(uint64_t srs_handle, uint64_t poly, uint64_t out_jac[12]) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!out_jac) { set_error("out_jac is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    return srs_msm_dev_by_handle(srs_handle, p->buf.as<uint4>(), p->n, ctx().stream, out_jac);
}

// op: 1 lagrange_to_coeff, 2 coeff_to_lagrange, 4 extended_to_coeff — in place on the handle
static int poly_domain_inplace(int op, uint64_t poly, uint32_t k, uint32_t ek) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    const uint64_t N = 1ull << (op == 4 ? ek : k);
    if (p->n != N) { set_error("polynomial holds %zu elements, the domain needs %zu", (size_t)p->n, (size_t)N); return ZKB_ERR_ARG; }
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(N * 32));
    return domain_dev_by_op(op, p->buf.as<uint4>(), p->buf.as<uint4>(), w.tmp.as<uint4>(), 1, k, ek, ctx().stream);
}
int zkb_poly_lagrange_to_coeff(uint64_t poly, uint32_t k) { return poly_domain_inplace(1, poly, k, k); }
int zkb_poly_coeff_to_lagrange(uint64_t poly, uint32_t k) { return poly_domain_inplace(2, poly, k, k); }
int zkb_poly_extended_to_coeff(uint64_t poly, uint32_t k, uint32_t extended_k) { return poly_domain_inplace(4, poly, k, extended_k); }

int zkb_poly_coeff_to_extended(uint64_t poly, uint32_t k, uint32_t extended_k, uint64_t* out_handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!out_handle) { set_error("out_handle is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (extended_k < k || extended_k > 28 || p->n != (1ull << k)) { set_error("polynomial length / domain mismatch"); return ZKB_ERR_ARG; }
    const uint64_t N = 1ull << extended_k;
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(N * 32));
    Poly* e;
    ZKB_TRY(new_poly(N, &e, out_handle));
    int rc = domain_dev_by_op(3, p->buf.as<uint4>(), e->buf.as<uint4>(), w.tmp.as<uint4>(), 1, k, extended_k, ctx().stream);
    if (rc != ZKB_OK) { pool_release(e->buf); delete e; poly_erase(*out_handle); *out_handle = 0; }
    return rc;
}

int zkb_poly_eval(uint64_t poly, const uint64_t x[4], uint64_t out[4]) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!x || !out) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n == 0) { memset(out, 0, 32); return ZKB_OK; }
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(poly_tmp_bytes(p->n)));
    if (!w.h_out) ZKB_CUDA_TRY(cudaMallocHost(&w.h_out, 64));
    uint4* d_res = nullptr;
    ZKB_TRY(poly_eval_dev(p->buf.as<uint4>(), p->n, fr_of(x), w.tmp.as<uint4>(), &d_res, ctx().stream));
    ZKB_CUDA_TRY(cudaMemcpyAsync(w.h_out, d_res, 32, cudaMemcpyDeviceToHost, ctx().stream));
    ZKB_CUDA_TRY(cudaStreamSynchronize(ctx().stream));
    memcpy(out, w.h_out, 32);
    return ZKB_OK;
}

int zkb_poly_kate_division(uint64_t poly, const uint64_t b[4], uint64_t* out_handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!b || !out_handle) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n < 1) { set_error("kate_division of an empty polynomial"); return ZKB_ERR_ARG; }
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(poly_tmp_bytes(p->n)));
    Poly* q;
    ZKB_TRY(new_poly(p->n, &q, out_handle));  // Q_0 .. Q_{n-1}; the quotient is the first n-1 of them (Q_{n-1} = 0)
    int rc = kate_q_dev(p->buf.as<uint4>(), p->n, fr_of(b), q->buf.as<uint4>(), w.tmp.as<uint4>(), ctx().stream);
    if (rc != ZKB_OK) { pool_release(q->buf); delete q; poly_erase(*out_handle); *out_handle = 0; return rc; }
    q->n = p->n - 1;  // upstream returns a.len() - 1 coefficients
    return ZKB_OK;
}

int zkb_poly_batch_invert(uint64_t poly) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n == 0) return ZKB_OK;
    BatchInvertArgs a{p->buf.as<uint4>(), p->n};
    fr_batch_invert_kernel<<<nblk((p->n + INV_CHUNK - 1) / INV_CHUNK, 128), 128, 0, ctx().stream>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

int zkb_poly_mul(uint64_t poly, uint64_t other) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly *p, *q;
    ZKB_TRY(find_poly(poly, &p));
    ZKB_TRY(find_poly(other, &q));
    if (p->n != q->n) { set_error("element-wise product of %zu and %zu elements", (size_t)p->n, (size_t)q->n); return ZKB_ERR_ARG; }
    if (p->n == 0) return ZKB_OK;
    PolyMulArgs a{p->buf.as<uint4>(), q->buf.as<uint4>(), p->n};
    poly_mul_kernel<<<nblk(p->n, 128), 128, 0, ctx().stream>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// poly[i] *= table[i mod period]
int zkb_poly_mul_periodic(uint64_t poly, const uint64_t* table, uint32_t period) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!table || period == 0 || period > POLY_PERIOD_MAX || (period & (period - 1))) { set_error("period must be a power of two <= %u", POLY_PERIOD_MAX); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n == 0) return ZKB_OK;
    PolyMulPeriodicArgs a{};
    a.a = p->buf.as<uint4>(); a.n = p->n; a.period = period;
    for (uint32_t i = 0; i < period; ++i) words_of(fr_of(table + 4 * i), a.t[i]);
    poly_mul_periodic_kernel<<<nblk(p->n, 128), 128, 0, ctx().stream>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// poly <- poly * k + other   (other == 0: poly <- poly * k); lengths must agree
int zkb_poly_scale_add(uint64_t poly, const uint64_t k[4], uint64_t other) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!k) { set_error("k is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    Poly* q = nullptr;
    ZKB_TRY(find_poly(poly, &p));
    if (other) {
        ZKB_TRY(find_poly(other, &q));
        if (q->n != p->n) { set_error("scale_add of %zu and %zu elements", (size_t)p->n, (size_t)q->n); return ZKB_ERR_ARG; }
    }
    if (p->n == 0) return ZKB_OK;
    PolyScaleAddArgs a{};
    a.acc = p->buf.as<uint4>(); a.other = q ? q->buf.as<uint4>() : nullptr; a.n = p->n;
    words_of(fr_of(k), a.k);
    poly_scale_add_kernel<<<nblk(p->n, 128), 128, 0, ctx().stream>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// coefficient 0 += c   (f(X) - f(x) before kate_division: pass c = -f(x))
int zkb_poly_add_const(uint64_t poly, const uint64_t c[4]) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!c) { set_error("c is NULL"); return ZKB_ERR_ARG; }
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n == 0) { set_error("empty polynomial"); return ZKB_ERR_ARG; }
    poly_add_const_kernel<<<1, 1, 0, ctx().stream>>>(p->buf.as<uint4>(), fr_of(c));
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

int zkb_poly_prefix_product(uint64_t poly) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Poly* p;
    ZKB_TRY(find_poly(poly, &p));
    if (p->n == 0) return ZKB_OK;
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(poly_tmp_bytes(p->n)));
    return prefix_product_dev(p->buf.as<uint4>(), p->n, p->buf.as<uint4>(), w.tmp.as<uint4>(), ctx().stream);
}

// ---- lookup argument: permute_expression_pair on resident columns (lookup.cuh) -------------------------------------------------------
// sorted order of `canon` (u canonical 256-bit values) into idx: block-wise bitonic sort + merge-path merges (lookup.cuh)
static int lookup_sort_dev(const uint4* canon, uint64_t u, uint32_t* idx, DevBuf& ws, cudaStream_t s) {
    LookupSortArgs a{};
    a.log_p = lookup_sort_log_p(u);
    const uint64_t P = (uint64_t)1 << a.log_p;
    const uint64_t tiles = P / LOOKUP_MERGE_TILE;
    ZKB_TRY(ws.reserve(2 * P * 36 + tiles * 4 + 64));
    // two record buffers (klo | khi | row each) and the merge partition
    char* b = ws.as<char>();
    uint4* lo[2]; uint4* hi[2]; uint32_t* row[2];
    for (int i = 0; i < 2; ++i) { lo[i] = reinterpret_cast<uint4*>(b); b += P * 16; hi[i] = reinterpret_cast<uint4*>(b); b += P * 16; }
    for (int i = 0; i < 2; ++i) { row[i] = reinterpret_cast<uint32_t*>(b); b += P * 4; }
    uint32_t* part = reinterpret_cast<uint32_t*>(b);
    a.klo = lo[0]; a.khi = hi[0]; a.row = row[0];
    a.canon = canon;
    a.u = u;
    lookup_sort_init_kernel<<<nblk(P, 128), 128, 0, s>>>(a);
    constexpr size_t block_smem = (size_t)LOOKUP_SORT_BLOCK * 36;
    struct LookupSortAttr { bool set = false; };
    bool& attr = per_device<LookupSortAttr>().set;
    if (!attr) {
        ZKB_CUDA_TRY(cudaFuncSetAttribute(lookup_sort_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_smem));
        attr = true;
    }
    lookup_sort_block_kernel<<<(unsigned)(P / LOOKUP_SORT_BLOCK), LOOKUP_SORT_THREADS, block_smem, s>>>(a);
    count_launch(2);
    int cur = 0;
    for (uint64_t run = LOOKUP_SORT_BLOCK; run < P; run <<= 1, cur ^= 1) {
        LookupMergeArgs m{};
        m.ilo = lo[cur]; m.ihi = hi[cur]; m.irow = row[cur];
        m.olo = lo[cur ^ 1]; m.ohi = hi[cur ^ 1]; m.orow = row[cur ^ 1];
        m.run = run; m.log_p = a.log_p; m.part = part;
        lookup_merge_partition_kernel<<<nblk(tiles, 128), 128, 0, s>>>(m);
        lookup_merge_tile_kernel<<<(unsigned)tiles, LOOKUP_MERGE_THREADS, 0, s>>>(m);
        count_launch(2);
    }
    lookup_take_rows_kernel<<<nblk(u, 128), 128, 0, s>>>(row[cur], idx, u);   // the padding sorted to the end
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// ---- quotient evaluation: GraphEvaluator::evaluate for every row (graph.cuh, graph_plan.hpp) --------------------------------------
struct GraphInfo { uint32_t instructions = 0, slots = 0, polys = 0, bytes_per_row = 0; };
static GraphInfo g_graph_info;
constexpr size_t GRAPH_STAGE_PROG_BYTES = 32 << 10;

// lowered program + operand tables -> one upload into `ws` and the launch on `s`
static int graph_launch(const GraphPlan& plan, const std::vector<const uint4*>& ptrs, uint4* d_values, uint64_t nrows, uint64_t mask,
                        cudaStream_t s, DevBuf& ws) {
    const uint32_t ninstr = (uint32_t)plan.prog.size();
    const bool stage = (size_t)ninstr * 16 <= GRAPH_STAGE_PROG_BYTES;
    int smem_limit = 0;
    ZKB_CUDA_TRY(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx().device));
    const uint32_t threads = graph_cta_threads(plan.nslots, stage ? ninstr : 0, (size_t)smem_limit);
    if (threads == 0) { set_error("graph: %u live intermediates do not fit in shared memory", plan.nslots); return ZKB_ERR_ARG; }
    const size_t smem = (size_t)plan.nslots * 32 * threads + (stage ? (size_t)ninstr * 16 : 0);

    // one upload: program | scalars | column queries
    std::vector<GraphQuery> queries;
    for (auto& q : plan.queries) queries.push_back(GraphQuery{ptrs[q.first], q.second});
    auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t o_prog = 0, o_sc = up16(o_prog + (size_t)ninstr * 16), o_q = up16(o_sc + plan.scalars.size() * 8),
                 total = up16(o_q + queries.size() * sizeof(GraphQuery)) + 16;
    std::vector<unsigned char> host(total, 0);
    if (ninstr) memcpy(host.data() + o_prog, plan.prog.data(), (size_t)ninstr * 16);
    if (!plan.scalars.empty()) memcpy(host.data() + o_sc, plan.scalars.data(), plan.scalars.size() * 8);
    if (!queries.empty()) memcpy(host.data() + o_q, queries.data(), queries.size() * sizeof(GraphQuery));
    ZKB_TRY(ws.reserve(total));
    ZKB_CUDA_TRY(cudaMemcpyAsync(ws.p, host.data(), total, cudaMemcpyHostToDevice, s));  // pageable source: staged before return
    count_h2d(total);
    char* d = reinterpret_cast<char*>(ws.p);
    GraphArgs a{};
    a.prog = reinterpret_cast<const uint4*>(d + o_prog);
    a.ninstr = ninstr;
    a.result_slot = plan.result_slot;
    a.scalars = reinterpret_cast<const uint4*>(d + o_sc);
    a.queries = reinterpret_cast<const GraphQuery*>(d + o_q);
    a.values = d_values;
    a.nrows = nrows;
    a.mask = mask;
    if (smem > 48 * 1024) ZKB_CUDA_TRY(cudaFuncSetAttribute(graph_evaluate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope prof("graph_evaluate", s);
        graph_evaluate_kernel<<<nblk(nrows, threads), threads, smem, s>>>(a, plan.nslots, stage ? 1u : 0u);
    }
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    g_graph_info.instructions = ninstr;
    g_graph_info.slots = plan.nslots;
    g_graph_info.polys = (uint32_t)ptrs.size();
    g_graph_info.bytes_per_row = 32 * ((uint32_t)ptrs.size() + (plan.uses_prev ? 1 : 0) + 1);
    return ZKB_OK;
}

int zkb_graph_evaluate(const zkb_graph* graph, const zkb_graph_inputs* inputs, uint64_t values) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!graph || !inputs) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    Poly* out;
    ZKB_TRY(find_poly(values, &out));
    const uint64_t isize = out->n;
    GraphPlan plan;
    const std::string err = graph_lower(*graph, *inputs, isize, plan);
    if (!err.empty()) { set_error("graph: %s", err.c_str()); return ZKB_ERR_ARG; }
    std::vector<const uint4*> ptrs;
    for (uint64_t h : plan.poly_handles) {
        Poly* p;
        ZKB_TRY(find_poly(h, &p));
        if (p->n != isize) { set_error("graph: polynomial %llu holds %zu elements, values holds %zu", (unsigned long long)h, (size_t)p->n, (size_t)isize); return ZKB_ERR_ARG; }
        if (p == out) { set_error("graph: values is also read as a column (rows are not evaluated in order)"); return ZKB_ERR_ARG; }
        ptrs.push_back(p->buf.as<uint4>());
    }
    return graph_launch(plan, ptrs, out->buf.as<uint4>(), isize, isize - 1, ctx().stream, graph_ws());  // stream-ordered reuse of the upload buffer
}

// Device pointers, the caller's stream, and optionally a row WINDOW (one rank's share of the extended domain): the `fixed` /
// `advice` / `instance` arrays of `inputs` hold device addresses instead of handles.
int zkb_graph_evaluate_dev(const zkb_graph* graph, const zkb_graph_inputs* inputs, void* d_values, size_t rows, int window,
                           size_t halo_lo, size_t halo_hi, void* stream) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!graph || !inputs || !d_values) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    GraphWindow win;
    win.on = window != 0;
    win.halo_lo = halo_lo;
    win.halo_hi = halo_hi;
    if (!win.on && (halo_lo || halo_hi)) { set_error("graph: a halo needs window mode"); return ZKB_ERR_ARG; }
    GraphPlan plan;
    const std::string err = graph_lower(*graph, *inputs, rows, plan, win);
    if (!err.empty()) { set_error("graph: %s", err.c_str()); return ZKB_ERR_ARG; }
    std::vector<const uint4*> ptrs;
    for (uint64_t h : plan.poly_handles) {
        if (!h || (h & 15)) { set_error("graph: column pointers must be non-NULL and 16-byte aligned"); return ZKB_ERR_ARG; }
        if (h == (uint64_t)(uintptr_t)d_values) { set_error("graph: values is also read as a column (rows are not evaluated in order)"); return ZKB_ERR_ARG; }
        ptrs.push_back(reinterpret_cast<const uint4*>((uintptr_t)h));
    }
    // the upload buffers are shared between calls on arbitrary streams: a ring, each slot reused only after the kernel that read it
    GraphRing& r = graph_ring();
    const unsigned k = r.next++ & 3;
    if (!r.ev[k]) ZKB_CUDA_TRY(cudaEventCreateWithFlags(&r.ev[k], cudaEventDisableTiming));
    else ZKB_CUDA_TRY(cudaEventSynchronize(r.ev[k]));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ZKB_TRY(graph_launch(plan, ptrs, reinterpret_cast<uint4*>(d_values), rows, win.on ? ~0ull : (uint64_t)rows - 1, s, r.buf[k]));
    ZKB_CUDA_TRY(cudaEventRecord(r.ev[k], s));
    return ZKB_OK;
}

int zkb_graph_last_info(uint32_t* instructions, uint32_t* slots, uint32_t* polys_read, uint32_t* bytes_per_row) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    if (instructions) *instructions = g_graph_info.instructions;
    if (slots) *slots = g_graph_info.slots;
    if (polys_read) *polys_read = g_graph_info.polys;
    if (bytes_per_row) *bytes_per_row = g_graph_info.bytes_per_row;
    return ZKB_OK;
}

int zkb_field_vec_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (field < 0 || field > 1 || op < 0 || op > 10) { set_error("bad field / op"); return ZKB_ERR_ARG; }
    if (n == 0) return ZKB_OK;
    if (!a || !b || !out) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    PolyWs& w = poly_ws();
    ZKB_TRY(w.tmp.reserve(3 * n * 32));
    char* d = reinterpret_cast<char*>(w.tmp.p);
    cudaStream_t s = ctx().stream;
    ZKB_CUDA_TRY(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, s));
    count_h2d(n * 32);
    ZKB_CUDA_TRY(cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, s));
    count_h2d(n * 32);
    field_vec_op_kernel<<<nblk(n, 128), 128, 0, s>>>(field, op, reinterpret_cast<const uint4*>(d), reinterpret_cast<const uint4*>(d + n * 32),
                                                     reinterpret_cast<uint4*>(d + 2 * n * 32), n);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    ZKB_CUDA_TRY(cudaMemcpyAsync(out, d + 2 * n * 32, n * 32, cudaMemcpyDeviceToHost, s));
    ZKB_CUDA_TRY(cudaStreamSynchronize(s));
    return ZKB_OK;
}

// host-buffer conveniences (upload + op + download), for callers that do not keep polynomials resident
int zkb_fr_eval_polynomial(const uint64_t* coeffs, size_t n, const uint64_t x[4], uint64_t out[4]) {
    uint64_t h = 0;
    ZKB_TRY(zkb_poly_upload(coeffs, n, &h));
    int rc = zkb_poly_eval(h, x, out);
    zkb_poly_free(h);
    return rc;
}
int zkb_fr_kate_division(const uint64_t* coeffs, size_t n, const uint64_t b[4], uint64_t* out) {
    if (n < 1) { set_error("kate_division of an empty polynomial"); return ZKB_ERR_ARG; }
    uint64_t h = 0, q = 0;
    ZKB_TRY(zkb_poly_upload(coeffs, n, &h));
    int rc = zkb_poly_kate_division(h, b, &q);
    if (rc == ZKB_OK) rc = zkb_poly_download(q, out, n - 1);
    zkb_poly_free(h);
    if (q) zkb_poly_free(q);
    return rc;
}
int zkb_fr_batch_invert(uint64_t* values, size_t n) {
    uint64_t h = 0;
    ZKB_TRY(zkb_poly_upload(values, n, &h));
    int rc = zkb_poly_batch_invert(h);
    if (rc == ZKB_OK) rc = zkb_poly_download(h, values, n);
    zkb_poly_free(h);
    return rc;
}


// plonk::lookup::prover::permute_expression_pair on resident columns: input / table hold the compressed expressions (>= usable_rows
// elements); two new polynomials of the same length receive A' and S' in rows [0, usable_rows), zeros above (the caller writes
// its blinding rows).  ZKB_ERR_ARG when an input value does not occur in the table (upstream: Error::ConstraintSystemFailure).
int zkb_lookup_permute_expression_pair(uint64_t input, uint64_t table, size_t usable_rows, uint64_t* permuted_input, uint64_t* permuted_table) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!permuted_input || !permuted_table) { set_error("NULL output handle"); return ZKB_ERR_ARG; }
    Poly *pi, *pt;
    ZKB_TRY(find_poly(input, &pi));
    ZKB_TRY(find_poly(table, &pt));
    const uint64_t u = usable_rows;
    if (pi->n != pt->n || u > pi->n || u >= (1ull << 31)) { set_error("lookup columns hold %zu / %zu elements, usable rows %zu", (size_t)pi->n, (size_t)pt->n, usable_rows); return ZKB_ERR_ARG; }
    Poly *oi, *ot;
    ZKB_TRY(new_poly(pi->n, &oi, permuted_input));
    int rc = new_poly(pi->n, &ot, permuted_table);
    if (rc != ZKB_OK) { zkb_poly_free(*permuted_input); *permuted_input = 0; return rc; }
    cudaStream_t s = ctx().stream;
    auto fail = [&](int code) {
        cudaStreamSynchronize(s);
        zkb_poly_free(*permuted_input); zkb_poly_free(*permuted_table);
        *permuted_input = *permuted_table = 0;
        return code;
    };
    if (cudaMemsetAsync(oi->buf.p, 0, pi->n ? pi->n * 32 : 32, s) != cudaSuccess || cudaMemsetAsync(ot->buf.p, 0, pi->n ? pi->n * 32 : 32, s) != cudaSuccess) return fail(ZKB_ERR_CUDA);
    if (u == 0) return ZKB_OK;
    // workspace: canonical copies (2 x 32 u) and the index / flag arrays in one buffer; the sort's records and the scan's tile sums in `sort`
    struct LookupWs { DevBuf buf, sort; uint32_t* h_flag = nullptr; };
    LookupWs& ws = per_device<LookupWs>();
    const size_t need = u * (64 + 4 * 10) + 256;
    rc = ws.buf.reserve(need);
    if (rc != ZKB_OK) return fail(rc);
    if (!ws.h_flag && cudaMallocHost(&ws.h_flag, 16) != cudaSuccess) { cudaGetLastError(); set_error("pinned allocation failed"); return fail(ZKB_ERR_OOM); }
    char* b = ws.buf.as<char>();
    uint4* canon_in = reinterpret_cast<uint4*>(b); b += u * 32;
    uint4* canon_tab = reinterpret_cast<uint4*>(b); b += u * 32;
    uint32_t* arr[10];
    for (auto& x : arr) { x = reinterpret_cast<uint32_t*>(b); b += u * 4; }
    uint32_t* d_flags = reinterpret_cast<uint32_t*>(b);   // [0] missing
    uint32_t *idx_in = arr[0], *idx_tab = arr[1], *first = arr[3], *consumed = arr[4], *rep_rank = arr[5], *left_rank = arr[6],
             *rep_rows = arr[7], *inv = arr[8];
    LookupArgs a{};
    a.input = pi->buf.as<uint4>(); a.table = pt->buf.as<uint4>(); a.u = u;
    a.canon_in = canon_in; a.canon_tab = canon_tab; a.idx_in = idx_in; a.idx_tab = idx_tab; a.first = first; a.consumed = consumed;
    a.rep_rank = rep_rank; a.left_rank = left_rank; a.out_in = oi->buf.as<uint4>(); a.out_tab = ot->buf.as<uint4>(); a.rep_rows = rep_rows;
    a.missing = d_flags;
    ProfScope prof("lookup_permute", s);
    cudaMemsetAsync(d_flags, 0, 16, s);
    cudaMemsetAsync(consumed, 0, u * 4, s);
    const unsigned nb = nblk(u, 128);
    lookup_canon_kernel<<<nb, 128, 0, s>>>(a);
    count_launch();
    rc = lookup_sort_dev(canon_in, u, idx_in, ws.sort, s);
    if (rc == ZKB_OK) rc = lookup_sort_dev(canon_tab, u, idx_tab, ws.sort, s);
    if (rc != ZKB_OK) return fail(rc);
    lookup_first_kernel<<<nb, 128, 0, s>>>(a);
    lookup_match_kernel<<<nb, 128, 0, s>>>(a);
    count_launch(2);
    // ranks of the repeated rows and of the leftover table elements (exclusive scans of the complemented flags)
    auto scan_not = [&](const uint32_t* flags, uint32_t* out) -> int {
        lookup_not_kernel<<<nb, 128, 0, s>>>(flags, inv, u);
        count_launch();
        ZKB_TRY(scan_u32_exclusive_dev(inv, out, u, ws.sort, s));
        return ZKB_OK;
    };
    rc = scan_not(first, rep_rank);
    if (rc != ZKB_OK) return fail(rc);
    // number of repeated rows = rank of the last row + its own flag
    uint32_t tail[2] = {0, 0};
    if (cudaMemcpyAsync(&ws.h_flag[0], rep_rank + (u - 1), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&ws.h_flag[1], first + (u - 1), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        cudaGetLastError(); set_error("lookup: device read-back failed"); return fail(ZKB_ERR_CUDA);
    }
    tail[0] = ws.h_flag[0]; tail[1] = ws.h_flag[1];
    a.repeated = tail[0] + (1u - tail[1]);
    rc = scan_not(consumed, left_rank);
    if (rc != ZKB_OK) return fail(rc);
    lookup_rep_rows_kernel<<<nb, 128, 0, s>>>(a);
    lookup_leftover_kernel<<<nb, 128, 0, s>>>(a);
    count_launch(2);
    if (cudaMemcpyAsync(&ws.h_flag[2], d_flags, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        set_error("lookup: kernels failed"); return fail(ZKB_ERR_CUDA);
    }
    if (ws.h_flag[2]) { set_error("lookup: an input value does not occur in the table (ConstraintSystemFailure)"); return fail(ZKB_ERR_ARG); }
    return ZKB_OK;
}

}  // extern "C"
