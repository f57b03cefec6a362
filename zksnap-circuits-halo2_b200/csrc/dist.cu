// dist.cu — one NTT sharded over the GPUs of a box: peer HBM over NVLink 5 / NVSwitch (CUDA IPC), exchange fused into the
// NTT passes, device-side barriers.  SURVEY.md §8(e) row "single NTT larger than one GPU's share".
//
// One process per GPU (CUDA IPC mappings exchanged by the caller), or one process driving all GPUs (zkb_init with several
// devices: plain peer access, ranks = device slots, every rank's kernels launched by its own worker thread).  Rank r owns the contiguous slice [r N/G, (r+1) N/G) of the natural-order input and receives the
// same slice of the natural-order output.  The multi-pass NTT of ntt.cuh is kept as is; only the addressing changes:
//
//   pass 0        every CTA gathers its R_1 x T tile from all ranks' input slices (peer loads: the "transpose" of the
//                 four-step algorithm happens inside the load phase) and writes k_1 where n_1 was, i.e. into the work
//                 slice of the rank that owns that address (peer stores);
//   barrier       device-side: one release/acquire flag per peer in IPC-mapped memory, no host round trip;
//   middle passes touch the rank's own work slice only;
//   final pass    reads local, stores each output to the slice of the rank that owns its natural-order position
//                 (peer stores fused into the store phase);
//   barrier.
//
// There is no separate transpose pass and no staging copy: the collective is the kernels' own loads and stores.
#include <cstdlib>
#include <cstring>

#include "ntt_host.hpp"

namespace zkb {

struct DistCtx {
    bool created = false, connected = false;
    bool inprocess = false;        // peers are other device slots of this process (peer access, no IPC handles)
    int rank = -1, world = 0;
    uint32_t log_g = 0, max_log_n = 0;
    size_t slice_bytes = 0;
    void* local[3] = {nullptr, nullptr, nullptr};  // A (input slice), W (work slice), O (output slice)
    uint32_t* flags = nullptr;                     // [NTT_MAX_RANKS] arrival epochs written by the peers, + status word
    void* peer[3][NTT_MAX_RANKS] = {};
    uint32_t* peer_flags[NTT_MAX_RANKS] = {};
    uint32_t epoch = 0;
    uint32_t* h_status = nullptr;  // pinned mirror of the barrier status word
};
static DistCtx& dctx() { return per_device<DistCtx>(); }

struct DistBarrierArgs {
    uint32_t* peer_flags[NTT_MAX_RANKS];
    uint32_t* my_flags;
    uint32_t rank, world, epoch;
    unsigned long long timeout_ns;
};

// Thread j signals peer j (flags_j[rank] = epoch, release at system scope: everything this GPU wrote before — including
// the previous kernels' peer stores — is visible to whoever acquires the flag) and waits for peer j's signal.
__global__ void dist_barrier_kernel(const DistBarrierArgs a) {
    const uint32_t j = threadIdx.x;
    if (j >= a.world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[j] + a.rank), "r"(a.epoch) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.my_flags + j) : "memory");
        if ((int32_t)(v - a.epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > a.timeout_ns) {  // a peer never arrived: report instead of hanging the GPU; the NTT passes that follow
            a.my_flags[NTT_MAX_RANKS] = 1;  // see the status word and return at once, so no half-exchanged data is touched
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

static unsigned long long g_dist_timeout_ms = 0;   // 0: ZKB_DIST_TIMEOUT_MS or 120 s
static unsigned long long dist_timeout_ms() {
    if (g_dist_timeout_ms) return g_dist_timeout_ms;
    static const unsigned long long env = getenv("ZKB_DIST_TIMEOUT_MS") ? strtoull(getenv("ZKB_DIST_TIMEOUT_MS"), nullptr, 10) : 120000ull;
    return env ? env : 120000ull;
}

static int dist_barrier(cudaStream_t s) {
    DistCtx& d = dctx();
    DistBarrierArgs a{};
    for (int r = 0; r < d.world; ++r) a.peer_flags[r] = d.peer_flags[r];
    a.my_flags = d.flags;
    a.rank = (uint32_t)d.rank;
    a.world = (uint32_t)d.world;
    a.epoch = ++d.epoch;
    // Ranks are independent processes (or threads): a first-call plan build or a large pageable staging copy on one of them can
    // delay its arrival by seconds.  Default 120 s, ZKB_DIST_TIMEOUT_MS overrides.
    a.timeout_ns = dist_timeout_ms() * 1000ull * 1000ull;
    ProfScope prof("dist_barrier", s);
    dist_barrier_kernel<<<1, 32, 0, s>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

static void dist_release() {
    DistCtx& d = dctx();
    if (!d.created) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < d.world; ++r) {
        if (r == d.rank || !d.connected || d.inprocess) continue;
        for (int b = 0; b < 3; ++b)
            if (d.peer[b][r]) cudaIpcCloseMemHandle(d.peer[b][r]);
        if (d.peer_flags[r]) cudaIpcCloseMemHandle(d.peer_flags[r]);
    }
    for (int b = 0; b < 3; ++b)
        if (d.local[b]) cudaFree(d.local[b]);
    if (d.flags) cudaFree(d.flags);
    if (d.h_status) cudaFreeHost(d.h_status);
    cudaGetLastError();
    d = DistCtx();
}

struct InprocShared;
static InprocShared& inproc();
static void inproc_reset();
void dist_shutdown() {
    dist_release();
    if (cur_slot() == 0) inproc_reset();
}

// Sharded NTT on device slices.  d_in == NULL: the input is already in the symmetric input slice (zkb_dist_buffers);
// d_out == NULL: leave the result in the symmetric output slice.
static int dist_ntt_dev(const void* d_in, void* d_out, const uint64_t omega[4], uint32_t log_n, cudaStream_t s,
                        const Fr* out_scale = nullptr) {
    DistCtx& d = dctx();
    if (!d.connected) { set_error("zkb_dist_connect has not been called"); return ZKB_ERR_ARG; }
    if (log_n > d.max_log_n) { set_error("log_n %u exceeds the dist context's max_log_n %u", log_n, d.max_log_n); return ZKB_ERR_ARG; }
    NttPlan* plan = nullptr;
    ZKB_TRY(ntt_get_plan(log_n, omega, s, &plan, true));  // small radix first: wide tiles for the exchange pass
    const NttGeometry& g = plan->geom;
    if (d.log_g == 0 || !ntt_dist_supported(g, d.log_g)) {
        set_error("a 2^%u NTT cannot be sharded over %d ranks (needs >= 2 passes and tiles inside a rank's share)", log_n, d.world);
        return ZKB_ERR_ARG;
    }
    const uint64_t N = 1ull << log_n;
    const size_t slice = (size_t)(N >> d.log_g) * 32;
    if (d_in) ZKB_CUDA_TRY(cudaMemcpyAsync(d.local[0], d_in, slice, cudaMemcpyDeviceToDevice, s));
    ZKB_TRY(dist_barrier(s));  // every rank's input slice is in place
    for (uint32_t p = 0; p < g.npass; ++p) {
        NttPassArgs a{};
        const bool fin = p + 1 == g.npass;
        a.log_n = g.log_n; a.npass = g.npass; a.pass = p;
        for (uint32_t q = 0; q < g.npass; ++q) a.lr[q] = g.lr[q];
        a.log_t = g.log_t[p];
        a.is_final = fin ? 1 : 0;
        a.tw_r = plan->tw_r[p].as<uint4>(); a.tw_hi = plan->tw_hi.as<uint4>(); a.tw_lo = plan->tw_lo.as<uint4>();
        a.tw_h = g.tw_h;
        a.tw_pass = (plan->has_tw_pass && p > 0) ? plan->tw_pass[p].as<uint4>() : nullptr;
        a.in_len = N;
        a.out_scale_on = (fin && out_scale) ? 1 : 0;
        if (a.out_scale_on)
            for (int m = 0; m < 3; ++m)
                for (int i = 0; i < 8; ++i) a.out_scale[m][i] = out_scale[m].l[i];
        a.dist_abort = d.flags + NTT_MAX_RANKS;
        a.dist_log_g = d.log_g; a.dist_rank = (uint32_t)d.rank; a.dist_log_slice = log_n - d.log_g;
        for (int r = 0; r < d.world; ++r) {
            a.peer_src[r] = reinterpret_cast<const uint4*>(d.peer[p == 0 ? 0 : 1][r]);
            a.peer_dst[r] = reinterpret_cast<uint4*>(d.peer[fin ? 2 : 1][r]);
        }
        dim3 grid((unsigned)(ntt_cta_count(g, p) >> d.log_g), 1);
        {
            ProfScope prof(p == 0 ? "dist_ntt_pass0" : (fin ? "dist_ntt_final" : "dist_ntt_middle"), s);
            ZKB_TRY(ntt_launch_pass(g.lr[p], a, grid, ntt_cta_threads(g, p), ntt_cta_smem_bytes(g, p), s));
        }
        if (p == 0 || fin) ZKB_TRY(dist_barrier(s));  // all peer stores of this pass have landed
    }
    if (d_out) ZKB_CUDA_TRY(cudaMemcpyAsync(d_out, d.local[2], slice, cudaMemcpyDeviceToDevice, s));
    return ZKB_OK;
}

static int dist_check_status(cudaStream_t s) {
    DistCtx& d = dctx();
    ZKB_CUDA_TRY(cudaMemcpyAsync(d.h_status, d.flags + NTT_MAX_RANKS, 4, cudaMemcpyDeviceToHost, s));
    ZKB_CUDA_TRY(cudaStreamSynchronize(s));
    if (*d.h_status) {
        // reported once: clear the word so that a later collective call (after the late rank has recovered) can succeed
        cudaMemsetAsync(d.flags + NTT_MAX_RANKS, 0, 4, s);
        cudaStreamSynchronize(s);
        set_error("distributed barrier timed out: a peer rank never arrived; the transform was abandoned (its output slice is undefined)");
        return ZKB_ERR_CUDA;
    }
    return ZKB_OK;
}

// ---- one process, several device slots ----------------------------------------------------------------------------------------
struct InprocShared {
    std::mutex mu;
    uint32_t max_log_n = 0;
    int world = 0;
    void* local[ZKB_MAX_DEVICES][3] = {};
    uint32_t* flags[ZKB_MAX_DEVICES] = {};
};
static InprocShared& inproc() {
    static InprocShared s;
    return s;
}
static void inproc_reset() { inproc().world = 0; inproc().max_log_n = 0; }

// Called on slot 0.  (Re)creates the symmetric slices on every slot when the requested size grows.
static int dist_inprocess_setup(uint32_t log_n, int world) {
    InprocShared& sh = inproc();
    if (sh.world == world && sh.max_log_n >= log_n && dctx().connected && dctx().inprocess) return ZKB_OK;
    uint32_t log_g = 0;
    while ((1 << log_g) < world) ++log_g;
    const uint32_t max_log_n = log_n > sh.max_log_n ? log_n : sh.max_log_n;
    const size_t slice_bytes = ((size_t)1 << (max_log_n - log_g)) * 32;
    ZKB_TRY(run_on_devices(world, [&](int slot) -> int {
        std::lock_guard<std::recursive_mutex> lk(ctx().mu);
        ZKB_TRY(require_init());
        dist_release();
        DistCtx& d = dctx();
        d.rank = slot; d.world = world; d.log_g = log_g; d.max_log_n = max_log_n; d.slice_bytes = slice_bytes;
        d.created = true; d.inprocess = true;
        for (int b = 0; b < 3; ++b) {
            if (cudaMalloc(&d.local[b], slice_bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) for the dist slice failed", slice_bytes); dist_release(); return ZKB_ERR_OOM; }
            sh.local[slot][b] = d.local[b];
        }
        ZKB_CUDA_TRY(cudaMalloc(&d.flags, 256));
        ZKB_CUDA_TRY(cudaMemset(d.flags, 0, 256));
        ZKB_CUDA_TRY(cudaMallocHost(&d.h_status, 64));
        ZKB_CUDA_TRY(cudaDeviceSynchronize());
        sh.flags[slot] = d.flags;
        return ZKB_OK;
    }));
    ZKB_TRY(run_on_devices(world, [&](int slot) -> int {
        DistCtx& d = dctx();
        for (int r = 0; r < world; ++r) {
            for (int b = 0; b < 3; ++b) d.peer[b][r] = sh.local[r][b];   // unified addressing + peer access (zkb_init)
            d.peer_flags[r] = sh.flags[r];
        }
        d.connected = true;
        return ZKB_OK;
    }));
    sh.world = world;
    sh.max_log_n = max_log_n;
    return ZKB_OK;
}

// best_fft (optionally with the fused output scaling of lagrange_to_coeff / extended_to_coeff) of ONE host vector over all
// device slots: slot r uploads the contiguous slice [r N/G, (r+1) N/G) over its own PCIe link, the slots run the sharded NTT
// together, every slot downloads its slice of the natural-order result.  Called on slot 0 with its mutex held.
bool dist_inprocess_supported(uint32_t log_n) {
    int world = 1;
    while (world * 2 <= device_slots() && world * 2 <= NTT_MAX_RANKS) world *= 2;
    uint32_t log_g = 0;
    while ((1 << log_g) < world) ++log_g;
    return world >= 2 && log_n <= 28 && ntt_dist_supported(ntt_geometry(log_n, true), log_g);
}

int dist_inprocess_ntt_host(const uint64_t* in, uint64_t* out, const uint64_t omega[4], uint32_t log_n, const Fr* out_scale) {
    int world = 1;
    while (world * 2 <= device_slots() && world * 2 <= NTT_MAX_RANKS) world *= 2;
    uint32_t log_g = 0;
    while ((1 << log_g) < world) ++log_g;
    if (world < 2 || !ntt_dist_supported(ntt_geometry(log_n, true), log_g)) { set_error("a 2^%u NTT cannot be sharded over %d devices", log_n, world); return ZKB_ERR_ARG; }
    ZKB_TRY(dist_inprocess_setup(log_n, world));
    const size_t slice = ((size_t)1 << (log_n - log_g)) * 32;
    const int rc = run_on_devices(world, [&](int slot) -> int {
        std::lock_guard<std::recursive_mutex> lk(ctx().mu);
        ZKB_TRY(require_init());
        DistCtx& d = dctx();
        cudaStream_t s = ctx().stream;
        ZKB_CUDA_TRY(cudaMemcpyAsync(d.local[0], reinterpret_cast<const char*>(in) + (size_t)slot * slice, slice, cudaMemcpyHostToDevice, s));
        count_h2d(slice);
        ZKB_TRY(dist_ntt_dev(nullptr, nullptr, omega, log_n, s, out_scale));
        ZKB_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(out) + (size_t)slot * slice, d.local[2], slice, cudaMemcpyDeviceToHost, s));
        return dist_check_status(s);
    });
    if (rc != ZKB_OK) inproc_reset();   // the ranks' barrier epochs may be out of step: rebuild the contexts next time
    return rc;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_dist_create(int rank, int world, uint32_t max_log_n, uint8_t* handle_out) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!handle_out) { set_error("handle_out is NULL"); return ZKB_ERR_ARG; }
    if (world < 1 || world > NTT_MAX_RANKS || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_error("world must be 1, 2, 4 or 8 and 0 <= rank < world (got rank %d of %d)", rank, world);
        return ZKB_ERR_ARG;
    }
    if (max_log_n < 1 || max_log_n > 28) { set_error("max_log_n %u out of range", max_log_n); return ZKB_ERR_ARG; }
    dist_release();
    DistCtx& d = dctx();
    d.rank = rank;
    d.world = world;
    while ((1 << d.log_g) < world) ++d.log_g;
    d.max_log_n = max_log_n;
    d.slice_bytes = ((size_t)1 << (max_log_n - d.log_g)) * 32;
    d.created = true;
    cudaIpcMemHandle_t h[4];
    for (int b = 0; b < 3; ++b) {
        cudaError_t e = cudaMalloc(&d.local[b], d.slice_bytes);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) for the dist slice failed", d.slice_bytes); dist_release(); return ZKB_ERR_OOM; }
        d.peer[b][rank] = d.local[b];
    }
    ZKB_CUDA_TRY(cudaMalloc(&d.flags, 256));
    ZKB_CUDA_TRY(cudaMemset(d.flags, 0, 256));
    ZKB_CUDA_TRY(cudaMallocHost(&d.h_status, 64));
    d.peer_flags[rank] = d.flags;
    for (int b = 0; b < 3; ++b) ZKB_CUDA_TRY(cudaIpcGetMemHandle(&h[b], d.local[b]));
    ZKB_CUDA_TRY(cudaIpcGetMemHandle(&h[3], d.flags));
    ZKB_CUDA_TRY(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    static_assert(ZKB_DIST_HANDLE_BYTES >= 4 * 64, "handle blob too small");
    memset(handle_out, 0, ZKB_DIST_HANDLE_BYTES);
    memcpy(handle_out, h, sizeof(h));
    return ZKB_OK;
}

int zkb_dist_connect(const uint8_t* all_handles) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    DistCtx& d = dctx();
    if (!d.created) { set_error("zkb_dist_create has not been called"); return ZKB_ERR_ARG; }
    if (!all_handles) { set_error("all_handles is NULL"); return ZKB_ERR_ARG; }
    if (d.connected) return ZKB_OK;
    for (int r = 0; r < d.world; ++r) {
        if (r == d.rank) continue;
        cudaIpcMemHandle_t h[4];
        memcpy(h, all_handles + (size_t)r * ZKB_DIST_HANDLE_BYTES, sizeof(h));
        for (int b = 0; b < 3; ++b) {
            cudaError_t e = cudaIpcOpenMemHandle(&d.peer[b][r], h[b], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("cudaIpcOpenMemHandle(rank %d buffer %d) failed: %s — peer access over NVLink is required", r, b, cudaGetErrorString(e));
                return ZKB_ERR_CUDA;
            }
        }
        void* f = nullptr;
        ZKB_CUDA_TRY(cudaIpcOpenMemHandle(&f, h[3], cudaIpcMemLazyEnablePeerAccess));
        d.peer_flags[r] = reinterpret_cast<uint32_t*>(f);
    }
    d.connected = true;
    return ZKB_OK;
}

// one process, several devices: the symmetric slices of every bound device, connected through peer access (no handles to exchange)
int zkb_dist_create_inprocess(uint32_t max_log_n) {
    if (cur_slot() != 0) { set_error("zkb_dist_create_inprocess must be called from a thread bound to the home device"); return ZKB_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    int world = 1;
    while (world * 2 <= device_slots() && world * 2 <= NTT_MAX_RANKS) world *= 2;
    if (world < 2) { set_error("needs >= 2 bound devices"); return ZKB_ERR_ARG; }
    if (max_log_n < 11 || max_log_n > 28) { set_error("max_log_n %u out of range [11, 28]", max_log_n); return ZKB_ERR_ARG; }
    return dist_inprocess_setup(max_log_n, world);
}

int zkb_dist_set_timeout_ms(uint64_t ms) {
    g_dist_timeout_ms = ms;
    return ZKB_OK;
}

int zkb_dist_destroy(void) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    dist_release();
    return ZKB_OK;
}

int zkb_dist_buffers(void** d_in_slice, void** d_out_slice, size_t* slice_bytes) {
    DistCtx& d = dctx();
    if (!d.created) { set_error("zkb_dist_create has not been called"); return ZKB_ERR_ARG; }
    if (d_in_slice) *d_in_slice = d.local[0];
    if (d_out_slice) *d_out_slice = d.local[2];
    if (slice_bytes) *slice_bytes = d.slice_bytes;
    return ZKB_OK;
}

int zkb_dist_ntt_fr_dev(const void* d_in_slice, void* d_out_slice, const uint64_t omega[4], uint32_t log_n, void* stream) {
    SlotScope scope(slot_of_device_ptr(d_in_slice ? d_in_slice : d_out_slice));
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!omega) { set_error("omega is NULL"); return ZKB_ERR_ARG; }
    return dist_ntt_dev(d_in_slice, d_out_slice, omega, log_n, (cudaStream_t)stream);
}

int zkb_dist_ntt_fr(const uint64_t* in_slice, uint64_t* out_slice, const uint64_t omega[4], uint32_t log_n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (!in_slice || !out_slice || !omega) { set_error("NULL argument"); return ZKB_ERR_ARG; }
    DistCtx& d = dctx();
    if (!d.connected) { set_error("zkb_dist_connect has not been called"); return ZKB_ERR_ARG; }
    if (log_n > d.max_log_n || log_n < d.log_g) { set_error("log_n %u out of range for this dist context", log_n); return ZKB_ERR_ARG; }
    cudaStream_t s = ctx().stream;
    const size_t slice = ((size_t)1 << (log_n - d.log_g)) * 32;
    ZKB_CUDA_TRY(cudaMemcpyAsync(d.local[0], in_slice, slice, cudaMemcpyHostToDevice, s));
    count_h2d(slice);
    ZKB_TRY(dist_ntt_dev(nullptr, nullptr, omega, log_n, s));
    ZKB_CUDA_TRY(cudaMemcpyAsync(out_slice, d.local[2], slice, cudaMemcpyDeviceToHost, s));
    return dist_check_status(s);
}

int zkb_dist_status(void* stream) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    DistCtx& d = dctx();
    if (!d.created) { set_error("zkb_dist_create has not been called"); return ZKB_ERR_ARG; }
    return dist_check_status((cudaStream_t)stream);
}

}  // extern "C"
