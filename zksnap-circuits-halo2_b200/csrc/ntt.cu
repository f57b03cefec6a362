// ntt.cu — kernels and host driver of the Fr NTT (see ntt.cuh for the algorithm).
#include <cstdlib>

#include "ntt_host.hpp"

namespace zkb {

#ifndef ZKB_NTT_MIN_CTAS
#define ZKB_NTT_MIN_CTAS 4
#endif
struct CtaBarrier {
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <int LOGR>
__global__ void __launch_bounds__(128, ZKB_NTT_MIN_CTAS) ntt_pass_kernel(const NttPassArgs a) {
    extern __shared__ uint4 ntt_smem[];
    CtaBarrier bar;
    if (a.dist_abort && *reinterpret_cast<const volatile uint32_t*>(a.dist_abort)) return;  // sharded NTT: a barrier timed out
    const uint32_t ncols = a.ncols ? a.ncols : 1;
    ntt_cta_program<LOGR>(a, ntt_smem, threadIdx.x, blockDim.x, (uint64_t)(blockIdx.x / ncols), blockIdx.x % ncols, bar);
}

// out[t] = omega^(t << shift)
__global__ void ntt_pow_table_kernel(uint4* out, Fr omega, uint64_t count, uint32_t shift) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    fr_store2(out, t, fp_pow_u64(omega, t << shift));
}

// out[t] = omega_N^((r*K) << shift), t = K * R + r
__global__ void ntt_pass_table_kernel(uint4* out, const uint4* tw_hi, const uint4* tw_lo, uint32_t tw_h, uint32_t log_r,
                                      uint32_t shift, uint64_t count) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    fr_store2(out, t, ntt_pass_twiddle(tw_hi, tw_lo, tw_h, log_r, shift, t));
}

template <int LOGR>
static int launch_pass(const NttPassArgs& a, dim3 grid, uint32_t threads, size_t smem, cudaStream_t s) {
    struct AttrSet { bool v = false; };  // the attribute belongs to the (function, device) pair
    bool& attr_set = per_device<AttrSet>().v;
    if (!attr_set) {
        ZKB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<LOGR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    ntt_pass_kernel<LOGR><<<grid, threads, smem, s>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

int ntt_launch_pass(uint32_t logr, const NttPassArgs& a, dim3 grid, uint32_t threads, size_t smem, cudaStream_t s) {
    switch (logr) {
#define C(L) case L: return launch_pass<L>(a, grid, threads, smem, s);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10)
#undef C
        default: set_error("unsupported pass radix 2^%u", logr); return ZKB_ERR_ARG;
    }
}

// ---- plan cache: twiddle tables for (log_n, omega) ---------------------------------------------------------------
using PlanCache = std::map<std::pair<uint32_t, std::array<uint64_t, 4>>, NttPlan*>;  // key.first = log_n | small_first << 8
static PlanCache& plan_cache() { return per_device<PlanCache>(); }  // twiddle tables live in one device's HBM

static int build_table(DevBuf& buf, const Fr& w, uint64_t count, uint32_t shift, cudaStream_t s) {
    ZKB_TRY(buf.reserve(count * 32));
    uint32_t threads = 128;
    ntt_pow_table_kernel<<<(unsigned)((count + threads - 1) / threads), threads, 0, s>>>(buf.as<uint4>(), w, count, shift);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

int ntt_get_plan(uint32_t log_n, const uint64_t omega[4], cudaStream_t s, NttPlan** out, bool small_first) {
    if (log_n < 1 || log_n > 28) { set_error("log_n %u out of range [1, 28]", log_n); return ZKB_ERR_ARG; }
    std::array<uint64_t, 4> key{omega[0], omega[1], omega[2], omega[3]};
    auto& cache = plan_cache();
    const uint32_t ckey = log_n | (small_first ? 256u : 0u);
    auto it = cache.find({ckey, key});
    if (it != cache.end()) { *out = it->second; return ZKB_OK; }
    NttPlan* p = new NttPlan();
    p->geom = ntt_geometry(log_n, small_first);
    for (int i = 0; i < 4; ++i) { p->omega.l[2 * i] = (uint32_t)omega[i]; p->omega.l[2 * i + 1] = (uint32_t)(omega[i] >> 32); }
    const uint64_t N = 1ull << log_n;
    int rc = build_table(p->tw_lo, p->omega, 1ull << p->geom.tw_h, 0, s);
    if (rc == ZKB_OK) rc = build_table(p->tw_hi, p->omega, (N >> p->geom.tw_h) ? (N >> p->geom.tw_h) : 1, p->geom.tw_h, s);
    for (uint32_t q = 0; q < p->geom.npass && rc == ZKB_OK; ++q)
        rc = build_table(p->tw_r[q], p->omega, 1ull << p->geom.lr[q], log_n - p->geom.lr[q], s);
    if (rc != ZKB_OK) { delete p; return rc; }
    // per-pass inter-pass twiddle tables (one lookup + one product per element instead of two); sizes Q_{p+1} * 32 B, the last
    // one N * 32 B.  Skipped (the two-level tables remain) when disabled or when HBM is short.
    {
        const char* e = getenv("ZKB_NTT_PASS_TABLES");
        size_t free_b = 0, total_b = 0;
        bool want = !(e && e[0] == '0') && p->geom.npass > 1 && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && N * 32 * 2 < free_b / 4;
        cudaGetLastError();
        uint32_t log_q = p->geom.lr[0];
        for (uint32_t q = 1; q < p->geom.npass && want; ++q) {
            log_q += p->geom.lr[q];  // log2 Q_{q+1}
            const uint64_t count = 1ull << log_q;
            if (p->tw_pass[q].reserve(count * 32) != ZKB_OK) { want = false; break; }
            ntt_pass_table_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(p->tw_pass[q].as<uint4>(), p->tw_hi.as<uint4>(),
                                                                                p->tw_lo.as<uint4>(), p->geom.tw_h, p->geom.lr[q],
                                                                                log_n - log_q, count);
            count_launch();
            if (cudaGetLastError() != cudaSuccess) { want = false; break; }
        }
        if (!want) for (auto& b : p->tw_pass) b.release();
        p->has_tw_pass = want;
    }
    // tables are built on `s`; later launches may use another stream, so make them visible now
    ZKB_CUDA_TRY(cudaStreamSynchronize(s));
    cache[{ckey, key}] = p;
    *out = p;
    return ZKB_OK;
}

void ntt_clear_plans() {
    for (auto& kv : plan_cache()) {
        kv.second->tw_lo.release();
        kv.second->tw_hi.release();
        for (auto& b : kv.second->tw_r) b.release();
        for (auto& b : kv.second->tw_pass) b.release();
        delete kv.second;
    }
    plan_cache().clear();
}

// ---- run -----------------------------------------------------------------------------------------------------------
int ntt_run(const NttPlan& plan, const NttIo& io, cudaStream_t s) {
    const NttGeometry& g = plan.geom;
    const uint64_t N = 1ull << g.log_n;
    if (io.cols == 0) return ZKB_OK;
    if (io.cols > 65535) { set_error("too many columns in one launch: %zu", io.cols); return ZKB_ERR_ARG; }
    if (io.out == io.work || (g.npass == 1 && (const void*)io.in == (const void*)io.out)) {
        set_error("ntt_run: the last pass cannot run in place");
        return ZKB_ERR_ARG;
    }
    ProfScope prof("ntt_pass", s);
    for (uint32_t p = 0; p < g.npass; ++p) {
        NttPassArgs a{};
        const bool fin = p + 1 == g.npass;
        a.src = p == 0 ? io.in : io.work;
        a.src_col_stride = p == 0 ? io.in_col_stride : N;
        a.dst = fin ? io.out : io.work;
        a.dst_col_stride = N;
        a.log_n = g.log_n; a.npass = g.npass; a.pass = p;
        for (uint32_t q = 0; q < g.npass; ++q) a.lr[q] = g.lr[q];
        a.log_t = g.log_t[p];
        a.is_final = fin ? 1 : 0;
        a.tw_r = plan.tw_r[p].as<uint4>(); a.tw_hi = plan.tw_hi.as<uint4>(); a.tw_lo = plan.tw_lo.as<uint4>();
        a.tw_h = g.tw_h;
        a.tw_pass = (plan.has_tw_pass && p > 0) ? plan.tw_pass[p].as<uint4>() : nullptr;
        a.in_len = p == 0 ? io.in_len : N;
        a.in_scale_on = (p == 0 && io.in_scale) ? 1 : 0;
        a.out_scale_on = (fin && io.out_scale) ? 1 : 0;
        for (int m = 0; m < 3; ++m)
            for (int i = 0; i < 8; ++i) {
                a.in_scale[m][i] = io.in_scale ? io.in_scale[m].l[i] : 0;
                a.out_scale[m][i] = io.out_scale ? io.out_scale[m].l[i] : 0;
            }
        a.ncols = (uint32_t)io.cols;
        const uint64_t blocks = ntt_cta_count(g, p) * io.cols;
        if (blocks >= (1ull << 31)) { set_error("NTT batch too large for one launch: %llu CTAs", (unsigned long long)blocks); return ZKB_ERR_ARG; }
        dim3 grid((unsigned)blocks, 1);
        ZKB_TRY(ntt_launch_pass(g.lr[p], a, grid, ntt_cta_threads(g, p), ntt_cta_smem_bytes(g, p), s));
    }
    return ZKB_OK;
}

}  // namespace zkb
