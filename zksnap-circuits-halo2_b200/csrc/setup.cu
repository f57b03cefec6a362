// setup.cu — kernels and host driver of SRS generation (ParamsKZG::setup), fixed-base multiplication and
// Curve::batch_normalize (see setup.cuh).
#include <cstring>

#include "msm_host.hpp"
#include "setup.cuh"

namespace zkb {

__global__ void __launch_bounds__(128) fb_table_scalar_kernel(const FbScalarArgs a) {
    fb_table_scalar_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) fb_mul_kernel(const FbMulArgs a) {
    fb_mul_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) g1_batch_to_affine_kernel(const BatchAffineArgs a) {
    g1_batch_to_affine_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) fr_powers_kernel(const FrPowersArgs a) {
    fr_powers_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) lagrange_scalars_kernel(const LagrangeScalarArgs a) {
    lagrange_scalars_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(128) g1_bitrev_kernel(const G1BitrevArgs a) {
    g1_bitrev_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) g1_fft_stage_kernel(const G1FftStageArgs a) {
    g1_fft_stage_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) fr_pow_canon_kernel(const FrPowCanonArgs a) {
    fr_pow_canon_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) g1_scale_kernel(const G1ScaleArgs a) {
    g1_scale_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) g1_lift_kernel(const G1LiftArgs a) {
    g1_lift_thread(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

static inline unsigned nblocks(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

struct SetupWs {
    DevBuf table;      // fixed-base table of G (built once)
    bool table_ready = false;
    DevBuf xyzz, scalars, points;
    uint32_t* d_status = nullptr;
};
static SetupWs& setup_ws() { return per_device<SetupWs>(); }
__global__ void __launch_bounds__(128) g1_on_curve_kernel(const OnCurveArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i >= a.n || g1_affine_is_valid(affine_load(a.pts + 4 * i));
    const unsigned bad = __popc(__ballot_sync(0xFFFFFFFFu, !ok));
    if (bad && (threadIdx.x & 31) == 0) atomicAdd(a.bad, (unsigned long long)bad);
}

// number of points of d_aff[0..n) that are neither the identity nor canonical points of the curve -> *bad (synchronises `s`)
int g1_check_on_curve_dev(const uint4* d_aff, uint64_t n, uint64_t* bad, cudaStream_t s) {
    *bad = 0;
    if (n == 0) return ZKB_OK;
    unsigned long long* d_bad = nullptr;
    ZKB_CUDA_TRY(cudaMalloc(&d_bad, 8));
    cudaError_t e = cudaMemsetAsync(d_bad, 0, 8, s);
    if (e == cudaSuccess) {
        OnCurveArgs a{d_aff, n, d_bad};
        g1_on_curve_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
        count_launch();
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_bad, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_bad);
    if (e != cudaSuccess) { set_error("on-curve check failed: %s", cudaGetErrorString(e)); return ZKB_ERR_CUDA; }
    *bad = h;
    return ZKB_OK;
}

void setup_release() {
    SetupWs& w = setup_ws();
    w.table.release(); w.xyzz.release(); w.scalars.release(); w.points.release();
    if (w.d_status) cudaFree(w.d_status);
    w.d_status = nullptr;
    w.table_ready = false;
}

// n points (XYZZ or Jacobian) -> affine, FB_BATCH per thread
int g1_batch_to_affine_dev(const uint4* d_in, uint64_t n, uint4* d_out, bool jacobian, cudaStream_t s) {
    if (n == 0) return ZKB_OK;
    BatchAffineArgs a{};
    a.in = d_in; a.n = n; a.out = d_out; a.jacobian = jacobian ? 1 : 0;
    a.nthreads = (n + FB_BATCH - 1) / FB_BATCH;
    g1_batch_to_affine_kernel<<<nblocks(a.nthreads, 128), 128, 0, s>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return ZKB_OK;
}

// T[w][d-1] = [d 2^(16 w)] G, built once with the naive double-and-add kernel (an independent code path, so the windowed
// kernel is cross-checked against it in the tests)
static int fb_table_ready(cudaStream_t s) {
    SetupWs& w = setup_ws();
    if (w.table_ready) return ZKB_OK;
    const uint64_t count = (uint64_t)FB_WINDOWS * FB_ROW;
    ZKB_TRY(w.table.reserve(count * 64));
    DevBuf sc;
    ZKB_TRY(sc.reserve(count * 32));
    FbScalarArgs a{sc.as<uint4>()};
    fb_table_scalar_kernel<<<nblocks(count, 128), 128, 0, s>>>(a);
    count_launch();
    int rc = cudaGetLastError() == cudaSuccess ? ZKB_OK : ZKB_ERR_CUDA;
    if (rc == ZKB_OK) rc = g1_fixed_base_mul_dev(sc.as<uint4>(), count, w.table.as<uint4>(), s);
    if (rc == ZKB_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = ZKB_ERR_CUDA;
    sc.release();
    if (rc != ZKB_OK) { cudaGetLastError(); set_error("building the fixed-base table failed"); return rc; }
    w.table_ready = true;
    return ZKB_OK;
}

// out[i] = [s_i] G (affine), windowed
int g1_fixed_base_window_dev(const uint4* d_scalars, uint64_t n, uint4* d_out, cudaStream_t s) {
    if (n == 0) return ZKB_OK;
    SetupWs& w = setup_ws();
    ZKB_TRY(fb_table_ready(s));
    ZKB_TRY(w.xyzz.reserve(n * 128));
    FbMulArgs a{d_scalars, n, w.table.as<uint4>(), w.xyzz.as<uint4>()};
    fb_mul_kernel<<<nblocks(n, 128), 128, 0, s>>>(a);
    count_launch();
    ZKB_CUDA_TRY(cudaGetLastError());
    return g1_batch_to_affine_dev(w.xyzz.as<uint4>(), n, d_out, false, s);
}

static void fr_words(const Fr& v, uint32_t (&o)[8]) {
    for (int i = 0; i < 8; ++i) o[i] = v.l[i];
}
static Fr fr_from_u64x4(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}

// d_g / d_gl: n affine points each (either may be NULL)
int kzg_setup_dev(uint32_t k, const uint64_t s_mont[4], uint4* d_g, uint4* d_gl, cudaStream_t st) {
    SetupWs& w = setup_ws();
    const uint64_t n = 1ull << k;
    const Fr s = fr_from_u64x4(s_mont);
    ZKB_TRY(w.scalars.reserve(n * 32));
    if (d_g) {
        FrPowersArgs a{};
        a.out = w.scalars.as<uint4>(); a.n = n;
        fr_words(s, a.s);
        fr_powers_kernel<<<nblocks((n + SETUP_CHUNK - 1) / SETUP_CHUNK, 128), 128, 0, st>>>(a);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        ZKB_TRY(g1_fixed_base_window_dev(w.scalars.as<uint4>(), n, d_g, st));
    }
    if (d_gl) {
        if (!w.d_status) ZKB_CUDA_TRY(cudaMalloc(&w.d_status, 4));
        ZKB_CUDA_TRY(cudaMemsetAsync(w.d_status, 0, 4, st));
        // omega = ROOT_OF_UNITY^(2^(28-k)); c = (s^n - 1) / n  (host: a handful of products)
        uint64_t om[4];
        if (zkb_fr_omega(k, om) != ZKB_OK) return ZKB_ERR_ARG;
        Fr sn = s;
        for (uint32_t i = 0; i < k; ++i) sn = fp_sqr(sn);
        Fr nn = Fr::one();
        for (uint32_t i = 0; i < k; ++i) nn = fp_dbl(nn);
        const Fr c = fp_mul(fp_sub(sn, Fr::one()), fp_inv(nn));
        LagrangeScalarArgs a{};
        a.out = w.scalars.as<uint4>(); a.n = n; a.status = w.d_status;
        fr_words(s, a.s); fr_words(fr_from_u64x4(om), a.omega); fr_words(c, a.c);
        lagrange_scalars_kernel<<<nblocks((n + SETUP_CHUNK - 1) / SETUP_CHUNK, 128), 128, 0, st>>>(a);
        count_launch();
        ZKB_CUDA_TRY(cudaGetLastError());
        uint32_t h_status = 0;
        ZKB_CUDA_TRY(cudaMemcpyAsync(&h_status, w.d_status, 4, cudaMemcpyDeviceToHost, st));
        ZKB_CUDA_TRY(cudaStreamSynchronize(st));
        if (h_status) { set_error("setup: s is an n-th root of unity (upstream panics on the failed inversion)"); return ZKB_ERR_ARG; }
        ZKB_TRY(g1_fixed_base_window_dev(w.scalars.as<uint4>(), n, d_gl, st));
    }
    return ZKB_OK;
}

// best_fft over G1: d_aff_in n affine points -> d_aff_out n affine points, out[i] = sum_j [omega^(i j)] in[j], optionally
// followed by the scalar `scale` (canonical-ised here) on every output — g_to_lagrange = (omega^-1, 1/n).
int g1_fft_dev(const uint4* d_aff_in, uint4* d_aff_out, uint32_t log_n, const Fr& omega, const Fr* scale, cudaStream_t st) {
    SetupWs& w = setup_ws();
    const uint64_t n = 1ull << log_n;
    ZKB_TRY(w.xyzz.reserve(n * 128));
    ZKB_TRY(w.scalars.reserve((n / 2 ? n / 2 : 1) * 32));
    uint4* a = w.xyzz.as<uint4>();
    G1LiftArgs la{d_aff_in, a, n};
    g1_lift_kernel<<<nblocks(n, 128), 128, 0, st>>>(la);
    count_launch();
    if (log_n >= 1) {
        FrPowCanonArgs pa{};
        pa.out = w.scalars.as<uint4>(); pa.n = n / 2;
        fr_words(omega, pa.w);
        fr_pow_canon_kernel<<<nblocks((n / 2 + SETUP_CHUNK - 1) / SETUP_CHUNK, 128), 128, 0, st>>>(pa);
        count_launch();
        G1BitrevArgs ba{a, log_n};
        g1_bitrev_kernel<<<nblocks(n, 128), 128, 0, st>>>(ba);
        count_launch();
        for (uint32_t stage = 1; stage <= log_n; ++stage) {
            G1FftStageArgs sa{a, w.scalars.as<uint4>(), log_n, stage};
            g1_fft_stage_kernel<<<nblocks(n / 2, 128), 128, 0, st>>>(sa);
            count_launch();
        }
    }
    if (scale) {
        G1ScaleArgs ga{};
        ga.a = a; ga.n = n;
        fr_words(fp_from_mont(*scale), ga.k);
        g1_scale_kernel<<<nblocks(n, 128), 128, 0, st>>>(ga);
        count_launch();
    }
    ZKB_CUDA_TRY(cudaGetLastError());
    return g1_batch_to_affine_dev(a, n, d_aff_out, false, st);
}

Fr fr_from_limbs_u64(const uint64_t* p) { return fr_from_u64x4(p); }

}  // namespace zkb
