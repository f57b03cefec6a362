// msm_host.hpp — host-side types of the MSM driver.
#pragma once
#include "common.hpp"
#include "msm.cuh"
#include "msm_plan.hpp"

namespace zkb {

struct MsmWorkspace {
    // keys / vals / sort_tmp / counter exist twice: a sliced MSM prepares (digits + sort) slice i+1 on `aux` while slice i
    // accumulates on the caller's stream; set = slice parity.  Whole MSMs use set 0.
    DevBuf keys[4], vals[4], sort_tmp[2], counter[2], buckets, pk[2], pv[2], seg[2];
    void* h_sums = nullptr;  // pinned staging for the per-set sums / finished results
    size_t h_sums_cap = 0;
    void* h_cnt = nullptr;   // pinned, 2 x 8 B: entry counts of the two sets
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_sorted[2] = {nullptr, nullptr}, ev_acc[2] = {nullptr, nullptr};
    bool acc_pending[2] = {false, false};  // ev_acc[set] recorded in the current MSM
    uint32_t slice_ix = 0;
    int overlap = 1;         // ZKB_MSM_OVERLAP=0 (environment, read once): run the slices on one stream
    bool env_read = false;
    uint64_t last_entries = 0;  // non-zero digits (bucket additions) of the last MSM / slice on this device
};

MsmWorkspace& msm_workspace();
void msm_release_workspace();
void msm_identity_out(uint64_t out[12]);

// precomputed SRS window table: rows[w * row_stride + i] = 2^(c w) * P_i
struct MsmTable {
    const uint4* rows;     // already offset to the first point of the MSM range
    uint64_t row_stride;   // registered SRS length
    uint32_t c;
    uint32_t nwin;
};

// sum_i scalars[i] * bases[i]; device pointers; synchronises `s`; result normalised (z = R) on the host.
// With `table` the bases come from the window table and all windows share one bucket set.
// `ncols` independent scalar columns (column-major contiguous) against the same bases are accumulated in ONE pass
// (column index folded into the bucket key); out_jac receives ncols x 12 limbs.
// `phase`: a pipelined MSM is run as point-range slices that share one bucket array (table mode, same c): the first
// slice (MSM_FIRST) clears the buckets, later ones add into them, the last one (MSM_LAST) reduces and synchronises;
// slices without MSM_LAST return after enqueueing their kernels.
enum : uint32_t { MSM_FIRST = 1, MSM_LAST = 2, MSM_WHOLE = 3 };
// `input_ready` (slices only): event after which the slice's scalars are in HBM.  When given, digits + sort of the slice run on a
// second stream that waits for it directly (an event recorded on `s` would also wait for the previous slice's accumulation) and
// overlap that accumulation; NULL keeps the slice on `s`.  Measured (profiles/r1_tuning.txt §12): the sort competes with the
// accumulation for the SMs, so the gain is small (2^24 from page-locked scalars 43.1 -> 42.4 ms) and with pageable scalars, where
// host staging threads set the pace, it is a loss — the caller passes the event only for page-locked memory.
int msm_run(const uint4* d_scalars, const uint4* d_bases, uint64_t n, cudaStream_t s, uint64_t* out_jac,
            const MsmTable* table = nullptr, uint32_t ncols = 1, uint32_t phase = MSM_WHOLE, cudaEvent_t input_ready = nullptr);
int srs_table_build(const uint4* d_bases, uint64_t n, uint32_t c, uint32_t nwin, uint4* d_table, cudaStream_t s);
int g1_fixed_base_mul_dev(const uint4* d_scalars, uint64_t n, uint4* d_out, cudaStream_t s);       // naive double-and-add
// setup.cu
int g1_fixed_base_window_dev(const uint4* d_scalars, uint64_t n, uint4* d_out, cudaStream_t s);  // 16-bit windows over a table of G
int g1_batch_to_affine_dev(const uint4* d_in, uint64_t n, uint4* d_out, bool jacobian, cudaStream_t s);
int kzg_setup_dev(uint32_t k, const uint64_t s_mont[4], uint4* d_g, uint4* d_gl, cudaStream_t st);
int g1_fft_dev(const uint4* d_aff_in, uint4* d_aff_out, uint32_t log_n, const Fr& omega, const Fr* scale, cudaStream_t st);
int g1_check_on_curve_dev(const uint4* d_aff, uint64_t n, uint64_t* bad, cudaStream_t s);
void setup_release();
int measure_imad_peak(double* macs_per_s);
int g1_sum_host(const uint64_t* pts, size_t count, uint64_t out[12]);

}  // namespace zkb
