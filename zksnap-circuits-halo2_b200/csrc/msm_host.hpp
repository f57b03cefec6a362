// msm_host.hpp — host-side types of the MSM driver.
#pragma once
#include "common.hpp"
#include "msm.cuh"
#include "msm_plan.hpp"

namespace zkb {

struct MsmWorkspace {
    DevBuf keys[2], vals[2], sort_tmp, buckets, pk[2], pv[2], seg[2];
    void* h_sums = nullptr;  // pinned, 64 XYZZ
};

MsmWorkspace& msm_workspace();
void msm_release_workspace();
void msm_identity_out(uint64_t out[12]);

// sum_i scalars[i] * bases[i]; device pointers; synchronises `s`; result normalised (z = R) on the host
int msm_run(const uint4* d_scalars, const uint4* d_bases, uint64_t n, cudaStream_t s, uint64_t out_jac[12]);
int g1_fixed_base_mul_dev(const uint4* d_scalars, uint64_t n, uint4* d_out, cudaStream_t s);
int measure_imad_peak(double* macs_per_s);
int g1_sum_host(const uint64_t* pts, size_t count, uint64_t out[12]);

}  // namespace zkb
