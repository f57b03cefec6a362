// ntt_host.hpp — host-side types of the NTT driver.
#pragma once
#include <array>

#include "common.hpp"
#include "ntt.cuh"
#include "ntt_plan.hpp"

namespace zkb {

struct NttPlan {
    NttGeometry geom;
    Fr omega;
    DevBuf tw_lo, tw_hi;
    DevBuf tw_r[NTT_MAX_PASSES];
    DevBuf tw_pass[NTT_MAX_PASSES];  // per-pass inter-pass twiddle tables ([K][r]); empty when they did not fit
    bool has_tw_pass = false;
};

struct NttIo {
    const uint4* in = nullptr;   // cols x in_len elements, column stride in_col_stride
    uint64_t in_len = 0;
    uint64_t in_col_stride = 0;
    uint4* work = nullptr;       // cols x N (inner passes run in place here; may alias `in` when in_len == N)
    uint4* out = nullptr;        // cols x N, must differ from `work`
    size_t cols = 1;
    const Fr* in_scale = nullptr;   // 3 constants or NULL
    const Fr* out_scale = nullptr;  // 3 constants or NULL
};

int ntt_get_plan(uint32_t log_n, const uint64_t omega[4], cudaStream_t s, NttPlan** out, bool small_first = false);
int ntt_run(const NttPlan& plan, const NttIo& io, cudaStream_t s);
void ntt_clear_plans();
// one pass of the multi-pass NTT (grid.x CTAs of R*T elements, grid.y columns); used by the sharded driver in dist.cu
int ntt_launch_pass(uint32_t logr, const NttPassArgs& a, dim3 grid, uint32_t threads, size_t smem, cudaStream_t s);
void dist_shutdown();

}  // namespace zkb
