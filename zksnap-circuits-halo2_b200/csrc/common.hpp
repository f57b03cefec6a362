// common.hpp — process-wide context of libzkb200: device binding, error reporting, grow-only workspaces,
// CUDA-event profiling and launch counting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/zkb200.h"

namespace zkb {

void set_error(const char* fmt, ...);

#define ZKB_CUDA_TRY(expr)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ::zkb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA;                        \
        }                                                                                               \
    } while (0)

#define ZKB_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != ZKB_OK) return _rc; \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct ProfTimer {
    double total_ms = 0;
    uint64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct Ctx {
    std::recursive_mutex mu;
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // library-owned stream for the host-buffer entry points
    // profiling
    bool prof_on = false;
    std::map<std::string, ProfTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    std::atomic<uint64_t> launches{0};
    // tuning
    uint32_t msm_c_override = 0;
    uint32_t msm_chunk_override = 0;
};

Ctx& ctx();
int require_init();

struct ProfScope {
    Ctx& c;
    ProfTimer* t = nullptr;
    cudaEvent_t start = nullptr, stop = nullptr;
    cudaStream_t s;
    ProfScope(const char* name, cudaStream_t stream);
    ~ProfScope();
};

inline void count_launch(uint64_t n = 1) { ctx().launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace zkb
