// common.hpp — context of libzkb200: device binding, error reporting, grow-only workspaces, CUDA-event profiling and
// launch counting.
//
// One process drives 1..8 GPUs.  Everything that lives on a GPU (streams, workspaces, twiddle plans, SRS replicas, staging
// pipelines) exists once per DEVICE SLOT; `cur_slot()` is a thread-local index, so the single-device code paths are written
// once and run unchanged on any slot: the calling thread is slot 0 (the home device), slots 1.. are served by one worker
// thread each (`run_on_devices`), and `SlotScope` lets a caller's thread act on another slot for device-pointer calls.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
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
    std::atomic<uint64_t> h2d_bytes{0};   // host->device bytes copied by the library on this device (zkb_transfer_stats)
    // tuning
    uint32_t msm_c_override = 0;
    uint32_t msm_chunk_override = 0;
};

constexpr int ZKB_MAX_DEVICES = 8;
int cur_slot();                       // device slot of the calling thread (0 = home device)
int device_slots();                   // number of bound devices (0 before zkb_init)
int slot_device(int slot);            // CUDA ordinal bound to a slot
// per-device instance of a (default-constructible) singleton, selected by the calling thread's slot
template <class T>
T* per_device_array() {
    static T inst[ZKB_MAX_DEVICES];
    return inst;
}
template <class T>
T& per_device() { return per_device_array<T>()[cur_slot()]; }
// Runs f(slot) for slot = 0 .. count-1: slot 0 on the calling thread, the others on the per-device worker threads, all at the
// same time.  Returns the first failing slot's code with its error text copied to the caller's thread.
int run_on_devices(int count, const std::function<int(int)>& f);
// Makes the calling thread act as `slot` (thread-local slot + cudaSetDevice) until destruction.
struct SlotScope {
    int prev;
    explicit SlotScope(int slot);
    ~SlotScope();
};
// slot that owns a device pointer (0 when there is one device or the pointer is unknown)
int slot_of_device_ptr(const void* p);

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
inline void count_h2d(uint64_t bytes) { ctx().h2d_bytes.fetch_add(bytes, std::memory_order_relaxed); }

}  // namespace zkb
