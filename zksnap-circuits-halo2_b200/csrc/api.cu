// api.cu — the C ABI of libzkb200.so (include/zkb200.h): context, SRS registry, host-buffer and device-buffer
// entry points.  No CPU fallback lives here: every compute call requires an initialised CUDA device.
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <thread>

#include "msm_host.hpp"
#include "ntt_host.hpp"

namespace zkb {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

// ---- device slots -----------------------------------------------------------------------------------------------------------
static thread_local int tl_slot = 0;
static int g_nslots = 0;                      // bound devices
static int g_slot_device[ZKB_MAX_DEVICES] = {0};
static std::mutex g_fanout_mu;                // one fan-out at a time (the workers hold one job each)

int cur_slot() { return tl_slot; }
int device_slots() { return g_nslots; }
int slot_device(int slot) { return g_slot_device[slot]; }

Ctx& ctx() { return per_device<Ctx>(); }

SlotScope::SlotScope(int slot) : prev(tl_slot) {
    tl_slot = slot;
    if (slot != prev) cudaSetDevice(g_slot_device[slot]);
}
SlotScope::~SlotScope() {
    if (tl_slot != prev) cudaSetDevice(g_slot_device[prev]);
    tl_slot = prev;
}

int slot_of_device_ptr(const void* p) {
    if (g_nslots <= 1 || !p) return tl_slot;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return tl_slot; }
    if (at.type != cudaMemoryTypeDevice) return tl_slot;
    for (int i = 0; i < g_nslots; ++i)
        if (g_slot_device[i] == at.device) return i;
    return tl_slot;
}

// one worker thread per device slot >= 1: bound to its GPU for life, runs the jobs of run_on_devices
struct DevWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = ZKB_OK;
    std::string err;

    void start(int slot) {
        if (th.joinable()) return;
        quit = false;
        th = std::thread([this, slot] {
            tl_slot = slot;
            cudaSetDevice(g_slot_device[slot]);
            for (;;) {
                std::function<int()> j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return quit || has_job; });
                    if (quit) return;
                    j = std::move(job);
                    has_job = false;
                }
                g_error.clear();
                int r = j();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    rc = r;
                    err = g_error;
                    done = true;
                }
                cv.notify_all();
            }
        });
    }
    void post(std::function<int()> j) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j);
            has_job = true;
            done = false;
        }
        cv.notify_all();
    }
    int wait(std::string* e) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        if (rc != ZKB_OK && e) *e = err;
        return rc;
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        th.join();
    }
};
static DevWorker g_workers[ZKB_MAX_DEVICES];

int run_on_devices(int count, const std::function<int(int)>& f) {
    if (count <= 1 || tl_slot != 0) return f(tl_slot);
    std::lock_guard<std::mutex> fan(g_fanout_mu);
    for (int i = 1; i < count; ++i) g_workers[i].post([&f, i] { return f(i); });
    int rc = f(0);
    std::string first_err = rc != ZKB_OK ? g_error : std::string();
    for (int i = 1; i < count; ++i) {
        std::string e;
        int r = g_workers[i].wait(&e);
        if (r != ZKB_OK && rc == ZKB_OK) { rc = r; first_err = "device slot " + std::to_string(i) + ": " + e; }
    }
    if (rc != ZKB_OK) g_error = first_err;
    return rc;
}

void poly_pool_flush();  // poly.cu

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return ZKB_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + (bytes >> 3);  // slack so slowly growing sizes do not reallocate every call
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {  // give back the cached polynomial buffers (poly.cu) and try once more
        cudaGetLastError();
        poly_pool_flush();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return ZKB_ERR_OOM;
    }
    cap = want;
    return ZKB_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

int require_init() {
    Ctx& c = ctx();
    if (c.inited) {
        cudaSetDevice(c.device);
        return ZKB_OK;
    }
    if (cur_slot() != 0) { set_error("device slot %d is not initialised", cur_slot()); return ZKB_ERR_NO_DEVICE; }
    return zkb_init(nullptr, 0);
}

ProfScope::ProfScope(const char* name, cudaStream_t stream) : c(ctx()), s(stream) {
    if (!c.prof_on) return;
    t = &c.timers[name];
    auto get = [&]() {
        cudaEvent_t e;
        if (!c.event_pool.empty()) { e = c.event_pool.back(); c.event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    start = get();
    stop = get();
    cudaEventRecord(start, s);
}
ProfScope::~ProfScope() {
    if (!t) return;
    cudaEventRecord(stop, s);
    t->pending.emplace_back(start, stop);
    t->launches++;
}

// ---- host-side field constants (EvaluationDomain::new) ---------------------------------------------------------------
static Fr fr_from_limbs64(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}
static void fr_to_limbs64(const Fr& v, uint64_t* p) {
    for (int i = 0; i < 4; ++i) p[i] = (uint64_t)v.l[2 * i] | ((uint64_t)v.l[2 * i + 1] << 32);
}
static Fr fr_root_of_unity() {  // Fr::ROOT_OF_UNITY = 7^((r-1)/2^28), Montgomery form
    const uint32_t w[8] = {0xb639feb8u, 0x9632c7c5u, 0x0d0ff299u, 0x985ce340u, 0x01b0ecd8u, 0xb2dd8800u, 0x6d98ce29u, 0x1d69070du};
    return fr_from_words(w);
}
static Fr fr_zeta() {  // Fr::ZETA, Montgomery form
    const uint32_t w[8] = {0x55fcd653u, 0x0363f299u, 0x5fc1e200u, 0x73e7950bu, 0x576d9d24u, 0xc5fce83eu, 0xa1c3a4d4u, 0x059c805du};
    return fr_from_words(w);
}
static Fr fr_inv_host(const Fr& a) {  // a^(r-2)
    Fr acc = Fr::one();
    for (int i = 255; i >= 0; --i) {
        acc = fp_sqr(acc);
        uint32_t limb = FrParams::M(i >> 5);
        if (i < 32) limb -= 2;  // r - 2: low limb 0xf0000001 - 2 borrows nothing
        if ((limb >> (i & 31)) & 1) acc = fp_mul(acc, a);
    }
    return acc;
}
// omega(k), omega(k)^-1 and (2^k)^-1 for k = 0 .. 28, built ONCE (two host inversions): lagrange_to_coeff / extended_to_coeff used to
// run two 256-step Fermat inversions on the portable 32-bit chains per CALL — ~0.1–0.25 ms of host time in front of transforms
// that take 20 us at the voter's k = 13
struct DomainConsts {
    Fr omega[29], omega_inv[29], pow2_inv[29];
    DomainConsts() {
        omega[28] = fr_root_of_unity();
        for (int k = 27; k >= 0; --k) omega[k] = fp_sqr(omega[k + 1]);
        omega_inv[28] = fr_inv_host(omega[28]);
        for (int k = 27; k >= 0; --k) omega_inv[k] = fp_sqr(omega_inv[k + 1]);
        const Fr inv2 = fr_inv_host(fp_dbl(Fr::one()));
        pow2_inv[0] = Fr::one();
        for (int k = 1; k <= 28; ++k) pow2_inv[k] = fp_mul(pow2_inv[k - 1], inv2);
    }
};
static const DomainConsts& domain_consts() {
    static const DomainConsts c;   // thread-safe one-time initialisation
    return c;
}
static Fr fr_omega_host(uint32_t k) { return domain_consts().omega[k <= 28 ? k : 28]; }
static Fr fr_omega_inv_host(uint32_t k) { return domain_consts().omega_inv[k <= 28 ? k : 28]; }
static Fr fr_pow2_inv_host(uint32_t k) { return domain_consts().pow2_inv[k <= 28 ? k : 28]; }  // (2^k)^-1

// ---- SRS registry ---------------------------------------------------------------------------------------------------------
struct Srs {
    DevBuf bases;
    size_t n = 0;
    // window table T[w][i] = 2^(c w) P_i, built lazily on the first commit (kept for the life of the handle)
    DevBuf table;
    uint32_t table_c = 0, table_nwin = 0;
    bool table_failed = false;
    uint64_t commits = 0;  // commitments served so far (automatic mode decides on this)
    uint64_t plan_n = 0;   // MSM length the window table is tuned for: n, or n / devices when commits are sharded by point range
};
// One registered SRS: a replica per device slot (the arrays are small next to 180 GB of HBM: 256 MiB at k = 22), so that one
// commit can be sharded by point range over the devices and a batch of commits by column, without moving bases per call.
struct SrsSet {
    Srs* dev[ZKB_MAX_DEVICES] = {};
    size_t n = 0;
};
// SRS window-table policy: 0 off, 1 eager (build on the first commit), 2 automatic (default): build once the handle has served
// PRECOMPUTE_AFTER commitments.  A table-mode commit is ~18 % faster (2^22: 11.5 vs 14.0 ms) and the build costs ~100 commits'
// worth of that gain at every size (2^22: 0.33 s), so a process that proves once should not pay for it while a long-running
// prover should — buying after that many "rents" is the 2-competitive ski-rental rule.  zkb_srs_precompute forces the build.
static int g_precompute = -1;  // -1: read ZKB_SRS_PRECOMPUTE on first use
constexpr uint64_t PRECOMPUTE_AFTER = 96;

static int precompute_mode() {
    if (g_precompute < 0) {
        const char* e = getenv("ZKB_SRS_PRECOMPUTE");
        g_precompute = !e ? 2 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2));
    }
    return g_precompute;
}

// Build (once) the window table of an SRS; returns false when disabled or when it does not fit comfortably in HBM.
static bool srs_table_ready(Srs* s, cudaStream_t stream, bool force = false) {
    if (s->table_c) return true;
    const int mode = precompute_mode();
    if (mode == 0 || s->table_failed || s->n < 64 || ctx().msm_c_override) return false;
    if (!force && mode == 2 && s->commits < PRECOMPUTE_AFTER) return false;
    MsmGeometry g = msm_geometry(s->plan_n ? s->plan_n : s->n, 0, 0, true);
    size_t bytes = (size_t)g.nwin * s->n * 64;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes > free_b / 3 || (uint64_t)g.nwin * s->n >= (1ull << 31)) {
        cudaGetLastError();
        s->table_failed = true;
        return false;
    }
    if (s->table.reserve(bytes) != ZKB_OK) { s->table_failed = true; return false; }
    if (srs_table_build(s->bases.as<uint4>(), s->n, g.c, g.nwin, s->table.as<uint4>(), stream) != ZKB_OK ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        cudaGetLastError();
        s->table.release();
        s->table_failed = true;
        return false;
    }
    s->table_c = g.c;
    s->table_nwin = g.nwin;
    return true;
}
static std::map<uint64_t, SrsSet*>& srs_map() {
    static std::map<uint64_t, SrsSet*> m;
    return m;
}
static std::mutex g_srs_mu;  // guards srs_map and g_next_handle (entry points of different slots may run concurrently)
static uint64_t g_next_handle = 1;

// Multi-device thresholds (log2): smallest per-device share of points worth sharding a commit for; smallest transform whose
// batch is split by column; smallest single transform sharded over the devices.  Environment ZKB_MULTI_MIN_SHARE_LOG /
// ZKB_MULTI_NTT_MIN_LOG / ZKB_MULTI_DIST_MIN_LOG, or zkb_multi_device_set.
static int g_multi[3] = {-1, -1, -1};
static int multi_threshold(int which) {
    static const char* const names[3] = {"ZKB_MULTI_MIN_SHARE_LOG", "ZKB_MULTI_NTT_MIN_LOG", "ZKB_MULTI_DIST_MIN_LOG"};
    static const int defaults[3] = {16, 16, 22};
    if (g_multi[which] < 0) {
        const char* e = getenv(names[which]);
        g_multi[which] = e ? atoi(e) : defaults[which];
        if (g_multi[which] < 1) g_multi[which] = 1;
    }
    return g_multi[which];
}
static int msm_min_share_log() { return multi_threshold(0); }
// devices a commit of n points is sharded over (by SRS point range)
static int msm_fanout(size_t n) {
    if (g_nslots <= 1 || cur_slot() != 0) return 1;
    size_t g = n >> msm_min_share_log();
    if (g > (size_t)g_nslots) g = (size_t)g_nslots;
    return g < 1 ? 1 : (int)g;
}
static void point_range(size_t n, int part, int parts, size_t* off, size_t* len) {
    const size_t base = n / parts, rem = n % parts;
    *off = (size_t)part * base + ((size_t)part < rem ? (size_t)part : rem);
    *len = base + ((size_t)part < rem ? 1 : 0);
}

// Publishes an SRS that is resident on the calling slot: replicates it to every other device slot (peer copies over NVLink,
// all devices at once) and returns the handle.  Takes ownership of `home` (released on failure).
static int srs_publish(Srs* home, uint64_t* handle) {
    SrsSet* set = new SrsSet();
    set->n = home->n;
    const int hs = cur_slot();
    const size_t bytes = home->n ? home->n * 64 : 64;
    home->plan_n = (g_nslots > 1 && (home->n >> msm_min_share_log()) >= (size_t)g_nslots) ? home->n / g_nslots : home->n;
    set->dev[hs] = home;
    int rc = ZKB_OK;
    if (g_nslots > 1 && hs == 0) {
        if (cudaStreamSynchronize(ctx().stream) != cudaSuccess) { cudaGetLastError(); set_error("SRS publish: stream sync failed"); rc = ZKB_ERR_CUDA; }
        if (rc == ZKB_OK) rc = run_on_devices(g_nslots, [&](int slot) -> int {
            if (slot == hs) return ZKB_OK;
            Srs* r = new Srs();
            r->n = home->n;
            r->plan_n = home->plan_n;
            int rr = r->bases.reserve(bytes);
            if (rr == ZKB_OK && home->n) {
                cudaError_t e = cudaMemcpyPeer(r->bases.p, slot_device(slot), home->bases.p, slot_device(hs), home->n * 64);
                if (e != cudaSuccess) { cudaGetLastError(); set_error("SRS replication to device %d failed: %s", slot_device(slot), cudaGetErrorString(e)); rr = ZKB_ERR_CUDA; }
            }
            if (rr != ZKB_OK) { r->bases.release(); delete r; return rr; }
            set->dev[slot] = r;
            return ZKB_OK;
        });
    }
    if (rc != ZKB_OK) {
        for (int i = 0; i < ZKB_MAX_DEVICES; ++i)
            if (set->dev[i]) { SlotScope sc(i); set->dev[i]->bases.release(); delete set->dev[i]; }
        delete set;
        return rc;
    }
    std::lock_guard<std::mutex> lk(g_srs_mu);
    *handle = g_next_handle++;
    srs_map()[*handle] = set;
    return ZKB_OK;
}

struct PinBuf {  // grow-only pinned host buffer
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ZKB_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ZKB_ERR_OOM; }
        cap = bytes;
        return ZKB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct HostIo {  // device staging for the host-buffer entry points
    DevBuf scalars, bases, x, y, in;
    PinBuf stage[2];            // pinned staging of pageable scalar slices
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};
static HostIo& hostio() { return per_device<HostIo>(); }

static int check_ptr(const void* p, const char* what) {
    if (!p) { set_error("%s is NULL", what); return ZKB_ERR_ARG; }
    return ZKB_OK;
}

enum DomainOp { OP_FFT, OP_L2C, OP_C2L, OP_C2E, OP_E2C };

// One NTT-family operation over `ncols` columns resident on the device.
//   d_in: columns of in_len elements; result (N elements per column) ends in d_a; d_b is scratch of the same size.
//   For OP_C2E d_in is separate from d_a; otherwise d_in == d_a.
static int domain_op_dev(DomainOp op, const uint4* d_in, uint4* d_a, uint4* d_b, size_t ncols, uint32_t k, uint32_t ek,
                         const uint64_t* omega_user, cudaStream_t s) {
    uint32_t log_n = (op == OP_C2E || op == OP_E2C) ? ek : k;
    if (log_n < 1 || log_n > 28 || k > log_n) { set_error("bad domain sizes k=%u extended_k=%u", k, ek); return ZKB_ERR_ARG; }
    uint64_t om[4];
    Fr scale[3];
    const Fr* in_scale = nullptr;
    const Fr* out_scale = nullptr;
    switch (op) {
        case OP_FFT: memcpy(om, omega_user, 32); break;
        case OP_C2L: fr_to_limbs64(fr_omega_host(k), om); break;
        case OP_L2C: {
            fr_to_limbs64(fr_omega_inv_host(k), om);
            scale[0] = scale[1] = scale[2] = fr_pow2_inv_host(k);
            out_scale = scale;
            break;
        }
        case OP_C2E: {
            fr_to_limbs64(fr_omega_host(ek), om);
            scale[0] = Fr::one(); scale[1] = fr_zeta(); scale[2] = fp_sqr(scale[1]);
            in_scale = scale;
            break;
        }
        case OP_E2C: {
            fr_to_limbs64(fr_omega_inv_host(ek), om);
            Fr ninv = fr_pow2_inv_host(ek), z = fr_zeta(), z2 = fp_sqr(z);
            scale[0] = ninv; scale[1] = fp_mul(ninv, z2); scale[2] = fp_mul(ninv, z);
            out_scale = scale;
            break;
        }
    }
    NttPlan* plan = nullptr;
    ZKB_TRY(ntt_get_plan(log_n, om, s, &plan));
    const uint64_t N = 1ull << log_n;
    NttIo io;
    io.in = d_in;
    io.in_len = op == OP_C2E ? (1ull << k) : N;
    io.in_col_stride = io.in_len;
    io.cols = ncols;
    io.in_scale = in_scale;
    io.out_scale = out_scale;
    if (plan->geom.npass == 1) {
        if (op == OP_C2E) { io.work = d_b; io.out = d_a; return ntt_run(*plan, io, s); }
        io.work = d_a; io.out = d_b;
        ZKB_TRY(ntt_run(*plan, io, s));
        ZKB_CUDA_TRY(cudaMemcpyAsync(d_a, d_b, ncols * N * 32, cudaMemcpyDeviceToDevice, s));
        return ZKB_OK;
    }
    io.work = d_b;  // first pass d_in -> d_b, inner passes in place in d_b, last pass d_b -> d_a
    io.out = d_a;
    return ntt_run(*plan, io, s);
}

// ---- host-buffer scheduler (north_star item 4): pinned staging, three streams, column groups in flight ---------------------
// The host-buffer entry points are synchronous (halo2 semantics), so the overlap lives inside one batched call: the
// columns are cut into groups; group i+1 is copied host->device on `s_h2d` while group i runs on the compute stream
// and group i-1 is copied device->host on `s_d2h` (PCIe is full duplex).  Pageable caller memory (a Rust Vec) is
// staged through per-slot pinned buffers by a small pool of host threads; pinned / registered caller memory is
// DMA'd directly.
struct HostCopy { void* dst; const void* src; size_t bytes; };
struct HostPool {  // parallel memcpy for the pageable <-> pinned staging copies
    std::vector<std::thread> workers;
    const HostCopy* list = nullptr;   // copy_many: the threads split the CONCATENATED byte range of the list
    size_t list_count = 0;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t gen = 0;
    unsigned pending = 0;
    bool quit = false;
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0;
    unsigned nparts = 0;

    void start() {
        if (!workers.empty()) return;
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw / 4;  // two pools (stage-in, stage-out) share the cores with the caller's own threads
        if (n < 2) n = 2;
        if (n > 8) n = 8;
        const char* e = getenv("ZKB_STAGE_THREADS");
        if (e && atoi(e) > 0) n = (unsigned)atoi(e);
        nparts = n;
        const uint64_t g0 = gen;  // a pool restarted after zkb_shutdown must not replay the last job of its previous life
        for (unsigned i = 1; i < n; ++i) workers.emplace_back([this, i, g0] { loop(i, g0); });
    }
    void part(unsigned i) {
        size_t per = ((bytes / nparts) + 4095) & ~(size_t)4095;
        size_t lo = (size_t)i * per, hi = lo + per;
        if (lo >= bytes) return;
        if (hi > bytes) hi = bytes;
        if (!list) { memcpy(dst + lo, src + lo, hi - lo); return; }
        size_t off = 0;   // start of the current list entry in the concatenation
        for (size_t e = 0; e < list_count && off < hi; ++e) {
            const size_t b = list[e].bytes;
            if (off + b > lo) {
                const size_t from = lo > off ? lo - off : 0, to = hi - off < b ? hi - off : b;
                memcpy((char*)list[e].dst + from, (const char*)list[e].src + from, to - from);
            }
            off += b;
        }
    }
    void loop(unsigned i, uint64_t seen) {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return quit || gen != seen; });
                if (quit) return;
                seen = gen;
            }
            part(i);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    // many copies as ONE parallel job: a batch of short columns (hundreds of 256 KiB .. 1 MiB arrays for the voter circuit) would
    // otherwise be copied one column at a time by one thread each (measured: a pageable 256-column coeff_to_extended batch 34 ms
    // against 7 ms from page-locked memory)
    void copy_many(const HostCopy* l, size_t count) {
        size_t total = 0;
        for (size_t e = 0; e < count; ++e) total += l[e].bytes;
        if (total >= (4u << 20)) start();
        if (total < (4u << 20) || nparts <= 1) {
            for (size_t e = 0; e < count; ++e) memcpy(l[e].dst, l[e].src, l[e].bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            list = l; list_count = count; bytes = total;
            pending = nparts - 1;
            ++gen;
        }
        cv_go.notify_all();
        part(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        list = nullptr; list_count = 0;
    }
    void copy(void* d, const void* s, size_t n) {
        if (n < (4u << 20)) { memcpy(d, s, n); return; }
        start();
        if (nparts <= 1) { memcpy(d, s, n); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            list = nullptr; list_count = 0;
            dst = (char*)d; src = (const char*)s; bytes = n;
            pending = nparts - 1;
            ++gen;
        }
        cv_go.notify_all();
        part(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~HostPool() { stop(); }
    void stop() {
        if (workers.empty()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_go.notify_all();
        for (auto& t : workers) t.join();
        workers.clear();
        quit = false;
    }
};
struct StageInPool : HostPool {};
static HostPool& host_pool() { return per_device<StageInPool>(); }
struct StageOutPool : HostPool {};
static HostPool& out_pool() { return per_device<StageOutPool>(); }

// Copy-out of pageable results runs on its own thread: it waits for a group's device->host event and scatters the pinned
// staging buffer into the caller's arrays while the calling thread is already staging the next group in — without it the two
// host copies serialise and a pageable batch runs at half the speed of a page-locked one (tools/ntt_e2e_ab.py).
struct Drainer {
    using Copy = HostCopy;
    struct Job { cudaEvent_t ev; std::vector<Copy> copies; uint64_t id; };
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> q;
    uint64_t next_id = 1, done_id = 0;
    bool quit = false, failed = false;
    int device = 0, slot = 0;

    void start(int dev) {
        if (th.joinable()) return;
        device = dev;
        slot = cur_slot();
        th = std::thread([this] { loop(); });
    }
    void loop() {
        tl_slot = slot;  // the stage-out pool of this device
        cudaSetDevice(device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return quit || !q.empty(); });
                if (q.empty()) return;
                j = std::move(q.front());
                q.pop_front();
            }
            bool ok = cudaEventSynchronize(j.ev) == cudaSuccess;
            if (ok) out_pool().copy_many(j.copies.data(), j.copies.size());
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!ok) failed = true;
                done_id = j.id;
            }
            cv_done.notify_all();
        }
    }
    uint64_t submit(cudaEvent_t ev, std::vector<Copy>&& copies) {
        std::lock_guard<std::mutex> lk(mu);
        const uint64_t id = next_id++;
        q.push_back(Job{ev, std::move(copies), id});
        cv_job.notify_one();
        return id;
    }
    bool wait(uint64_t id) {  // false if any job failed
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return done_id >= id; });
        bool ok = !failed;
        return ok;
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_job.notify_all();
        th.join();
        quit = false;
        failed = false;
    }
    ~Drainer() { stop(); }
};
static Drainer& drainer() { return per_device<Drainer>(); }

constexpr int PIPE_SLOTS = 3;
struct PipeSlot {
    DevBuf x, y, in;
    PinBuf h_in, h_out;
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    bool busy = false;
    // deferred copy-out of a pageable group
    size_t c0 = 0, nc = 0;
    uint64_t drain_ticket = 0;  // non-zero: the drainer thread copies this group out
};
struct Pipeline {
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    PipeSlot slot[PIPE_SLOTS];
    bool ready = false;
    size_t group_bytes_override = 0;  // 0 = automatic
    int depth = PIPE_SLOTS;           // 1 = serial (for A/B measurements)
};
static Pipeline& pipeline() { return per_device<Pipeline>(); }
static int pipeline_init() {
    Pipeline& pl = pipeline();
    if (pl.ready) return ZKB_OK;
    ZKB_CUDA_TRY(cudaStreamCreateWithFlags(&pl.s_h2d, cudaStreamNonBlocking));
    ZKB_CUDA_TRY(cudaStreamCreateWithFlags(&pl.s_d2h, cudaStreamNonBlocking));
    for (auto& s : pl.slot) {
        ZKB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        ZKB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        ZKB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
    pl.ready = true;
    return ZKB_OK;
}
static void pipeline_release() {
    Pipeline& pl = pipeline();
    if (!pl.ready) return;
    for (auto& s : pl.slot) {
        s.x.release(); s.y.release(); s.in.release(); s.h_in.release(); s.h_out.release();
        cudaEventDestroy(s.ev_h2d); cudaEventDestroy(s.ev_comp); cudaEventDestroy(s.ev_d2h);
        s.busy = false;
    }
    cudaStreamDestroy(pl.s_h2d);
    cudaStreamDestroy(pl.s_d2h);
    pl.ready = false;
    drainer().stop();
    host_pool().stop();
    out_pool().stop();
}

// true when the driver can DMA straight from/to this host pointer (cudaMallocHost / cudaHostRegister memory)
static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// host-buffer flavour: columns are separate host arrays
static int domain_op_host_one(DomainOp op, const uint64_t* const* in, uint64_t* const* out, size_t ncols, uint32_t k, uint32_t ek,
                              const uint64_t* omega_user) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(in, "input column array"));
    ZKB_TRY(check_ptr(out, "output column array"));
    uint32_t log_n = (op == OP_C2E || op == OP_E2C) ? ek : k;
    if (log_n < 1 || log_n > 28 || k > log_n) { set_error("bad domain sizes k=%u extended_k=%u", k, ek); return ZKB_ERR_ARG; }
    for (size_t cidx = 0; cidx < ncols; ++cidx) {
        ZKB_TRY(check_ptr(in[cidx], "input column"));
        ZKB_TRY(check_ptr(out[cidx], "output column"));
    }
    if (ncols == 0) return ZKB_OK;
    ZKB_TRY(pipeline_init());
    Ctx& c = ctx();
    Pipeline& pl = pipeline();
    const uint64_t N = 1ull << log_n, in_len = op == OP_C2E ? (1ull << k) : N;
    const size_t col_out = (size_t)N * 32, col_in = (size_t)in_len * 32;
    // group size: about a quarter of the batch, between 8 MiB and 256 MiB of output per group
    size_t target = pl.group_bytes_override ? pl.group_bytes_override : (ncols * col_out) / 4;
    if (!pl.group_bytes_override) {
        if (target < ((size_t)8 << 20)) target = (size_t)8 << 20;
        if (target > ((size_t)256 << 20)) target = (size_t)256 << 20;
    }
    size_t per = target / col_out;
    if (per < 1) per = 1;
    if (per > 4096) per = 4096;
    const int depth = pl.depth < 1 ? 1 : (pl.depth > PIPE_SLOTS ? PIPE_SLOTS : pl.depth);
    std::vector<uint8_t> pinned_in(ncols), pinned_out(ncols);
    for (size_t i = 0; i < ncols; ++i) {
        pinned_in[i] = host_ptr_is_pinned(in[i]);
        pinned_out[i] = (const void*)in[i] == (const void*)out[i] ? pinned_in[i] : host_ptr_is_pinned(out[i]);
    }

    // wait for a slot's device->host copies and hand pageable columns back to the caller
    auto finish = [&](PipeSlot& s) -> int {
        if (!s.busy) return ZKB_OK;
        if (s.drain_ticket) {
            const bool ok = drainer().wait(s.drain_ticket);
            s.drain_ticket = 0;
            if (!ok) { set_error("device->host copy failed in the drainer thread"); return ZKB_ERR_CUDA; }
        } else {
            ZKB_CUDA_TRY(cudaEventSynchronize(s.ev_d2h));
        }
        s.busy = false;
        return ZKB_OK;
    };
    auto fail = [&](int rc) {
        cudaStreamSynchronize(pl.s_h2d); cudaStreamSynchronize(c.stream); cudaStreamSynchronize(pl.s_d2h);
        for (auto& s : pl.slot) {
            if (s.drain_ticket) drainer().wait(s.drain_ticket);
            s.drain_ticket = 0;
            s.busy = false;
        }
        return rc;
    };

    static const bool trace = getenv("ZKB_TRACE") != nullptr;
    size_t g = 0;
    for (size_t c0 = 0; c0 < ncols; c0 += per, ++g) {
        const size_t nc = ncols - c0 < per ? ncols - c0 : per;
        PipeSlot& s = pl.slot[g % depth];
        if (trace) fprintf(stderr, "[zkb] pipeline group %zu: columns [%zu, %zu) slot %zu\n", g, c0, c0 + nc, g % depth);
        int rc = finish(s);
        if (rc == ZKB_OK) rc = s.x.reserve(nc * col_out);
        if (rc == ZKB_OK) rc = s.y.reserve(nc * col_out);
        if (rc == ZKB_OK && op == OP_C2E) rc = s.in.reserve(nc * col_in);
        bool any_pg_in = false, any_pg_out = false;
        for (size_t i = 0; i < nc; ++i) { any_pg_in |= !pinned_in[c0 + i]; any_pg_out |= !pinned_out[c0 + i]; }
        if (rc == ZKB_OK && any_pg_in) rc = s.h_in.reserve(nc * col_in);
        if (rc == ZKB_OK && any_pg_out) rc = s.h_out.reserve(nc * col_out);
        if (rc != ZKB_OK) return fail(rc);
        char* d_in = op == OP_C2E ? (char*)s.in.p : (char*)s.x.p;
        cudaError_t e = cudaSuccess;
        bool all_pg_in = true, all_pg_out = true;
        for (size_t i = 0; i < nc; ++i) { all_pg_in &= !pinned_in[c0 + i]; all_pg_out &= !pinned_out[c0 + i]; }
        if (any_pg_in) {   // every pageable column of the group in one parallel copy
            std::vector<HostCopy> stage;
            for (size_t i = 0; i < nc; ++i)
                if (!pinned_in[c0 + i]) stage.push_back({(char*)s.h_in.p + i * col_in, in[c0 + i], col_in});
            host_pool().copy_many(stage.data(), stage.size());
        }
        if (all_pg_in) {   // the staged group is contiguous on both sides: one DMA
            e = cudaMemcpyAsync(d_in, s.h_in.p, nc * col_in, cudaMemcpyHostToDevice, pl.s_h2d);
            count_h2d(nc * col_in);
        } else {
            for (size_t i = 0; i < nc && e == cudaSuccess; ++i) {
                const void* src = pinned_in[c0 + i] ? (const void*)in[c0 + i] : (const void*)((char*)s.h_in.p + i * col_in);
                e = cudaMemcpyAsync(d_in + i * col_in, src, col_in, cudaMemcpyHostToDevice, pl.s_h2d);
                count_h2d(col_in);
            }
        }
        if (e == cudaSuccess) e = cudaEventRecord(s.ev_h2d, pl.s_h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, s.ev_h2d, 0);
        if (e != cudaSuccess) { set_error("host->device staging failed: %s", cudaGetErrorString(e)); return fail(ZKB_ERR_CUDA); }
        rc = domain_op_dev(op, reinterpret_cast<const uint4*>(d_in), s.x.as<uint4>(), s.y.as<uint4>(), nc, k, ek, omega_user, c.stream);
        if (rc != ZKB_OK) return fail(rc);
        e = cudaEventRecord(s.ev_comp, c.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(pl.s_d2h, s.ev_comp, 0);
        if (all_pg_out) {
            if (e == cudaSuccess) e = cudaMemcpyAsync(s.h_out.p, s.x.p, nc * col_out, cudaMemcpyDeviceToHost, pl.s_d2h);
        } else {
            for (size_t i = 0; i < nc && e == cudaSuccess; ++i) {
                void* dst = pinned_out[c0 + i] ? (void*)out[c0 + i] : (void*)((char*)s.h_out.p + i * col_out);
                e = cudaMemcpyAsync(dst, (char*)s.x.p + i * col_out, col_out, cudaMemcpyDeviceToHost, pl.s_d2h);
            }
        }
        if (e == cudaSuccess) e = cudaEventRecord(s.ev_d2h, pl.s_d2h);
        if (e != cudaSuccess) { set_error("device->host staging failed: %s", cudaGetErrorString(e)); return fail(ZKB_ERR_CUDA); }
        s.busy = true;
        s.c0 = c0;
        s.nc = nc;
        if (any_pg_out) {  // hand the scatter of the pageable columns to the drainer thread
            std::vector<Drainer::Copy> copies;
            for (size_t i = 0; i < nc; ++i)
                if (!pinned_out[c0 + i]) copies.push_back({(void*)out[c0 + i], (char*)s.h_out.p + i * col_out, col_out});
            drainer().start(c.device);
            s.drain_ticket = drainer().submit(s.ev_d2h, std::move(copies));
        }
    }
    // drain in submission order
    for (size_t i = 0; i < (size_t)depth; ++i) {
        PipeSlot& s = pl.slot[(g + i) % depth];
        int rc = finish(s);
        if (rc != ZKB_OK) return fail(rc);
    }
    return ZKB_OK;
}

int dist_inprocess_ntt_host(const uint64_t* in, uint64_t* out, const uint64_t omega[4], uint32_t log_n, const Fr* out_scale);  // dist.cu
bool dist_inprocess_supported(uint32_t log_n);

// Host-buffer NTT-family op over all bound devices: a batch is split by column (contiguous column ranges, one per device,
// each running its own three-stream pipeline over its own PCIe link — no communication); ONE large transform is sharded
// over the devices with the exchange fused into the NTT passes over NVLink peer memory (dist.cu).
static int domain_op_host(DomainOp op, const uint64_t* const* in, uint64_t* const* out, size_t ncols, uint32_t k, uint32_t ek,
                          const uint64_t* omega_user) {
    const int nd = cur_slot() == 0 ? device_slots() : 1;
    if (nd <= 1 || ncols == 0 || !in || !out) return domain_op_host_one(op, in, out, ncols, k, ek, omega_user);
    const uint32_t log_n = (op == OP_C2E || op == OP_E2C) ? ek : k;
    const int min_cols_log = multi_threshold(1), dist_min_log = multi_threshold(2);
    if (ncols == 1) {
        if (op == OP_C2E || (int)log_n < dist_min_log || !dist_inprocess_supported(log_n) || !in[0] || !out[0])
            return domain_op_host_one(op, in, out, ncols, k, ek, omega_user);
        std::lock_guard<std::recursive_mutex> lock(ctx().mu);
        ZKB_TRY(require_init());
        uint64_t om[4];
        Fr scale[3];
        const Fr* out_scale = nullptr;
        switch (op) {
            case OP_FFT: ZKB_TRY(check_ptr(omega_user, "omega")); memcpy(om, omega_user, 32); break;
            case OP_C2L: fr_to_limbs64(fr_omega_host(k), om); break;
            case OP_L2C:
                fr_to_limbs64(fr_omega_inv_host(k), om);
                scale[0] = scale[1] = scale[2] = fr_pow2_inv_host(k);
                out_scale = scale;
                break;
            default: {  // OP_E2C
                fr_to_limbs64(fr_omega_inv_host(ek), om);
                Fr ninv = fr_pow2_inv_host(ek), z = fr_zeta(), z2 = fp_sqr(z);
                scale[0] = ninv; scale[1] = fp_mul(ninv, z2); scale[2] = fp_mul(ninv, z);
                out_scale = scale;
            }
        }
        return dist_inprocess_ntt_host(in[0], out[0], om, log_n, out_scale);
    }
    if ((int)log_n < min_cols_log && ncols < 64) return domain_op_host_one(op, in, out, ncols, k, ek, omega_user);
    const int g = ncols < (size_t)nd ? (int)ncols : nd;
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);  // lock order: home slot first, then the fan-out
    return run_on_devices(g, [&](int slot) -> int {
        size_t c0, nc;
        point_range(ncols, slot, g, &c0, &nc);
        return domain_op_host_one(op, in + c0, out + c0, nc, k, ek, omega_user);
    });
}

// MSM of device scalars against srs[offset .. offset+n), through the window table when available
static int msm_srs_dev(Srs* srs, size_t offset, const uint4* d_scalars, size_t n, cudaStream_t stream, uint64_t* out,
                       uint32_t ncols = 1, uint32_t phase = MSM_WHOLE, cudaEvent_t input_ready = nullptr) {
    if (phase & MSM_LAST) srs->commits += ncols;
    if (srs_table_ready(srs, stream)) {
        MsmTable t{srs->table.as<uint4>() + 4 * offset, srs->n, srs->table_c, srs->table_nwin};
        return msm_run(d_scalars, nullptr, n, stream, out, &t, ncols, phase, input_ready);
    }
    if (phase != MSM_WHOLE) { set_error("sliced MSM needs the SRS window table"); return ZKB_ERR_ARG; }
    return msm_run(d_scalars, srs->bases.as<uint4>() + 4 * offset, n, stream, out, nullptr, ncols);
}

static int upload_scalars(const uint64_t* scalars, size_t n) {
    Ctx& c = ctx();
    HostIo& h = hostio();
    ZKB_TRY(check_ptr(scalars, "scalars"));
    ZKB_TRY(h.scalars.reserve(n * 32));
    ZKB_CUDA_TRY(cudaMemcpyAsync(h.scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, c.stream));
    count_h2d(n * 32);
    return ZKB_OK;
}

// Host scalars -> MSM against srs[offset .. offset+n).  Large commits are cut into point-range slices whose upload
// (through pinned staging when the caller's memory is pageable) overlaps the previous slice's digit / sort / accumulate
// kernels; the slices add into one bucket array and the last one reduces (msm_run `phase`).
static int g_msm_slices = 0;  // 0 = automatic
static int msm_srs_host(Srs* srs, size_t offset, const uint64_t* scalars, size_t n, uint64_t* out_jac) {
    Ctx& c = ctx();
    HostIo& h = hostio();
    ZKB_TRY(check_ptr(scalars, "scalars"));
    // Slice boundaries (profiles/r1_tuning.txt §3).  Explicit count: equal slices.  Automatic: a SHORT first slice (its upload
    // is the only exposed one) and growing later ones, so most of the points still sit in long bucket runs; pageable memory is
    // staged by host threads at a fraction of PCIe speed, so it gets one more, even shorter, leading slice.
    const bool pinned = host_ptr_is_pinned(scalars);
    std::vector<size_t> bounds;  // exclusive upper ends
    if (g_msm_slices > 1) {
        const size_t per = (n + g_msm_slices - 1) / g_msm_slices;
        for (size_t hi = per; hi < n; hi += per) bounds.push_back(hi);
    } else if (g_msm_slices == 0 && n >= ((size_t)1 << 21)) {
        const bool big = n >= ((size_t)1 << 23);
        if (pinned) {
            if (big) { bounds.push_back(n / 8); bounds.push_back(n / 2); }
            else bounds.push_back(n / 4);
        } else {
            if (big) { bounds.push_back(n / 16); bounds.push_back(n / 4); bounds.push_back(n / 8 * 5); }
            else { bounds.push_back(n / 8); bounds.push_back(n / 2); }
        }
    }
    bounds.push_back(n);
    if (bounds.size() > 1 && !srs_table_ready(srs, c.stream)) { bounds.clear(); bounds.push_back(n); }
    const size_t slices = bounds.size();
    if (slices == 1) {
        ZKB_TRY(upload_scalars(scalars, n));
        return msm_srs_dev(srs, offset, h.scalars.as<uint4>(), n, c.stream, out_jac);
    }
    ZKB_TRY(pipeline_init());
    Pipeline& pl = pipeline();
    ZKB_TRY(h.scalars.reserve(n * 32));
    size_t per = 0;  // largest slice
    for (size_t i = 0, lo = 0; i < slices; lo = bounds[i++]) per = bounds[i] - lo > per ? bounds[i] - lo : per;
    if (!pinned) {
        for (auto& b : h.stage) ZKB_TRY(b.reserve(per * 32));
        for (auto& e : h.stage_ev)
            if (!e) ZKB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    auto fail = [&](int rc) {
        cudaStreamSynchronize(pl.s_h2d); cudaStreamSynchronize(c.stream);
        return rc;
    };
    for (size_t sidx = 0, lo = 0; sidx < slices; lo = bounds[sidx++]) {
        const size_t cnt = bounds[sidx] - lo;
        const void* src = scalars + 4 * lo;
        PipeSlot& ev = pl.slot[sidx % PIPE_SLOTS];  // only the events of the slot are used here
        if (!pinned) {
            PinBuf& st = h.stage[sidx & 1];
            if (sidx >= 2) {  // the upload that last read this staging buffer must be done
                cudaError_t e = cudaEventSynchronize(h.stage_ev[sidx & 1]);
                if (e != cudaSuccess) { set_error("staging wait failed: %s", cudaGetErrorString(e)); return fail(ZKB_ERR_CUDA); }
            }
            host_pool().copy(st.p, src, cnt * 32);
            src = st.p;
        }
        cudaError_t e = cudaMemcpyAsync((char*)h.scalars.p + lo * 32, src, cnt * 32, cudaMemcpyHostToDevice, pl.s_h2d);
        count_h2d(cnt * 32);
        if (e == cudaSuccess && !pinned) e = cudaEventRecord(h.stage_ev[sidx & 1], pl.s_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(ev.ev_h2d, pl.s_h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, ev.ev_h2d, 0);
        if (e != cudaSuccess) { set_error("scalar upload failed: %s", cudaGetErrorString(e)); return fail(ZKB_ERR_CUDA); }
        const uint32_t phase = (lo == 0 ? MSM_FIRST : 0u) | (lo + cnt == n ? MSM_LAST : 0u);
        int rc = msm_srs_dev(srs, offset + lo, h.scalars.as<uint4>() + 2 * lo, cnt, c.stream, out_jac, 1, phase, pinned ? ev.ev_h2d : nullptr);
        if (rc != ZKB_OK) return fail(rc);
    }
    return ZKB_OK;
}

static int find_srs_set(uint64_t handle, SrsSet** out) {
    std::lock_guard<std::mutex> lk(g_srs_mu);
    auto it = srs_map().find(handle);
    if (it == srs_map().end()) { set_error("unknown SRS handle %llu", (unsigned long long)handle); return ZKB_ERR_HANDLE; }
    *out = it->second;
    return ZKB_OK;
}
// the replica of a registered SRS on the calling thread's device slot
static int find_srs(uint64_t handle, Srs** out) {
    SrsSet* set;
    ZKB_TRY(find_srs_set(handle, &set));
    Srs* s = set->dev[cur_slot()];
    if (!s) { set_error("SRS handle %llu has no replica on device slot %d", (unsigned long long)handle, cur_slot()); return ZKB_ERR_HANDLE; }
    *out = s;
    return ZKB_OK;
}

// hooks for poly.cu (polynomials resident under handles)
int srs_msm_dev_by_handle(uint64_t srs_handle, const uint4* d_scalars, size_t n, cudaStream_t s, uint64_t* out) {
    Srs* srs;
    ZKB_TRY(find_srs(srs_handle, &srs));
    if (n > srs->n) { set_error("MSM length %zu exceeds the registered SRS length %zu", n, srs->n); return ZKB_ERR_ARG; }
    if (n == 0) { msm_identity_out(out); return ZKB_OK; }
    return msm_srs_dev(srs, 0, d_scalars, n, s, out);
}
int domain_dev_by_op(int op, const uint4* d_in, uint4* d_a, uint4* d_b, size_t ncols, uint32_t k, uint32_t ek, cudaStream_t s) {
    static const DomainOp ops[5] = {OP_FFT, OP_L2C, OP_C2L, OP_C2E, OP_E2C};
    if (op < 1 || op > 4) { set_error("bad domain op"); return ZKB_ERR_ARG; }
    return domain_op_dev(ops[op], d_in, d_a, d_b, ncols, k, ek, nullptr, s);
}
void poly_release_all();

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* zkb_version(void) { return "zkb200 0.1 (sm_100a)"; }
const char* zkb_last_error(void) { return g_error.c_str(); }

static int init_slot(int slot, int dev) {
    SlotScope sc(slot);
    Ctx& c = ctx();
    ZKB_CUDA_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    ZKB_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor); return ZKB_ERR_NO_DEVICE; }
    ZKB_CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = dev;
    c.sm_count = prop.multiProcessorCount;
    c.inited = true;
    return ZKB_OK;
}

int zkb_init(const int* devices, int ndev) {
    Ctx& c = per_device<Ctx>();  // the caller's slot (0 for application threads)
    std::lock_guard<std::recursive_mutex> lock(c.mu);
    if (ndev < 0 || ndev > ZKB_MAX_DEVICES) { set_error("ndev must be in [0, %d] (got %d)", ZKB_MAX_DEVICES, ndev); return ZKB_ERR_ARG; }
    if (cur_slot() != 0) { set_error("zkb_init must be called from an application thread"); return ZKB_ERR_ARG; }
    int count = zkb_device_count();
    if (count == 0) { set_error("no CUDA device visible; libzkb200 has no CPU fallback"); return ZKB_ERR_NO_DEVICE; }
    int list[ZKB_MAX_DEVICES];
    int n = 0;
    if (devices && ndev >= 1) {
        for (int i = 0; i < ndev; ++i) list[n++] = devices[i];
    } else if (const char* e = getenv("ZKB_DEVICES")) {  // "all" or "0,1,2,3": lets an unmodified caller (zkb_init(NULL, 0)) drive the whole box
        if (!strcmp(e, "all")) { for (int i = 0; i < count && i < ZKB_MAX_DEVICES; ++i) list[n++] = i; }
        else for (const char* q = e; *q && n < ZKB_MAX_DEVICES;) {
            list[n++] = atoi(q);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    }
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
        list[n++] = dev;
    }
    for (int i = 0; i < n; ++i) {
        if (list[i] < 0 || list[i] >= count) { set_error("device %d out of range (count %d)", list[i], count); return ZKB_ERR_ARG; }
        for (int j = 0; j < i; ++j)
            if (list[j] == list[i]) { set_error("device %d listed twice", list[i]); return ZKB_ERR_ARG; }
    }
    if (c.inited) {
        if (!devices) return ZKB_OK;
        bool same = n == g_nslots;
        for (int i = 0; same && i < n; ++i) same = list[i] == g_slot_device[i];
        if (!same) { set_error("already bound to %d device(s) starting with device %d; call zkb_shutdown first", g_nslots, g_slot_device[0]); return ZKB_ERR_ARG; }
        return ZKB_OK;
    }
    for (int i = 0; i < n; ++i) g_slot_device[i] = list[i];
    for (int i = 0; i < n; ++i) {
        int rc = init_slot(i, list[i]);
        if (rc != ZKB_OK) {
            for (int j = 0; j < i; ++j) { SlotScope sc(j); Ctx& cj = ctx(); cudaStreamDestroy(cj.stream); cj.stream = nullptr; cj.inited = false; }
            return rc;
        }
    }
    // all pairs reach each other's HBM over NVLink (SRS replication, the sharded NTT's fused exchange)
    for (int i = 0; i < n && n > 1; ++i) {
        SlotScope sc(i);
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(list[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                set_error("device %d cannot map device %d's memory (%s); multi-device mode needs peer access", list[i], list[j], cudaGetErrorString(e));
                for (int q = 0; q < n; ++q) { SlotScope s2(q); Ctx& cq = ctx(); cudaStreamDestroy(cq.stream); cq.stream = nullptr; cq.inited = false; }
                return ZKB_ERR_CUDA;
            }
            cudaGetLastError();
        }
    }
    g_nslots = n;
    for (int i = 1; i < n; ++i) g_workers[i].start(i);
    cudaSetDevice(list[0]);
    return ZKB_OK;
}

// releases everything the calling slot holds on its device
static void shutdown_slot() {
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> lock(c.mu);
    if (!c.inited) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    dist_shutdown();
    setup_release();
    ntt_clear_plans();
    msm_release_workspace();
    HostIo& h = hostio();
    h.scalars.release(); h.bases.release(); h.x.release(); h.y.release(); h.in.release();
    for (auto& b : h.stage) b.release();
    for (auto& e : h.stage_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    pipeline_release();
    for (auto& kv : c.timers)
        for (auto& pr : kv.second.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    c.timers.clear();
    for (auto e : c.event_pool) cudaEventDestroy(e);
    c.event_pool.clear();
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.inited = false;
}

void zkb_shutdown(void) {
    if (cur_slot() != 0 || g_nslots == 0) return;
    {
        std::lock_guard<std::recursive_mutex> lock(ctx().mu);
        if (!ctx().inited) return;
        cudaSetDevice(ctx().device);
        cudaDeviceSynchronize();
        poly_release_all();
        std::lock_guard<std::mutex> lk(g_srs_mu);
        for (auto& kv : srs_map()) {
            for (int i = 0; i < ZKB_MAX_DEVICES; ++i) {
                Srs* r = kv.second->dev[i];
                if (!r) continue;
                SlotScope sc(i);
                cudaDeviceSynchronize();
                r->bases.release(); r->table.release();
                delete r;
            }
            delete kv.second;
        }
        srs_map().clear();
    }
    const int n = g_nslots;
    for (int i = n - 1; i >= 1; --i) {
        g_workers[i].stop();
        SlotScope sc(i);
        shutdown_slot();
    }
    shutdown_slot();
    g_nslots = 0;
}

// ---- MSM ---------------------------------------------------------------------------------------------------------------------
int zkb_msm_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_jac[12]) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(out_jac, "out_jac"));
    if (n == 0) { msm_identity_out(out_jac); return ZKB_OK; }
    ZKB_TRY(check_ptr(bases, "bases"));
    ZKB_TRY(check_ptr(scalars, "scalars"));
    auto one = [&](size_t off, size_t len, uint64_t* out) -> int {
        std::lock_guard<std::recursive_mutex> lk(ctx().mu);
        ZKB_TRY(require_init());
        HostIo& h = hostio();
        ZKB_TRY(h.bases.reserve(len * 64));
        ZKB_CUDA_TRY(cudaMemcpyAsync(h.bases.p, bases + 8 * off, len * 64, cudaMemcpyHostToDevice, ctx().stream));
        count_h2d(len * 64);
        ZKB_TRY(upload_scalars(scalars + 4 * off, len));
        return msm_run(h.scalars.as<uint4>(), h.bases.as<uint4>(), len, ctx().stream, out);
    };
    const int g = msm_fanout(n);
    if (g <= 1) return one(0, n, out_jac);
    uint64_t parts[ZKB_MAX_DEVICES][12];
    ZKB_TRY(run_on_devices(g, [&](int slot) -> int {
        size_t off, len;
        point_range(n, slot, g, &off, &len);
        return one(off, len, parts[slot]);
    }));
    return g1_sum_host(&parts[0][0], (size_t)g, out_jac);
}

int zkb_srs_register(const uint64_t* bases, size_t n, uint64_t* handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(handle, "handle"));
    if (n) ZKB_TRY(check_ptr(bases, "bases"));
    Srs* s = new Srs();
    s->n = n;
    int rc = s->bases.reserve(n ? n * 64 : 64);
    if (rc != ZKB_OK) { delete s; return rc; }
    if (n) {
        cudaError_t e = cudaMemcpy(s->bases.p, bases, n * 64, cudaMemcpyHostToDevice);
        count_h2d(n * 64);
        if (e != cudaSuccess) { set_error("SRS upload failed: %s", cudaGetErrorString(e)); s->bases.release(); delete s; return ZKB_ERR_CUDA; }
    }
    return srs_publish(s, handle);
}

// n raw G1Affine (64 B each, Montgomery limbs — what SerdeFormat::RawBytes / RawBytesUnchecked write) from `path` at byte
// `offset` straight into HBM: read(2) into one pinned buffer while the other one is in flight to the device.
int zkb_srs_load_file(const char* path, uint64_t offset, size_t n, int check_points, uint64_t* handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(handle, "handle"));
    ZKB_TRY(check_ptr(path, "path"));
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return ZKB_ERR_ARG; }
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    if (fseeko(f, 0, SEEK_END) != 0) { set_error("cannot seek in %s", path); return ZKB_ERR_ARG; }
    const uint64_t fsize = (uint64_t)ftello(f);
    if (offset > fsize || (uint64_t)n * 64 > fsize - offset) {
        set_error("%s holds %llu bytes, %zu points at offset %llu need %llu", path, (unsigned long long)fsize, n,
                  (unsigned long long)offset, (unsigned long long)(offset + (uint64_t)n * 64));
        return ZKB_ERR_ARG;
    }
    if (fseeko(f, (off_t)offset, SEEK_SET) != 0) { set_error("cannot seek in %s", path); return ZKB_ERR_ARG; }
    Srs* s = new Srs();
    s->n = n;
    int rc = s->bases.reserve(n ? n * 64 : 64);
    if (rc != ZKB_OK) { delete s; return rc; }
    const size_t chunk = (size_t)32 << 20;
    void* pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = ctx().stream;
    auto cleanup = [&](int code) {
        cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i) {
            if (pin[i]) cudaFreeHost(pin[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (code != ZKB_OK) { s->bases.release(); delete s; }
        return code;
    };
    for (int i = 0; i < 2; ++i) {
        if (cudaMallocHost(&pin[i], chunk) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            set_error("pinned staging allocation failed");
            return cleanup(ZKB_ERR_OOM);
        }
    }
    const size_t total = n * 64;
    for (size_t done = 0, i = 0; done < total; ++i) {
        const size_t len = total - done < chunk ? total - done : chunk;
        if (i >= 2 && cudaEventSynchronize(ev[i & 1]) != cudaSuccess) { set_error("SRS upload failed"); return cleanup(ZKB_ERR_CUDA); }
        if (fread(pin[i & 1], 1, len, f) != len) { set_error("short read from %s", path); return cleanup(ZKB_ERR_ARG); }
        cudaError_t e = cudaMemcpyAsync((char*)s->bases.p + done, pin[i & 1], len, cudaMemcpyHostToDevice, st);
        count_h2d(len);
        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
        if (e != cudaSuccess) { set_error("SRS upload failed: %s", cudaGetErrorString(e)); return cleanup(ZKB_ERR_CUDA); }
        done += len;
    }
    if (check_points) {
        uint64_t bad = 0;
        rc = g1_check_on_curve_dev(s->bases.as<uint4>(), n, &bad, st);
        if (rc != ZKB_OK) return cleanup(rc);
        if (bad) { set_error("%s: %llu of %zu points are not on the curve", path, (unsigned long long)bad, n); return cleanup(ZKB_ERR_ARG); }
    }
    cleanup(ZKB_OK);
    return srs_publish(s, handle);
}

int zkb_srs_release(uint64_t handle) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    SrsSet* set;
    ZKB_TRY(find_srs_set(handle, &set));
    {
        std::lock_guard<std::mutex> lk(g_srs_mu);
        srs_map().erase(handle);
    }
    for (int i = 0; i < ZKB_MAX_DEVICES; ++i) {
        Srs* s = set->dev[i];
        if (!s) continue;
        SlotScope sc(i);
        cudaDeviceSynchronize();
        s->bases.release();
        s->table.release();
        delete s;
    }
    delete set;
    return ZKB_OK;
}

int zkb_msm_g1_srs_range(uint64_t handle, size_t offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(out_jac, "out_jac"));
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    if (offset > s->n || n > s->n - offset) {
        set_error("MSM range [%zu, %zu) exceeds the registered SRS length %zu", offset, offset + n, s->n);
        return ZKB_ERR_ARG;
    }
    if (n == 0) { msm_identity_out(out_jac); return ZKB_OK; }
    ZKB_TRY(check_ptr(scalars, "scalars"));
    // several devices: the commit is sharded by SRS point range, every device uploads only its share of the scalars over its own
    // PCIe link, and the 96-byte partial results are folded here on the host (north_star: "combined on the host")
    const int g = msm_fanout(n);
    if (g <= 1) return msm_srs_host(s, offset, scalars, n, out_jac);
    SrsSet* set;
    ZKB_TRY(find_srs_set(handle, &set));
    uint64_t parts[ZKB_MAX_DEVICES][12];
    ZKB_TRY(run_on_devices(g, [&](int slot) -> int {
        std::lock_guard<std::recursive_mutex> lk(ctx().mu);
        ZKB_TRY(require_init());
        Srs* r = set->dev[slot];
        if (!r) { set_error("SRS has no replica on device slot %d", slot); return ZKB_ERR_HANDLE; }
        size_t off, len;
        point_range(n, slot, g, &off, &len);
        return msm_srs_host(r, offset + off, scalars + 4 * off, len, parts[slot]);
    }));
    return g1_sum_host(&parts[0][0], (size_t)g, out_jac);
}

int zkb_msm_g1_srs(uint64_t handle, const uint64_t* scalars, size_t n, uint64_t out_jac[12]) {
    return zkb_msm_g1_srs_range(handle, 0, scalars, n, out_jac);
}

// one device's share of a batch: columns [0, ncols) of `scalars` against the replica `s`
static int msm_batch_one(Srs* s, const uint64_t* const* scalars, size_t ncols, size_t n, uint64_t* out_jac) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    // all columns of a group are digit-decomposed, sorted and accumulated in ONE pass (column folded into the bucket
    // key); groups bound the sort size to 2^28 (key, index) pairs, the bucket keys to 24 bits and the bucket arrays to a
    // quarter of the free HBM (in table mode every column owns 2^(c-1) buckets whatever its length)
    const size_t per_col = n * 16;  // generous bound on windows per scalar
    size_t group = ((size_t)1 << 28) / per_col;
    if (group < 1) group = 1;
    if (group > 4096) group = 4096;
    Ctx& c = ctx();
    HostIo& h = hostio();
    {
        const bool table = srs_table_ready(s, c.stream);
        const MsmGeometry g1 = table ? msm_geometry(n, s->table_c, c.msm_chunk_override, true, 1)
                                     : msm_geometry(n, c.msm_c_override, c.msm_chunk_override, false, 1);
        const size_t per_col_buckets = (size_t)g1.bucket_sets << (g1.c - 1);
        // bucket keys of a group stay within the 24 bits the library's own bucket sort covers (bucket_sort.cuh), so no commit
        // ever leaves it for the toolkit's radix sort; a single column never needs more (W x 2^(c-1) <= 13 x 2^19)
        size_t by_keys = ((size_t)1 << 24) / per_col_buckets;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
        size_t by_mem = (free_b / 4) / (per_col_buckets * 128);
        if (by_keys < 1) by_keys = 1;
        if (by_mem < 1) by_mem = 1;
        if (group > by_keys) group = by_keys;
        if (group > by_mem) group = by_mem;
    }
    for (size_t c0 = 0; c0 < ncols; c0 += group) {
        size_t nc = ncols - c0 < group ? ncols - c0 : group;
        ZKB_TRY(h.scalars.reserve(nc * n * 32));
        bool all_pageable = nc * n * 32 >= ((size_t)4 << 20);
        for (size_t i = 0; i < nc && all_pageable; ++i) all_pageable = !host_ptr_is_pinned(scalars[c0 + i]);
        if (all_pageable) {   // many short pageable columns (the voter circuit): one parallel staging copy, one DMA
            ZKB_TRY(h.stage[0].reserve(nc * n * 32));
            std::vector<HostCopy> stage(nc);
            for (size_t i = 0; i < nc; ++i) stage[i] = {(char*)h.stage[0].p + i * n * 32, scalars[c0 + i], n * 32};
            host_pool().copy_many(stage.data(), stage.size());
            ZKB_CUDA_TRY(cudaMemcpyAsync(h.scalars.p, h.stage[0].p, nc * n * 32, cudaMemcpyHostToDevice, c.stream));   // msm_run synchronises: the buffer is free again afterwards
        } else {
            for (size_t i = 0; i < nc; ++i)
                ZKB_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(h.scalars.p) + i * n * 32, scalars[c0 + i], n * 32,
                                             cudaMemcpyHostToDevice, c.stream));
        }
        count_h2d(nc * n * 32);
        ZKB_TRY(msm_srs_dev(s, 0, h.scalars.as<uint4>(), n, c.stream, out_jac + 12 * c0, (uint32_t)nc));
    }
    return ZKB_OK;
}

int zkb_msm_g1_srs_batch(uint64_t handle, const uint64_t* const* scalars, size_t ncols, size_t n, uint64_t* out_jac) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (ncols == 0) return ZKB_OK;
    ZKB_TRY(check_ptr(scalars, "scalars"));
    ZKB_TRY(check_ptr(out_jac, "out_jac"));
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    if (n > s->n) { set_error("MSM length %zu exceeds the registered SRS length %zu", n, s->n); return ZKB_ERR_ARG; }
    if (n == 0) { for (size_t i = 0; i < ncols; ++i) msm_identity_out(out_jac + 12 * i); return ZKB_OK; }
    for (size_t i = 0; i < ncols; ++i) ZKB_TRY(check_ptr(scalars[i], "scalars column"));
    // several devices: the batch is split by column (contiguous column ranges), no communication
    int g = cur_slot() == 0 ? device_slots() : 1;
    if ((size_t)g > ncols) g = (int)ncols;
    if (g <= 1 || ncols * n < ((size_t)1 << 17)) return msm_batch_one(s, scalars, ncols, n, out_jac);
    SrsSet* set;
    ZKB_TRY(find_srs_set(handle, &set));
    return run_on_devices(g, [&](int slot) -> int {
        Srs* r = set->dev[slot];
        if (!r) { set_error("SRS has no replica on device slot %d", slot); return ZKB_ERR_HANDLE; }
        size_t c0, nc;
        point_range(ncols, slot, g, &c0, &nc);
        return msm_batch_one(r, scalars + c0, nc, n, out_jac + 12 * c0);
    });
}

int zkb_msm_g1_srs_dev(uint64_t handle, size_t offset, const void* d_scalars, size_t n, uint64_t out_jac[12], void* stream) {
    SlotScope scope(slot_of_device_ptr(d_scalars));  // several devices in one process: the pointer's owner runs the call
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(out_jac, "out_jac"));
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    if (offset > s->n || n > s->n - offset) { set_error("MSM range exceeds the registered SRS length %zu", s->n); return ZKB_ERR_ARG; }
    if (n) ZKB_TRY(check_ptr(d_scalars, "d_scalars"));
    return msm_srs_dev(s, offset, reinterpret_cast<const uint4*>(d_scalars), n, (cudaStream_t)stream, out_jac);
}

int zkb_srs_set_precompute(int mode) {
    if (mode < 0 || mode > 2) { set_error("mode must be 0 (off), 1 (eager) or 2 (automatic)"); return ZKB_ERR_ARG; }
    g_precompute = mode;
    return ZKB_OK;
}

int zkb_srs_precompute(uint64_t handle, uint32_t* window_bits, uint64_t* table_bytes) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    SrsSet* set;
    ZKB_TRY(find_srs_set(handle, &set));
    if (cur_slot() == 0 && device_slots() > 1)   // every replica builds its own table, all devices at once
        run_on_devices(device_slots(), [&](int slot) -> int {
            std::lock_guard<std::recursive_mutex> lk(ctx().mu);
            if (require_init() == ZKB_OK && set->dev[slot]) srs_table_ready(set->dev[slot], ctx().stream, true);
            return ZKB_OK;
        });
    bool ok = srs_table_ready(s, ctx().stream, true);
    if (window_bits) *window_bits = ok ? s->table_c : 0;
    if (table_bytes) *table_bytes = ok ? (uint64_t)s->table_nwin * s->n * 64 : 0;
    return ZKB_OK;
}

int zkb_srs_table_info(uint64_t handle, uint32_t* window_bits, uint64_t* table_bytes, uint64_t* commits) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    if (window_bits) *window_bits = s->table_c;
    if (table_bytes) *table_bytes = s->table_c ? (uint64_t)s->table_nwin * s->n * 64 : 0;
    if (commits) *commits = s->commits;
    return ZKB_OK;
}

int zkb_g1_sum(const uint64_t* points_jac, size_t count, uint64_t out_jac[12]) {
    if (count) ZKB_TRY(check_ptr(points_jac, "points_jac"));
    ZKB_TRY(check_ptr(out_jac, "out_jac"));
    return g1_sum_host(points_jac, count, out_jac);
}

static int fixed_base_host(const uint64_t* scalars, size_t n, uint64_t* out_affine, bool windowed) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (n == 0) return ZKB_OK;
    ZKB_TRY(check_ptr(scalars, "scalars"));
    ZKB_TRY(check_ptr(out_affine, "out_affine"));
    HostIo& h = hostio();
    Ctx& c = ctx();
    ZKB_TRY(h.scalars.reserve(n * 32));
    ZKB_TRY(h.bases.reserve(n * 64));
    ZKB_CUDA_TRY(cudaMemcpyAsync(h.scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, c.stream));
    count_h2d(n * 32);
    if (windowed) ZKB_TRY(g1_fixed_base_window_dev(h.scalars.as<uint4>(), n, h.bases.as<uint4>(), c.stream));
    else ZKB_TRY(g1_fixed_base_mul_dev(h.scalars.as<uint4>(), n, h.bases.as<uint4>(), c.stream));
    ZKB_CUDA_TRY(cudaMemcpyAsync(out_affine, h.bases.p, n * 64, cudaMemcpyDeviceToHost, c.stream));
    ZKB_CUDA_TRY(cudaStreamSynchronize(c.stream));
    return ZKB_OK;
}
int zkb_g1_fixed_base_mul(const uint64_t* scalars, size_t n, uint64_t* out_affine) {
    return fixed_base_host(scalars, n, out_affine, true);
}
int zkb_g1_fixed_base_mul_naive(const uint64_t* scalars, size_t n, uint64_t* out_affine) {
    return fixed_base_host(scalars, n, out_affine, false);
}

int zkb_g1_batch_normalize(const uint64_t* points_jac, size_t n, uint64_t* out_affine) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    if (n == 0) return ZKB_OK;
    ZKB_TRY(check_ptr(points_jac, "points_jac"));
    ZKB_TRY(check_ptr(out_affine, "out_affine"));
    HostIo& h = hostio();
    Ctx& c = ctx();
    ZKB_TRY(h.x.reserve(n * 96));
    ZKB_TRY(h.bases.reserve(n * 64));
    ZKB_CUDA_TRY(cudaMemcpyAsync(h.x.p, points_jac, n * 96, cudaMemcpyHostToDevice, c.stream));
    count_h2d(n * 96);
    ZKB_TRY(g1_batch_to_affine_dev(h.x.as<uint4>(), n, h.bases.as<uint4>(), true, c.stream));
    ZKB_CUDA_TRY(cudaMemcpyAsync(out_affine, h.bases.p, n * 64, cudaMemcpyDeviceToHost, c.stream));
    ZKB_CUDA_TRY(cudaStreamSynchronize(c.stream));
    return ZKB_OK;
}

// ---- ParamsKZG::setup ----------------------------------------------------------------------------------------------------------
static int kzg_setup_common(uint32_t k, const uint64_t* s, uint64_t* g_out, uint64_t* gl_out, uint64_t* hg, uint64_t* hgl) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(s, "s"));
    if (k < 1 || k > 28) { set_error("k %u out of range [1, 28]", k); return ZKB_ERR_ARG; }
    const size_t n = (size_t)1 << k;
    Ctx& c = ctx();
    Srs* sg = nullptr;
    Srs* sgl = nullptr;
    auto make = [&](Srs** out) -> int {
        Srs* p = new Srs();
        p->n = n;
        int rc = p->bases.reserve(n * 64);
        if (rc != ZKB_OK) { delete p; return rc; }
        *out = p;
        return ZKB_OK;
    };
    const bool want_g = g_out || hg, want_gl = gl_out || hgl;
    int rc = ZKB_OK;
    if (want_g) rc = make(&sg);
    if (rc == ZKB_OK && want_gl) rc = make(&sgl);
    if (rc == ZKB_OK) rc = kzg_setup_dev(k, s, sg ? sg->bases.as<uint4>() : nullptr, sgl ? sgl->bases.as<uint4>() : nullptr, c.stream);
    if (rc == ZKB_OK && g_out && cudaMemcpyAsync(g_out, sg->bases.p, n * 64, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) rc = ZKB_ERR_CUDA;
    if (rc == ZKB_OK && gl_out && cudaMemcpyAsync(gl_out, sgl->bases.p, n * 64, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess) rc = ZKB_ERR_CUDA;
    if (rc == ZKB_OK && cudaStreamSynchronize(c.stream) != cudaSuccess) rc = ZKB_ERR_CUDA;
    if (rc == ZKB_ERR_CUDA) { set_error("ParamsKZG::setup failed: %s", cudaGetErrorString(cudaGetLastError())); }
    auto finish = [&](Srs* p, uint64_t* handle) {
        if (!p) return;
        if (rc == ZKB_OK && handle) rc = srs_publish(p, handle);
        else { p->bases.release(); delete p; }
    };
    finish(sg, hg);
    finish(sgl, hgl);
    return rc;
}
int zkb_kzg_setup(uint32_t k, const uint64_t s[4], uint64_t* g_out, uint64_t* g_lagrange_out) {
    if (!g_out && !g_lagrange_out) { set_error("both outputs are NULL"); return ZKB_ERR_ARG; }
    return kzg_setup_common(k, s, g_out, g_lagrange_out, nullptr, nullptr);
}
int zkb_kzg_setup_resident(uint32_t k, const uint64_t s[4], uint64_t* handle_g, uint64_t* handle_g_lagrange) {
    if (!handle_g && !handle_g_lagrange) { set_error("both handles are NULL"); return ZKB_ERR_ARG; }
    return kzg_setup_common(k, s, nullptr, nullptr, handle_g, handle_g_lagrange);
}
// best_fft::<Fr, G1>(a, omega, log_n): affine in, affine out (host buffers)
int zkb_g1_ntt(const uint64_t* points_affine, uint64_t* out_affine, const uint64_t omega[4], uint32_t log_n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(points_affine, "points_affine"));
    ZKB_TRY(check_ptr(out_affine, "out_affine"));
    ZKB_TRY(check_ptr(omega, "omega"));
    if (log_n > 26) { set_error("log_n %u out of range [0, 26]", log_n); return ZKB_ERR_ARG; }
    const size_t n = (size_t)1 << log_n;
    HostIo& h = hostio();
    Ctx& c = ctx();
    ZKB_TRY(h.bases.reserve(n * 64));
    ZKB_TRY(h.x.reserve(n * 64));
    ZKB_CUDA_TRY(cudaMemcpyAsync(h.bases.p, points_affine, n * 64, cudaMemcpyHostToDevice, c.stream));
    count_h2d(n * 64);
    ZKB_TRY(g1_fft_dev(h.bases.as<uint4>(), h.x.as<uint4>(), log_n, fr_from_limbs64(omega), nullptr, c.stream));
    ZKB_CUDA_TRY(cudaMemcpyAsync(out_affine, h.x.p, n * 64, cudaMemcpyDeviceToHost, c.stream));
    ZKB_CUDA_TRY(cudaStreamSynchronize(c.stream));
    return ZKB_OK;
}

// halo2_proofs::arithmetic::g_to_lagrange(g, k): g_lagrange = (1/n) * iFFT_G1(g) — derives the Lagrange-basis SRS from the
// monomial one when only g is known (an SRS file without g_lagrange); resident in, resident out.
int zkb_srs_g_to_lagrange(uint64_t handle_g, uint32_t k, uint64_t* handle_g_lagrange) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(handle_g_lagrange, "handle_g_lagrange"));
    Srs* g;
    ZKB_TRY(find_srs(handle_g, &g));
    if (k > 26 || g->n < ((size_t)1 << k)) { set_error("SRS holds %zu points, 2^%u needed", g->n, k); return ZKB_ERR_ARG; }
    const size_t n = (size_t)1 << k;
    Srs* gl = new Srs();
    gl->n = n;
    int rc = gl->bases.reserve(n * 64);
    const Fr omega_inv = fr_omega_inv_host(k), n_inv = fr_pow2_inv_host(k);
    if (rc == ZKB_OK) rc = g1_fft_dev(g->bases.as<uint4>(), gl->bases.as<uint4>(), k, omega_inv, &n_inv, ctx().stream);
    if (rc == ZKB_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) { set_error("g_to_lagrange failed"); rc = ZKB_ERR_CUDA; }
    if (rc != ZKB_OK) { cudaGetLastError(); gl->bases.release(); delete gl; return rc; }
    return srs_publish(gl, handle_g_lagrange);
}

int zkb_srs_download(uint64_t handle, uint64_t* bases_out, size_t n) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    Srs* s;
    ZKB_TRY(find_srs(handle, &s));
    if (n > s->n) { set_error("SRS holds %zu points, %zu requested", s->n, n); return ZKB_ERR_ARG; }
    if (n == 0) return ZKB_OK;
    ZKB_TRY(check_ptr(bases_out, "bases_out"));
    ZKB_CUDA_TRY(cudaMemcpy(bases_out, s->bases.p, n * 64, cudaMemcpyDeviceToHost));
    return ZKB_OK;
}

// ---- NTT ---------------------------------------------------------------------------------------------------------------------
int zkb_ntt_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n) {
    ZKB_TRY(check_ptr(omega, "omega"));
    const uint64_t* in[1] = {a};
    uint64_t* out[1] = {a};
    return domain_op_host(OP_FFT, in, out, 1, log_n, log_n, omega);
}
int zkb_ntt_fr_batch(uint64_t* const* cols, size_t ncols, const uint64_t omega[4], uint32_t log_n) {
    ZKB_TRY(check_ptr(omega, "omega"));
    return domain_op_host(OP_FFT, const_cast<const uint64_t* const*>(cols), cols, ncols, log_n, log_n, omega);
}
int zkb_lagrange_to_coeff(uint64_t* a, uint32_t k) {
    const uint64_t* in[1] = {a};
    uint64_t* out[1] = {a};
    return domain_op_host(OP_L2C, in, out, 1, k, k, nullptr);
}
int zkb_lagrange_to_coeff_batch(uint64_t* const* cols, size_t ncols, uint32_t k) {
    return domain_op_host(OP_L2C, const_cast<const uint64_t* const*>(cols), cols, ncols, k, k, nullptr);
}
int zkb_coeff_to_lagrange(uint64_t* a, uint32_t k) {
    const uint64_t* in[1] = {a};
    uint64_t* out[1] = {a};
    return domain_op_host(OP_C2L, in, out, 1, k, k, nullptr);
}
int zkb_coeff_to_extended(const uint64_t* in, uint64_t* out, uint32_t k, uint32_t extended_k) {
    const uint64_t* i1[1] = {in};
    uint64_t* o1[1] = {out};
    return domain_op_host(OP_C2E, i1, o1, 1, k, extended_k, nullptr);
}
int zkb_coeff_to_extended_batch(const uint64_t* const* in, uint64_t* const* out, size_t ncols, uint32_t k, uint32_t extended_k) {
    return domain_op_host(OP_C2E, in, out, ncols, k, extended_k, nullptr);
}
int zkb_extended_to_coeff(uint64_t* a, uint32_t k, uint32_t extended_k) {
    const uint64_t* in[1] = {a};
    uint64_t* out[1] = {a};
    return domain_op_host(OP_E2C, in, out, 1, k, extended_k, nullptr);
}
int zkb_fr_zeta(uint64_t out[4]) {
    ZKB_TRY(check_ptr(out, "out"));
    fr_to_limbs64(fr_zeta(), out);
    return ZKB_OK;
}
int zkb_thread_bind_device(int device) {
    if (g_nslots == 0) ZKB_TRY(require_init());
    for (int i = 0; i < g_nslots; ++i)
        if (g_slot_device[i] == device) {
            tl_slot = i;
            ZKB_CUDA_TRY(cudaSetDevice(device));
            return ZKB_OK;
        }
    set_error("device %d is not bound (zkb_init)", device);
    return ZKB_ERR_ARG;
}
int zkb_multi_device_set(int msm_min_share_log, int batch_ntt_min_log, int dist_ntt_min_log) {
    const int v[3] = {msm_min_share_log, batch_ntt_min_log, dist_ntt_min_log};
    for (int i = 0; i < 3; ++i) {
        if (v[i] < 0 || v[i] > 40) { set_error("thresholds are log2 sizes in [1, 40], 0 = default"); return ZKB_ERR_ARG; }
        g_multi[i] = v[i] ? v[i] : -1;
    }
    return ZKB_OK;
}
int zkb_bound_devices(int* devices, int capacity) {
    for (int i = 0; i < g_nslots && i < capacity && devices; ++i) devices[i] = g_slot_device[i];
    return g_nslots;
}
int zkb_fr_omega(uint32_t k, uint64_t out[4]) {
    ZKB_TRY(check_ptr(out, "out"));
    if (k > 28) { set_error("k %u exceeds the two-adicity 28", k); return ZKB_ERR_ARG; }
    fr_to_limbs64(fr_omega_host(k), out);
    return ZKB_OK;
}

static int dev_common(DomainOp op, const void* d_in, void* d_a, void* d_b, size_t ncols, uint32_t k, uint32_t ek,
                      const uint64_t* omega, void* stream) {
    SlotScope scope(slot_of_device_ptr(d_a));
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(d_in, "device input"));
    ZKB_TRY(check_ptr(d_a, "device data"));
    ZKB_TRY(check_ptr(d_b, "device scratch"));
    return domain_op_dev(op, reinterpret_cast<const uint4*>(d_in), reinterpret_cast<uint4*>(d_a), reinterpret_cast<uint4*>(d_b),
                         ncols, k, ek, omega, (cudaStream_t)stream);
}
int zkb_ntt_fr_dev(void* d_data, void* d_scratch, size_t ncols, const uint64_t omega[4], uint32_t log_n, void* stream) {
    ZKB_TRY(check_ptr(omega, "omega"));
    return dev_common(OP_FFT, d_data, d_data, d_scratch, ncols, log_n, log_n, omega, stream);
}
int zkb_coeff_to_extended_dev(const void* d_in, void* d_out, void* d_scratch, size_t ncols, uint32_t k, uint32_t extended_k, void* stream) {
    return dev_common(OP_C2E, d_in, d_out, d_scratch, ncols, k, extended_k, nullptr, stream);
}
int zkb_extended_to_coeff_dev(void* d_data, void* d_scratch, size_t ncols, uint32_t k, uint32_t extended_k, void* stream) {
    return dev_common(OP_E2C, d_data, d_data, d_scratch, ncols, k, extended_k, nullptr, stream);
}
int zkb_lagrange_to_coeff_dev(void* d_data, void* d_scratch, size_t ncols, uint32_t k, void* stream) {
    return dev_common(OP_L2C, d_data, d_data, d_scratch, ncols, k, k, nullptr, stream);
}

// ---- host memory / scheduler controls ---------------------------------------------------------------------------------------
int zkb_host_register(void* ptr, size_t bytes) {
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(ptr, "ptr"));
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return ZKB_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaHostRegister(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); return ZKB_ERR_CUDA; }
    return ZKB_OK;
}
int zkb_host_unregister(void* ptr) {
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(ptr, "ptr"));
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaHostUnregister failed: %s", cudaGetErrorString(e)); return ZKB_ERR_CUDA; }
    return ZKB_OK;
}
int zkb_msm_set_slices(int slices) {
    if (slices < 0 || slices > 64) { set_error("slices must be in [0, 64] (0 = automatic)"); return ZKB_ERR_ARG; }
    g_msm_slices = slices;
    return ZKB_OK;
}
int zkb_pipeline_set(int depth, size_t group_bytes) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    if (depth < 0 || depth > PIPE_SLOTS) { set_error("pipeline depth must be in [0, %d] (0 = default)", PIPE_SLOTS); return ZKB_ERR_ARG; }
    pipeline().depth = depth ? depth : PIPE_SLOTS;
    pipeline().group_bytes_override = group_bytes;
    return ZKB_OK;
}

// ---- tuning / measurement ----------------------------------------------------------------------------------------------------
int zkb_msm_set_params(uint32_t window_bits, uint32_t chunk) {
    if (window_bits && (window_bits < 2 || window_bits > 22)) { set_error("window_bits must be 0 or in [2, 22]"); return ZKB_ERR_ARG; }
    ctx().msm_c_override = window_bits;
    ctx().msm_chunk_override = chunk;
    return ZKB_OK;
}
int zkb_msm_get_params(size_t n, uint32_t* window_bits, uint32_t* num_windows, uint32_t* chunk) {
    MsmGeometry g = msm_geometry(n, ctx().msm_c_override, ctx().msm_chunk_override);
    if (window_bits) *window_bits = g.c;
    if (num_windows) *num_windows = g.nwin;
    if (chunk) *chunk = g.chunk0;
    return ZKB_OK;
}
int zkb_transfer_stats(uint64_t* h2d_bytes, int reset) {
    Ctx* all = per_device_array<Ctx>();
    uint64_t total = 0;
    for (int i = 0; i < ZKB_MAX_DEVICES; ++i) {
        total += all[i].h2d_bytes.load();
        if (reset) all[i].h2d_bytes.store(0);
    }
    if (h2d_bytes) *h2d_bytes = total;
    return ZKB_OK;
}
int zkb_msm_last_entries(uint64_t* entries) {
    ZKB_TRY(check_ptr(entries, "entries"));
    *entries = msm_workspace().last_entries;
    return ZKB_OK;
}
int zkb_prof_enable(int on) {
    ctx().prof_on = on != 0;
    return ZKB_OK;
}
static void prof_drain(ProfTimer& t) {
    Ctx& c = ctx();
    for (auto& pr : t.pending) {
        float ms = 0;
        cudaEventSynchronize(pr.second);
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) t.total_ms += ms;
        else cudaGetLastError();
        c.event_pool.push_back(pr.first);
        c.event_pool.push_back(pr.second);
    }
    t.pending.clear();
}
int zkb_prof_reset(void) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    for (auto& kv : ctx().timers) { prof_drain(kv.second); kv.second.total_ms = 0; kv.second.launches = 0; }
    return ZKB_OK;
}
int zkb_prof_get(const char* name, double* total_ms, uint64_t* launches) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(check_ptr(name, "name"));
    auto it = ctx().timers.find(name);
    if (it == ctx().timers.end()) {
        if (total_ms) *total_ms = 0;
        if (launches) *launches = 0;
        return ZKB_OK;
    }
    prof_drain(it->second);
    if (total_ms) *total_ms = it->second.total_ms;
    if (launches) *launches = it->second.launches;
    return ZKB_OK;
}
uint64_t zkb_launch_count(void) {  // all device slots
    uint64_t total = 0;
    Ctx* all = per_device_array<Ctx>();
    for (int i = 0; i < ZKB_MAX_DEVICES; ++i) total += all[i].launches.load();
    return total;
}
int zkb_measure_imad_peak(double* wide_macs_per_s) {
    std::lock_guard<std::recursive_mutex> lock(ctx().mu);
    ZKB_TRY(require_init());
    ZKB_TRY(check_ptr(wide_macs_per_s, "wide_macs_per_s"));
    return measure_imad_peak(wide_macs_per_s);
}

}  // extern "C"
