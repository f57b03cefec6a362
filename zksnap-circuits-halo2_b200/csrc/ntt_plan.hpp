// ntt_plan.hpp — host-side geometry of the multi-pass NTT (shared by the CUDA driver and the CPU emulator).
#pragma once
#include <cstdint>
#include <cstdlib>

#include "ntt.cuh"

namespace zkb {

struct NttGeometry {
    uint32_t log_n;
    uint32_t npass;
    uint32_t lr[NTT_MAX_PASSES];
    uint32_t log_t[NTT_MAX_PASSES];
    uint32_t tw_h;
};

// R <= 2^10 in a single pass, otherwise ceil(log_n / 9) passes with radices as equal as possible.
// small_first: the smaller radices go to the first passes — used by the sharded NTT, whose pass 0 gathers its rows from
// peer HBM over NVLink: a smaller R_1 means wider tiles (T = 1024 / R_1 adjacent columns), i.e. 128-256 byte row segments
// instead of 64 (profiles/r1_dist_ntt_8gpu.txt).
inline NttGeometry ntt_geometry(uint32_t log_n, bool small_first = false) {
    NttGeometry g{};
    g.log_n = log_n;
    if (log_n <= 10) {
        g.npass = 1;
        g.lr[0] = log_n;
    } else {
        g.npass = (log_n + 8) / 9;
        uint32_t base = log_n / g.npass, rem = log_n % g.npass;
        for (uint32_t p = 0; p < g.npass; ++p) g.lr[p] = base + ((small_first ? p >= g.npass - rem : p < rem) ? 1 : 0);
    }
    // experiment hook: ZKB_NTT_GEOM="8,10,8" forces the radices of the transform whose log_n equals their sum
    if (const char* e = getenv("ZKB_NTT_GEOM")) {
        uint32_t lr[NTT_MAX_PASSES] = {0, 0, 0, 0}, np = 0, sum = 0;
        for (const char* c = e; *c && np < NTT_MAX_PASSES;) {
            lr[np] = (uint32_t)strtoul(c, nullptr, 10);
            sum += lr[np++];
            while (*c && *c != ',') ++c;
            if (*c == ',') ++c;
        }
        bool ok = sum == log_n && np >= 1;
        for (uint32_t p = 0; p < np; ++p) ok = ok && lr[p] >= 1 && lr[p] <= 10;
        if (ok) {
            g.npass = np;
            for (uint32_t p = 0; p < NTT_MAX_PASSES; ++p) g.lr[p] = p < np ? lr[p] : 0;
        }
    }
    g.tw_h = (log_n + 1) / 2;
    for (uint32_t p = 0; p < g.npass; ++p) {
        uint32_t want = g.lr[p] >= 10 ? 0 : 10 - g.lr[p];  // R*T = 1024 elements (32 KiB) per CTA, 4 CTAs per SM
        uint32_t cap;
        if (p + 1 < g.npass) {  // strided pass: T <= S_p
            cap = 0;
            for (uint32_t q = p + 1; q < g.npass; ++q) cap += g.lr[q];
        } else {  // final pass: T <= Q_P
            cap = log_n - g.lr[p];
        }
        g.log_t[p] = want < cap ? want : cap;
    }
    return g;
}

// One transform sharded over 2^log_g GPUs (ntt.cuh "distributed mode"): needs >= 2 passes (pass 0 is the exchange
// step), at least one pass-0 tile per rank, and T-wide k_1 tiles of the final pass inside one rank's k_1 block.
inline bool ntt_dist_supported(const NttGeometry& g, uint32_t log_g) {
    if (log_g == 0) return true;
    if (log_g > 3 || g.npass < 2) return false;
    uint32_t log_s1 = g.log_n - g.lr[0];
    if (log_s1 < g.log_t[0] + log_g) return false;
    if (g.lr[0] < log_g + g.log_t[g.npass - 1]) return false;
    return true;
}

inline uint32_t ntt_cta_threads(const NttGeometry& g, uint32_t p) {
    uint32_t e = 1u << (g.lr[p] + g.log_t[p]);
    uint32_t t = e / 8;
    return t < 32 ? 32 : (t > 128 ? 128 : t);
}
inline uint64_t ntt_cta_count(const NttGeometry& g, uint32_t p) {
    return (1ull << g.log_n) >> (g.lr[p] + g.log_t[p]);
}
inline size_t ntt_cta_smem_bytes(const NttGeometry& g, uint32_t p) {
    uint32_t R = 1u << g.lr[p];
    bool fin = (p + 1 == g.npass);
    size_t plane = fin ? ((size_t)(R + 4) << g.log_t[p]) : ((size_t)R << g.log_t[p]);  // R + 4 >= every ntt_final_pitch
    return plane * 2 * 16;
}

}  // namespace zkb
