// ubench.cu — instruction-throughput and field-multiplication microbenchmarks for sm_100a (standalone binary).
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I zksnap-circuits-halo2_b200/csrc -I tools tools/ubench.cu -o tools/ubench
#include <cstdio>
#include <cuda_runtime.h>

#include "field29.cuh"

using namespace zkb;

#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) instr_kernel(uint32_t* out, uint32_t x, uint32_t y) {
    uint32_t a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    x += threadIdx.x;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {  // 8 independent mad.wide.u32 chains, multiplicand = own low word (not hoistable)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint64_t acc = ((uint64_t)a[2 * j + 1] << 32) | a[2 * j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[2 * j]), "r"(y));
                a[2 * j] = (uint32_t)acc; a[2 * j + 1] = (uint32_t)(acc >> 32);
            }
        } else if (MODE == 1) {  // two 8-limb carry chains of mad.lo.cc / madc.hi.cc pairs (-> IMAD.WIDE.U32.X)
            asm volatile(
                "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
                "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
                "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
                "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
                "mad.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
                "madc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
                "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
                "madc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;\n\t"
                : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                  "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
                : "r"(x), "r"(y));
        } else if (MODE == 2) {  // 8 independent mad.lo.u32
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(x), "r"(y));
        } else if (MODE == 3) {  // 8 independent mad.hi.u32
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(x), "r"(y));
        } else if (MODE == 4) {  // two 8-limb add.cc / addc chains (IADD3 / IADD3.X)
            asm volatile(
                "add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %16;\n\taddc.cc.u32 %3, %3, %17;\n\t"
                "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %16;\n\taddc.u32 %7, %7, %17;\n\t"
                "add.cc.u32 %8, %8, %16;\n\taddc.cc.u32 %9, %9, %17;\n\taddc.cc.u32 %10, %10, %16;\n\taddc.cc.u32 %11, %11, %17;\n\t"
                "addc.cc.u32 %12, %12, %16;\n\taddc.cc.u32 %13, %13, %17;\n\taddc.cc.u32 %14, %14, %16;\n\taddc.u32 %15, %15, %17;\n\t"
                : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                  "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
                : "r"(x), "r"(y));
        } else if (MODE == 5) {  // 4 wide-MAC chains + 8 xor/shift chains (ALU pipe): do the two pipes overlap?
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t acc = ((uint64_t)a[2 * j + 1] << 32) | a[2 * j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[2 * j]), "r"(y));
                a[2 * j] = (uint32_t)acc; a[2 * j + 1] = (uint32_t)(acc >> 32);
            }
#pragma unroll
            for (int j = 8; j < 16; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(x), "r"(a[(j + 1) & 15]));
        } else if (MODE == 6) {  // the 8 LOP3 alone
#pragma unroll
            for (int j = 8; j < 16; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(x), "r"(a[(j + 1) & 15]));
        } else if (MODE == 7) {  // the 4 wide-MAC chains alone
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t acc = ((uint64_t)a[2 * j + 1] << 32) | a[2 * j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[2 * j]), "r"(y));
                a[2 * j] = (uint32_t)acc; a[2 * j + 1] = (uint32_t)(acc >> 32);
            }
        } else if (MODE == 8) {  // 4 wide-MAC chains + 8 IADD3 (3-input adds)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t acc = ((uint64_t)a[2 * j + 1] << 32) | a[2 * j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[2 * j]), "r"(y));
                a[2 * j] = (uint32_t)acc; a[2 * j + 1] = (uint32_t)(acc >> 32);
            }
#pragma unroll
            for (int j = 8; j < 16; ++j) a[j] = a[j] + a[(j + 1) & 15] + x;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;
}

// dependent chains of field multiplications, CH independent chains per thread
template <int KIND, int CH, int PERTHREAD>
__global__ void __launch_bounds__(128) mul_kernel(uint32_t* out, int iters) {
    if (KIND == 0) {
        Fq x[CH], y;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[c].l[i] = threadIdx.x * 977u + i * 131u + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) y.l[i] = blockIdx.x * 31u + i + (PERTHREAD ? threadIdx.x * 7u : 0u);
        x[0].l[7] &= 0x0fffffff; y.l[7] &= 0x0fffffff;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fp_mul(x[c], y);
        uint32_t r = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) r ^= x[c].l[i];
        if (r == 0x12345678u) out[0] = r;
    } else {
        Fq29 x[CH], y;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 9; ++i) x[c].l[i] = (threadIdx.x * 977u + i * 131u + c) & MASK29;
#pragma unroll
        for (int i = 0; i < 9; ++i) y.l[i] = (blockIdx.x * 31u + i + (PERTHREAD ? threadIdx.x * 7u : 0u)) & MASK29;
        x[0].l[8] &= 0xfffff; y.l[8] &= 0xfffff;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = mul29(x[c], y);
        uint32_t r = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int i = 0; i < 9; ++i) r ^= x[c].l[i];
        if (r == 0x12345678u) out[0] = r;
    }
}

template <class F>
static float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t* d;
    cudaMalloc(&d, 64);
    const char* names[9] = {"IMAD.WIDE.U32 (mad.wide, no carry)", "IMAD.WIDE.U32.X (mad.lo.cc/madc.hi.cc chain)", "IMAD lo", "IMAD.HI",
                            "IADD3.X chain (add.cc/addc)", "4 IMAD.WIDE + 8 LOP3", "8 LOP3 alone", "4 IMAD.WIDE alone", "4 IMAD.WIDE + 8 IADD3"};
    double per_iter[9] = {8, 8, 8, 8, 16, 12, 8, 4, 12};  // "useful" warp-instructions per loop iteration (pairs fused for mode 1)
    int blocks = sms * 8, threads = 256;
    for (int mode = 0; mode < 9; ++mode) {
        float ms = 0;
        switch (mode) {
            case 0: ms = time_it([&] { instr_kernel<0><<<blocks, threads>>>(d, 3, 5); }); break;
            case 1: ms = time_it([&] { instr_kernel<1><<<blocks, threads>>>(d, 3, 5); }); break;
            case 2: ms = time_it([&] { instr_kernel<2><<<blocks, threads>>>(d, 3, 5); }); break;
            case 3: ms = time_it([&] { instr_kernel<3><<<blocks, threads>>>(d, 3, 5); }); break;
            case 4: ms = time_it([&] { instr_kernel<4><<<blocks, threads>>>(d, 3, 5); }); break;
            case 5: ms = time_it([&] { instr_kernel<5><<<blocks, threads>>>(d, 3, 5); }); break;
            case 6: ms = time_it([&] { instr_kernel<6><<<blocks, threads>>>(d, 3, 5); }); break;
            case 7: ms = time_it([&] { instr_kernel<7><<<blocks, threads>>>(d, 3, 5); }); break;
            case 8: ms = time_it([&] { instr_kernel<8><<<blocks, threads>>>(d, 3, 5); }); break;
        }
        double warp_instr = (double)blocks * (threads / 32) * ITERS * per_iter[mode];
        double clk = prop.clockRate * 1e3;  // Hz (max)
        double per_clk_sm = warp_instr * 32 / (ms * 1e-3) / clk / sms;
        printf("%-48s %8.3f ms  %7.1f lane-ops/clk/SM (at %.0f MHz)  %.3e lane-ops/s\n", names[mode], ms, per_clk_sm, clk / 1e6,
               warp_instr * 32 / (ms * 1e-3));
    }
    int iters = 512;
    struct { const char* name; float ms; int ch; } res[8];
    res[0] = {"fp_mul 8x32 (carry chains), 1 chain, uniform y", time_it([&] { mul_kernel<0, 1, 0><<<sms * 16, 128>>>(d, iters); }), 1};
    res[1] = {"fp_mul 8x32 (carry chains), 2 chains, uniform y", time_it([&] { mul_kernel<0, 2, 0><<<sms * 16, 128>>>(d, iters); }), 2};
    res[2] = {"mul29 9x29 (carry-free), 1 chain, uniform y", time_it([&] { mul_kernel<1, 1, 0><<<sms * 16, 128>>>(d, iters); }), 1};
    res[3] = {"mul29 9x29 (carry-free), 2 chains, uniform y", time_it([&] { mul_kernel<1, 2, 0><<<sms * 16, 128>>>(d, iters); }), 2};
    res[4] = {"fp_mul 8x32, 1 chain, per-thread y", time_it([&] { mul_kernel<0, 1, 1><<<sms * 16, 128>>>(d, iters); }), 1};
    res[5] = {"fp_mul 8x32, 2 chains, per-thread y", time_it([&] { mul_kernel<0, 2, 1><<<sms * 16, 128>>>(d, iters); }), 2};
    res[6] = {"mul29 9x29, 1 chain, per-thread y", time_it([&] { mul_kernel<1, 1, 1><<<sms * 16, 128>>>(d, iters); }), 1};
    res[7] = {"mul29 9x29, 2 chains, per-thread y", time_it([&] { mul_kernel<1, 2, 1><<<sms * 16, 128>>>(d, iters); }), 2};
    for (auto& r : res) {
        double muls = (double)sms * 16 * 128 * iters * r.ch;
        printf("%-50s %8.3f ms  %.3e modmul/s\n", r.name, r.ms, muls / (r.ms * 1e-3));
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
