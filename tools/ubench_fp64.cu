// ubench_fp64.cu — does the FP64 pipe of sm_100a (B200) run beside the integer pipe?  Standalone binary.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench_fp64.cu -o tools/ubench_fp64
// Modes: 0 DFMA chains alone, 1 IMAD.WIDE.U32.X carry chains alone, 2 both interleaved in one thread,
//        3 plain IMAD.WIDE.U32 alone, 4 DFMA + plain IMAD.WIDE.U32, 5 DADD alone, 6 FFMA alone, 7 FFMA + IMAD.WIDE.X
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t x, uint32_t y, double fx, double fy) {
    uint32_t a[16];
    double d[8];
    float f[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
#pragma unroll
    for (int j = 0; j < 8; ++j) { d[j] = threadIdx.x * 1e-3 + j; f[j] = threadIdx.x * 1e-3f + j; }
    x += threadIdx.x;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(fx), "d"(fy));
        }
        if (MODE == 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(fx));
        }
        if (MODE == 6 || MODE == 7) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"((float)fx), "f"((float)fy));
        }
        if (MODE == 1 || MODE == 2 || MODE == 7) {
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
        }
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint64_t acc = ((uint64_t)a[2 * j + 1] << 32) | a[2 * j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[2 * j]), "r"(y));
                a[2 * j] = (uint32_t)acc; a[2 * j + 1] = (uint32_t)(acc >> 32);
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) r ^= a[j];
    double s = 0; float sf = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += d[j]; sf += f[j]; }
    if (r == 0x12345678u || s == 1.2345 || sf == 1.2345f) out[0] = r;
}

template <class F> static float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, 64);
    int blocks = sms * 8, threads = 256;
    const char* names[8] = {"8 DFMA", "8 IMAD.WIDE.U32.X (carry chains)", "8 DFMA + 8 IMAD.WIDE.U32.X", "8 IMAD.WIDE.U32", "8 DFMA + 8 IMAD.WIDE.U32",
                            "8 DADD", "8 FFMA", "8 FFMA + 8 IMAD.WIDE.U32.X"};
    float ms[8];
    ms[0] = time_it([&] { k<0><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[1] = time_it([&] { k<1><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[2] = time_it([&] { k<2><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[3] = time_it([&] { k<3><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[4] = time_it([&] { k<4><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[5] = time_it([&] { k<5><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[6] = time_it([&] { k<6><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    ms[7] = time_it([&] { k<7><<<blocks, threads>>>(d, 3, 5, 1.0000001, 1e-9); });
    double clk = prop.clockRate * 1e3;
    for (int m = 0; m < 8; ++m) {
        double cyc_per_iter = ms[m] * 1e-3 * clk / ((double)blocks * (threads / 32) / (sms * 4.0)) / ITERS;  // SMSP cycles per warp-iteration
        printf("%-36s %8.3f ms   %7.2f SMSP-cycles per warp-iteration\n", names[m], ms[m], cyc_per_iter);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
