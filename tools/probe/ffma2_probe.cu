// Microbenchmark: throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) and how each mixes with ALU-pipe work.
// Prints FMA lanes per clock per SM for each variant.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

template <int MODE>  // 0: FFMA x8, 1: FFMA2 x8, 2: FFMA x8 + LOP3 x2, 3: FFMA2 x8 + LOP3 x2, 4: FFMA2 x8 + LOP3 x4, 5: FFMA2 x4 + FFMA x4
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
    float a[8]; u64 p[8]; unsigned z[4];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    for (int i = 0; i < 4; ++i) z[i] = threadIdx.x + i;
    const u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma1(a[i], s, s);
            } else if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], s2, s2); a[i] = fma1(a[i], s, s); }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s2, s2);
            }
            if (MODE == 2 || MODE == 3) { z[0] = lop(z[0], z[1], it); z[1] = lop(z[1], z[0], r); }
            if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) z[i] = lop(z[i], z[(i + 1) & 3], it);
            }
        }
    }
    float acc = 0; for (int i = 0; i < 8; ++i) acc += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + z[0] + z[1] + z[2] + z[3];
}
template <int MODE> void run(const char *name, int fma_per_round, float *d, int sms, int khz) {
    const int iters = 20000, ctas = sms * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(d, 1000, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<ctas, 256>>>(d, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)ctas * 256 * iters * 4 * fma_per_round;
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-28s %8.3f ms  %7.1f FMA lanes/clk/SM  (%.1f TFLOP/s)\n", name, ms, fmas / clk / sms, 2 * fmas / (ms * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", pr.name, pr.multiProcessorCount, khz);
    float *d; cudaMalloc(&d, pr.multiProcessorCount * 4 * 256 * 4);
    const int sms = pr.multiProcessorCount;
    run<0>("FFMA x8", 8, d, sms, khz);
    run<1>("FFMA2 x8", 16, d, sms, khz);
    run<2>("FFMA x8 + LOP3 x2", 8, d, sms, khz);
    run<3>("FFMA2 x8 + LOP3 x2", 16, d, sms, khz);
    run<4>("FFMA2 x8 + LOP3 x4", 16, d, sms, khz);
    run<5>("FFMA2 x4 + FFMA x4", 12, d, sms, khz);
    return 0;
}
