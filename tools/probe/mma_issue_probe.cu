// Microbenchmark: how long does ONE thread take to issue a tcgen05.mma, and how long does the tensor pipe take to
// retire it?  For each (kind, N): `issue` = cycles per instruction spent in the issuing thread for a back-to-back stream
// of 256 MMAs, `total` = cycles per instruction until the commit arrives.  Operands: zero-filled shared memory (A and B
// K-major SWIZZLE_128B tiles), accumulator in TMEM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_issue_probe mma_issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>  // 0: tf32 (K=8), 1: f16 (K=16)
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N>
__global__ void __launch_bounds__(128) probe(long long *out, int reps) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *sm = raw + (base - smem_u32(raw));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (KIND == 0 ? ((2u << 7) | (2u << 10)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t a = sdesc(base), b = sdesc(base + 128 * 128);
        uint32_t phase = 0;
        long long issue = 0, total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll 4
            for (int i = 0; i < 256; ++i) mma<KIND>(tmem + (uint32_t)((i & 1) * 256), a + (uint64_t)((i & 3) * 2), b + (uint64_t)((i & 3) * 2), idesc, i > 1);
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            const long long t2 = clock64();
            phase ^= 1u;
            if (r > 0) { issue += t1 - t0; total += t2 - t0; }
        }
        out[0] = issue; out[1] = total;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int N>
void run(const char *name, long long *d) {
    const int reps = 9, smem = (128 + 256) * 128 + 1024;
    cudaFuncSetAttribute(probe<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<KIND, N><<<1, 128, smem>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = 256.0 * (reps - 1);
    const double k = KIND == 0 ? 8 : 16;
    printf("%-10s N=%3d  issue %6.1f cyc/MMA   total %6.1f cyc/MMA  -> %6.0f FLOP/clk/SM  (%s)\n", name, N, h[0] / n, h[1] / n,
           2.0 * 128 * N * k / (h[1] / n), cudaGetErrorString(e));
}
int main() {
    long long *d;
    cudaMalloc(&d, 64);
    run<0, 32>("tf32", d); run<0, 64>("tf32", d); run<0, 128>("tf32", d); run<0, 160>("tf32", d); run<0, 256>("tf32", d);
    run<1, 32>("f16", d); run<1, 64>("f16", d); run<1, 128>("f16", d); run<1, 192>("f16", d); run<1, 256>("f16", d);
    return 0;
}
