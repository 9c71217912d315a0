// Host emulation of the small part of the CUDA execution model that csrc/orb.cu uses — TEST INFRASTRUCTURE.
// It lets the CPU-only build container run the *same* kernels and the same launch sequence (vo_orb_extract) against
// the pinned CPU restatement: blocks run one after another; the threads of a block run sequentially (kernels without a
// barrier) or as real OS threads meeting at a barrier (kernels launched with VO_LAUNCH_BAR).  __shared__ becomes a
// function-local static (one block at a time, so it is private to the running block), atomics are host atomics.
// It does not model warps, memory spaces or timing: it checks indexing, control flow and arithmetic, nothing else.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <math.h>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyDeviceToHost = 2, cudaMemcpyDefault = 4 };
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 { unsigned x, y, z; };
inline thread_local emu_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

using std::max;
using std::min;
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline int __float2int_rn(float x) { return (int)lrintf(x); }
inline int __double2int_rn(double x) { return (int)lrint(x); }

namespace vo_emu {
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, waiting = 0, phase = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int ph = phase;
        if (++waiting == n) { waiting = 0; ++phase; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != ph; });
    }
};
inline Barrier *g_barrier = nullptr;
inline long long g_launches = 0;

inline void launch(dim3 grid, dim3 block, bool barriers, const std::function<void()> &body) {
    ++g_launches;
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                auto run = [&](unsigned t) {
                    threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    blockIdx = {bx, by, bz};
                    blockDim = block;
                    gridDim = grid;
                    body();
                };
                if (!barriers) {
                    for (unsigned t = 0; t < nthreads; ++t) run(t);
                } else {
                    Barrier bar;
                    bar.n = (int)nthreads;
                    g_barrier = &bar;
                    std::vector<std::thread> th;
                    th.reserve(nthreads);
                    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(run, t);
                    for (auto &x : th) x.join();
                    g_barrier = nullptr;
                }
            }
}
}  // namespace vo_emu
inline void __syncthreads() {
    if (!vo_emu::g_barrier) abort();   // a kernel with a barrier was launched with VO_LAUNCH instead of VO_LAUNCH_BAR
    vo_emu::g_barrier->wait();
}
