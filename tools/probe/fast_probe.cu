#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "orb_math.cuh"
__global__ void k(const uint8_t* in, int n, int thr, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t ring[16];
    for (int j = 0; j < 16; ++j) ring[j] = in[i * 17 + 1 + j];
    out[i] = vo::orb::fast_corner_score(in[i * 17], ring, thr);
}
int main() {
    const int n = 1 << 16;
    std::vector<uint8_t> h(n * 17);
    srand(1);
    for (auto& v : h) v = rand() & 255;
    // make half of them corner-ish
    for (int i = 0; i < n; i += 2) { int v = h[i*17]; for (int j = 0; j < 12; ++j) h[i*17+1+((j + i) & 15)] = (uint8_t)(v > 128 ? v - 30 - (rand() % 60) : v + 30 + (rand() % 60)); }
    uint8_t* d; int* o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, n * 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    k<<<n / 128, 128>>>(d, n, 20, o);
    std::vector<int> r(n);
    cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t ring[16];
        for (int j = 0; j < 16; ++j) ring[j] = h[i * 17 + 1 + j];
        int want = vo::orb::fast_corner_score(h[i * 17], ring, 20);
        if (want != r[i]) { if (bad < 5) printf("i=%d dev %d host %d\n", i, r[i], want); ++bad; }
    }
    printf("bad %d / %d  (%s)\n", bad, n, cudaGetErrorString(cudaGetLastError()));
}
