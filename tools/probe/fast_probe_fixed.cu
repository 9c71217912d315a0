#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "orb_math.cuh"
static int ref_score(int v, const uint8_t* ring, int thr) {   // plain reference, host only
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int lo = 1000, lob = 1000;
        for (int j = 0; j < 9; ++j) { int d = v - ring[(k + j) & 15]; if (d < lo) lo = d; if (-d < lob) lob = -d; }
        int m = lo > lob ? lo : lob;
        if (m > best) best = m;
    }
    return best > thr ? best - 1 : 0;
}
__global__ void k(const uint8_t* in, int n, int thr, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t ring[16];
    for (int j = 0; j < 16; ++j) ring[j] = in[i * 17 + 1 + j];
    out[i] = vo::orb::fast_corner_score(in[i * 17], ring, thr);
}
int main() {
    const int n = 1 << 20;
    std::vector<uint8_t> h((size_t)n * 17);
    srand(1);
    for (auto& v : h) v = rand() & 255;
    for (int i = 0; i < n; i += 2) { int v = h[i*17]; int len = 7 + rand() % 6; for (int j = 0; j < len; ++j) h[i*17+1+((j + i) & 15)] = (uint8_t)(v > 128 ? v - 15 - (rand() % 60) : v + 15 + (rand() % 60)); }
    uint8_t* d; int* o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, n * 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    k<<<n / 128, 128>>>(d, n, 20, o);
    std::vector<int> r(n);
    cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0, badh = 0, nz = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t ring[16];
        for (int j = 0; j < 16; ++j) ring[j] = h[(size_t)i * 17 + 1 + j];
        int want = ref_score(h[(size_t)i * 17], ring, 20);
        int hostv = vo::orb::fast_corner_score(h[(size_t)i * 17], ring, 20);
        nz += want != 0;
        if (hostv != want) ++badh;
        if (want != r[i]) { if (bad < 5) printf("i=%d dev %d ref %d\n", i, r[i], want); ++bad; }
    }
    printf("device bad %d, host-header bad %d / %d (corners %d) (%s)\n", bad, badh, n, nz, cudaGetErrorString(cudaGetLastError()));
    return bad != 0;
}
