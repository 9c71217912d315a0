// csrc/orb.cu compiled for the host on top of tests/cuda_emu.h: the same kernels and the same launch sequence
// (vo_orb_create / vo_orb_extract), run against the pinned CPU restatement by tests/test_orb_emulation.py.
// Test infrastructure only.  Build: g++ -x c++ -std=c++17 -O2 -ffp-contract=off -pthread -DVO_HOST_EMU='"cuda_emu.h"' -I tests
#include "../visual-odometry-pipeline_b200/csrc/orb.cu"
#include <stdio.h>

namespace vo {
static char g_err[512];
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }
void clear_error() { g_err[0] = 0; }
}  // namespace vo

extern "C" {
const char *emu_last_error() { return vo::get_error(); }
long long emu_launches() { return vo_emu::g_launches; }
// returns rows written (>= 0) or a negative error code; outputs must hold nlevels * 4096 rows
int emu_orb_run(const unsigned char *image, int H, int W, int channels, int nfeatures, int nlevels, int fast_threshold,
                float *kp, unsigned char *desc, float *aux, int *count2) {
    vo_ctx ctx;
    memset(&ctx, 0, sizeof(ctx));
    vo_orb_config cfg = {H, W, nfeatures, nlevels, fast_threshold};
    vo_orb *o = nullptr;
    int rc = vo_orb_create(&ctx, &cfg, &o);
    if (rc) return -100 + rc;
    rc = vo_orb_extract(o, image, channels, kp, desc, aux, count2, nullptr);
    const int cap = vo_orb_capacity(o);
    vo_orb_destroy(o);
    if (rc) return -200 + rc;
    return count2[0] <= cap ? count2[0] : -300;
}
}

#ifdef VO_EMU_MAIN   // stand-alone driver for sanitizer runs: orb_emu <raw image file> H W channels [nfeatures]
int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int H = atoi(argv[2]), W = atoi(argv[3]), ch = atoi(argv[4]), nf = argc > 5 ? atoi(argv[5]) : 500;
    std::vector<unsigned char> img((size_t)H * W * ch);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    const int cap = 8 * 4096;
    std::vector<float> kp((size_t)cap * 2), aux((size_t)cap * 4);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int cnt[2] = {0, 0};
    const int n = emu_orb_run(img.data(), H, W, ch, nf, 8, 20, kp.data(), desc.data(), aux.data(), cnt);
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < n * 32; ++i) h = (h ^ desc[i]) * 1099511628211ull;
    printf("%d keypoints, overflow %d, descriptor hash %016llx, %s\n", n, cnt[1], h, n < 0 ? emu_last_error() : "ok");
    return n < 0;
}
#endif
