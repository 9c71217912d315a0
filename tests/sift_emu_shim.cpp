// csrc/sift.cu compiled for the host on top of tests/cuda_emu.h (same kernels, same launch sequence), compared with the
// CPU restatement by tests/test_sift_emulation.py.  Test infrastructure only.
// Build: g++ -x c++ -std=c++17 -O2 -ffp-contract=off -pthread -DVO_HOST_EMU='"cuda_emu.h"' -I tests
#include "../visual-odometry-pipeline_b200/csrc/sift.cu"
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
int emu_sift_run(const unsigned char *image, int H, int W, int channels, int cap, float *kp, float *desc, float *aux, int *count2) {
    vo_ctx ctx;
    memset(&ctx, 0, sizeof(ctx));
    vo_sift_config cfg = {H, W, cap};
    vo_sift *s = nullptr;
    int rc = vo_sift_create(&ctx, &cfg, &s);
    if (rc) return -100 + rc;
    rc = vo_sift_extract(s, image, channels, kp, desc, aux, count2, nullptr);
    vo_sift_destroy(s);
    return rc ? -200 + rc : count2[0];
}
}

#ifdef VO_EMU_MAIN   // sanitizer driver: sift_emu <raw image file> H W channels
int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int H = atoi(argv[2]), W = atoi(argv[3]), ch = atoi(argv[4]), cap = 20000;
    std::vector<unsigned char> img((size_t)H * W * ch);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    std::vector<float> kp((size_t)cap * 2), aux((size_t)cap * 4), desc((size_t)cap * 128);
    int cnt[2] = {0, 0};
    const int n = emu_sift_run(img.data(), H, W, ch, cap, kp.data(), desc.data(), aux.data(), cnt);
    double s = 0;
    for (int i = 0; i < n * 128; ++i) s += desc[i];
    printf("%d keypoints (%d raw), descriptor sum %.0f, %s\n", n, cnt[1], s, n < 0 ? emu_last_error() : "ok");
    return n < 0;
}
#endif
