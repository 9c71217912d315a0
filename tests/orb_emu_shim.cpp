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
