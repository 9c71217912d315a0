// Host build of the device math header (csrc/pnp_math.cuh) so the CPU-only container can check, bit for bit,
// that the code the kernels run equals the oracle's restatement.  Test infrastructure only.
#include "../visual-odometry-pipeline_b200/csrc/pnp_math.cuh"
#include "../visual-odometry-pipeline_b200/csrc/hamming_math.cuh"
extern "C" {
int hm_p3p4(const double *P, const double *uv, const double *K, double *out) {
    double Pm[4][3], uvm[4][2];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) Pm[i][j] = P[3 * i + j];
        uvm[i][0] = uv[2 * i]; uvm[i][1] = uv[2 * i + 1];
    }
    vo::PoseD best;
    if (!vo::p3p_solve4(Pm, uvm, K[0], K[1], K[2], K[3], best)) return 0;
    for (int j = 0; j < 9; ++j) out[j] = best.r[j];
    for (int j = 0; j < 3; ++j) out[9 + j] = best.t[j];
    return 1;
}
void hm_draw(unsigned long long seed, long long pair, int h, int n, int *out) { vo::draw_hypothesis(seed, pair, h, n, out); }
int hm_is_inlier(const float *pose, const float *k, float thr, float X, float Y, float Z, float u, float v) {
    vo::PoseF p; vo::IntrF kk{k[0], k[1], k[2], k[3]};
    for (int j = 0; j < 9; ++j) p.r[j] = pose[j];
    for (int j = 0; j < 3; ++j) p.t[j] = pose[9 + j];
    const vo::ScoreModel m = vo::score_model(p, kk);
    return vo::is_inlier(m, thr, X, Y, Z, u - kk.cx, v - kk.cy) ? 1 : 0;
}
// 256-bit Hamming distance exactly as match_u8_kernel forms it: both descriptors to prefix-XOR form, the 13-operation
// adder tree, four weighted popcounts.
int hm_hamming256(const unsigned int *a_in, const unsigned int *b_in) {
    unsigned int a[8], b[8];
    for (int w = 0; w < 8; ++w) { a[w] = a_in[w]; b[w] = b_in[w]; }
    vo::hamming_prefix_form(a);
    vo::hamming_prefix_form(b);
    const vo::HammingPlanes p = vo::hamming_planes(a, b);
    return __builtin_popcount(p.ones_a) + __builtin_popcount(p.ones_b) + 2 * __builtin_popcount(p.twos) +
           4 * __builtin_popcount(p.fours);
}
void hm_hamming256_many(const unsigned int *a, const unsigned int *b, int n, int *out) {
    for (int i = 0; i < n; ++i) out[i] = hm_hamming256(a + 8 * i, b + 8 * i);
}
}
