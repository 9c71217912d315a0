// Host build of the device math header (csrc/pnp_math.cuh) so the CPU-only container can check, bit for bit,
// that the code the kernels run equals the oracle's restatement.  Test infrastructure only.
#include "../visual-odometry-pipeline_b200/csrc/pnp_math.cuh"
#include "../visual-odometry-pipeline_b200/csrc/hamming_math.cuh"
#include "../visual-odometry-pipeline_b200/csrc/orb_math.cuh"
#include "../visual-odometry-pipeline_b200/csrc/pnp_ref_math.cuh"
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
// ---- ORB front-end arithmetic (csrc/orb_math.cuh)
int hm_orb_umax(int v) { return vo::orb::umax(v); }
int hm_orb_gray(int b, int g, int r) { return vo::orb::bgr_to_gray((unsigned char)b, (unsigned char)g, (unsigned char)r); }
void hm_orb_coeff(int dst, int src, int *ofs, int *c1, int *inside) {
    for (int x = 0; x < dst; ++x) { bool in; vo::orb::linear_exact_coeff(x, dst, src, ofs[x], c1[x], in); inside[x] = in; }
}
// whole-image INTER_LINEAR_EXACT through the per-pixel function
void hm_orb_resize(const unsigned char *src, int W, int H, unsigned char *dst, int dw, int dh) {
    for (int y = 0; y < dh; ++y) {
        int oy, cy; bool iny; vo::orb::linear_exact_coeff(y, dh, H, oy, cy, iny);
        const int oy1 = oy + 1 < H ? oy + 1 : H - 1;
        for (int x = 0; x < dw; ++x) {
            int ox, cx; bool inx; vo::orb::linear_exact_coeff(x, dw, W, ox, cx, inx);
            const int ox1 = ox + 1 < W ? ox + 1 : W - 1;
            dst[y * dw + x] = vo::orb::linear_exact_pixel(src[oy * W + ox], src[oy * W + ox1], src[oy1 * W + ox], src[oy1 * W + ox1], cx, inx, cy, iny);
        }
    }
}
// FAST score map (3-pixel frame left at 0)
void hm_orb_fast_map(const unsigned char *img, int W, int H, int thr, int *out) {
    static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int i = 0; i < W * H; ++i) out[i] = 0;
    for (int y = 3; y < H - 3; ++y)
        for (int x = 3; x < W - 3; ++x) {
            unsigned char ring[16];
            for (int k = 0; k < 16; ++k) ring[k] = img[(y + dy[k]) * W + x + dx[k]];
            out[y * W + x] = vo::orb::fast_corner_score(img[y * W + x], ring, thr);
        }
}
float hm_orb_harris(int a, int b, int c) { return vo::orb::harris_response(a, b, c); }
float hm_orb_atan2(float y, float x) { return vo::orb::fast_atan2(y, x); }
void hm_orb_blur(const unsigned char *img, int W, int H, unsigned char *out) {   // reflect-101 border
    float k[7]; vo::orb::gaussian_kernel(k);
    auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
    float *h = new float[(size_t)W * H];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            unsigned char p[7];
            for (int i = 0; i < 7; ++i) p[i] = img[y * W + refl(x + i - 3, W)];
            h[y * W + x] = vo::orb::blur_row(k, p);
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float c[7];
            for (int i = 0; i < 7; ++i) c[i] = h[refl(y + i - 3, H) * W + x];
            out[y * W + x] = vo::orb::blur_col(k, c);
        }
    delete[] h;
}
void hm_orb_rotate(float angle_deg, const int *pat, int n, int *ix, int *iy) {
    float ca, sb; vo::orb::angle_cos_sin(angle_deg, ca, sb);
    for (int i = 0; i < n; ++i) vo::orb::rotate_pattern_point(pat[2 * i], pat[2 * i + 1], ca, sb, ix[i], iy[i]);
}
// ---- reference-sampler PnP-RANSAC arithmetic (csrc/pnp_ref_math.cuh)
void hm_ref_table(int n, int iters, int *out) { vo::refpnp::mwc_table(n, iters, out); }
// the two Jacobi orders on the same symmetric 12 x 12 matrix (row-major): eigenvalues (descending) and eigenvectors (columns of V)
void hm_jacobi12(const double *A_in, int parallel_order, double *V_out, double *d_out) {
    double A[12][12], V[12][12], d[12];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) A[i][j] = A_in[12 * i + j];
    if (parallel_order) vo::refpnp::jacobi_eig12_rr(A, V, d);
    else vo::refpnp::jacobi_eig<12>(A, V, d);
    for (int i = 0; i < 12; ++i) {
        d_out[i] = d[i];
        for (int j = 0; j < 12; ++j) V_out[12 * i + j] = V[i][j];
    }
}
int hm_ref_epnp5(const double *X, const double *uv, const double *K, double *out) {
    double Xm[5][3], uvm[5][2];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 3; ++j) Xm[i][j] = X[3 * i + j];
        uvm[i][0] = uv[2 * i]; uvm[i][1] = uv[2 * i + 1];
    }
    vo::refpnp::Pose p;
    if (!vo::refpnp::epnp5(Xm, uvm, K[0], K[1], K[2], K[3], p)) return 0;
    for (int j = 0; j < 9; ++j) out[j] = p.R[j];
    for (int j = 0; j < 3; ++j) out[9 + j] = p.t[j];
    return 1;
}
void hm_ref_err2(const double *pose, const double *K, const float *xyz, const float *uv, int n, float *out) {
    vo::refpnp::Pose p;
    for (int j = 0; j < 9; ++j) p.R[j] = pose[j];
    for (int j = 0; j < 3; ++j) p.t[j] = pose[9 + j];
    for (int i = 0; i < n; ++i) out[i] = vo::refpnp::reproj_err2(p, K[0], K[1], K[2], K[3], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1]);
}
int hm_ref_scan(const int *counts, int n, int iters, double conf, int *iters_run, int *best_count) {
    return vo::refpnp::ransac_scan(counts, n, iters, conf, iters_run, best_count);
}
}
