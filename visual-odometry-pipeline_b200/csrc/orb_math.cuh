// Per-pixel / per-keypoint arithmetic of the ORB front-end (feature_extractors/ORB.py:8-21 -> cv2.ORB_create()
// .detectAndCompute), written once for device code and compilable for the host: tests/test_host_math.py checks every
// function bit for bit against the CPU restatement of the front-end, which is pinned against OpenCV.  The kernels that call
// these (pyramid, FAST map, retainBest, Harris, orientation, Gaussian, rBRIEF) are csrc/orb.cu.
//
// Exactness rules: integer work is exact; fp32 work spells out every rounding (no contraction on either side: the
// device uses __f*_rn intrinsics, the host build uses -ffp-contract=off); the two fused operations OpenCV itself
// performs (the row pass of its float separable filter) are explicit fmaf.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VO_ORB_HD __host__ __device__ __forceinline__
#else
#define VO_ORB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define VO_ORB_MUL(a, b) __fmul_rn((a), (b))
#define VO_ORB_ADD(a, b) __fadd_rn((a), (b))
#define VO_ORB_SUB(a, b) __fsub_rn((a), (b))
#define VO_ORB_DIV(a, b) __fdiv_rn((a), (b))
#define VO_ORB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define VO_ORB_RINT(x) __float2int_rn(x)
#else
#define VO_ORB_MUL(a, b) ((a) * (b))
#define VO_ORB_ADD(a, b) ((a) + (b))
#define VO_ORB_SUB(a, b) ((a) - (b))
#define VO_ORB_DIV(a, b) ((a) / (b))
#define VO_ORB_FMA(a, b, c) fmaf((a), (b), (c))
#define VO_ORB_RINT(x) ((int)lrintf(x))
#endif

namespace vo {
namespace orb {

constexpr int BORDER = 32;        // max(edgeThreshold 31, ceil(15 sqrt 2), HARRIS_BLOCK_SIZE / 2) + 1
constexpr int HALF_PATCH = 15;
// Row half-width of the circular orientation patch at row offset v = 0..15 (orb.cpp computeKeyPoints, umax):
// {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3}, one nibble each; checked against the CPU restatement.
VO_ORB_HD int umax(int v) { return (int)((0x3689abcddeeeffffull >> (4 * v)) & 15ull); }

// cv2.cvtColor(BGR2GRAY) on 8-bit pixels.
VO_ORB_HD uint8_t bgr_to_gray(uint8_t b, uint8_t g, uint8_t r) {
    return (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
}

// INTER_LINEAR_EXACT: source offset and 8.8 fixed-point weight of the right / lower neighbour for destination index x
// of a dst-long axis resized from src (resize.cpp interpolationLinear<ufixedpoint16>::getCoeffs).  `inside` false:
// the destination pixel copies source pixel `ofs` (clamped edge).
VO_ORB_HD void linear_exact_coeff(int x, int dst, int src, int &ofs, int &c1, bool &inside) {
    const double scale = 1.0 / ((double)dst / (double)src);
#if defined(__CUDA_ARCH__)   // OpenCV's softdouble rounds the product and the difference separately: no contraction
    const double fval = __dsub_rn(__dmul_rn(scale, (double)x + 0.5), 0.5);
#else
    const double fval = scale * ((double)x + 0.5) - 0.5;
#endif
    const int iv = (int)floor(fval);
    inside = iv >= 0 && iv < src - 1 && src > 1;
    if (inside) {
        ofs = iv;
        c1 = (int)rint((fval - (double)iv) * 256.0);
    } else {
        ofs = (iv >= 0 && src > 1) ? src - 1 : 0;
        c1 = 0;
    }
}
// One destination pixel from its four source neighbours (p00 p01 / p10 p11), weights cx, cy in 8.8.
VO_ORB_HD uint8_t linear_exact_pixel(int p00, int p01, int p10, int p11, int cx, bool in_x, int cy, bool in_y) {
    const int h0 = in_x ? (256 - cx) * p00 + cx * p01 : p00 * 256;
    const int h1 = in_x ? (256 - cx) * p10 + cx * p11 : p10 * 256;
    const int v = in_y ? (256 - cy) * h0 + cy * h1 : h0 * 256;
    return (uint8_t)((v + (1 << 15)) >> 16);
}

// FAST-9/16 corner score of a pixel with value v and ring values ring[0..15] (clockwise from (0, +3)): 0 if no arc of
// nine contiguous ring pixels is uniformly brighter than v + thr or darker than v - thr, else the largest threshold
// for which the pixel stays a corner (fast_score.cpp cornerScore<16>): max over the 16 arcs of min |v - p|, minus 1.
VO_ORB_HD int fast_corner_score(int v, const uint8_t (&ring)[16], int thr) {
    // Sliding-window form: lo9[k] / hi9[k] = min / max of d over the arc k..k+8, built from windows of 2, 4, 8.  The
    // darker-arc score is max_k lo9[k]; the brighter-arc score is max_k min(p - v) = -(min_k hi9[k]): ONE negation at
    // the end.  (The direct form, max(lo, -hi) per arc, is miscompiled by ptxas 12.9 for sm_100a: it folds 15 of the
    // 16 negations away when it fuses the min / max chains into VIMNMX3 — found with tools/orb_bisect.py on a B200;
    // tools/probe/fast_probe.cu keeps the reproducer.)
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)ring[k];
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a = d[k], b = d[(k + 1) & 15];
        lo2[k] = a < b ? a : b;
        hi2[k] = a > b ? a : b;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo4[k] = lo2[k] < lo2[(k + 2) & 15] ? lo2[k] : lo2[(k + 2) & 15];
        hi4[k] = hi2[k] > hi2[(k + 2) & 15] ? hi2[k] : hi2[(k + 2) & 15];
    }
    int dark = -256, nbright = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int lo = lo4[k] < lo4[(k + 4) & 15] ? lo4[k] : lo4[(k + 4) & 15];
        int hi = hi4[k] > hi4[(k + 4) & 15] ? hi4[k] : hi4[(k + 4) & 15];
        lo = lo < d[(k + 8) & 15] ? lo : d[(k + 8) & 15];
        hi = hi > d[(k + 8) & 15] ? hi : d[(k + 8) & 15];
        dark = lo > dark ? lo : dark;
        nbright = hi < nbright ? hi : nbright;
    }
    const int bright = 0 - nbright;
    const int best = dark > bright ? dark : bright;
    return best > thr ? best - 1 : 0;
}

// orb.cpp HarrisResponses from the integer sums a = sum Ix^2, b = sum Iy^2, c = sum Ix Iy over the 7 x 7 block.
VO_ORB_HD float harris_response(int a, int b, int c) {
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float scale = VO_ORB_DIV(1.0f, VO_ORB_MUL((float)((1 << 2) * 7), 255.0f));
    const float s2 = VO_ORB_MUL(scale, scale);
    const float s4 = VO_ORB_MUL(VO_ORB_MUL(s2, scale), scale);
    const float sum = VO_ORB_ADD(fa, fb);
    const float det = VO_ORB_SUB(VO_ORB_MUL(fa, fb), VO_ORB_MUL(fc, fc));
    const float tr = VO_ORB_MUL(VO_ORB_MUL(0.04f, sum), sum);
    return VO_ORB_MUL(VO_ORB_SUB(det, tr), s4);
}

// cv::fastAtan2(y, x) in degrees (mathfuncs_core.simd.hpp, scalar path).
VO_ORB_HD float fast_atan2(float y, float x) {
    const float p1 = VO_ORB_MUL(0.9997878412794807f, (float)(180 / 3.141592653589793238462643383279502884));
    const float p3 = VO_ORB_MUL(-0.3258083974640975f, (float)(180 / 3.141592653589793238462643383279502884));
    const float p5 = VO_ORB_MUL(0.1555786518463281f, (float)(180 / 3.141592653589793238462643383279502884));
    const float p7 = VO_ORB_MUL(-0.04432655554792128f, (float)(180 / 3.141592653589793238462643383279502884));
    const float ax = fabsf(x), ay = fabsf(y), eps = (float)2.2204460492503131e-16;
    const bool wide = ax >= ay;
    const float c = wide ? VO_ORB_DIV(ay, VO_ORB_ADD(ax, eps)) : VO_ORB_DIV(ax, VO_ORB_ADD(ay, eps));
    const float c2 = VO_ORB_MUL(c, c);
    float a = VO_ORB_ADD(VO_ORB_MUL(p7, c2), p5);
    a = VO_ORB_ADD(VO_ORB_MUL(a, c2), p3);
    a = VO_ORB_ADD(VO_ORB_MUL(a, c2), p1);
    a = VO_ORB_MUL(a, c);
    if (!wide) a = VO_ORB_SUB(90.0f, a);
    if (x < 0) a = VO_ORB_SUB(180.0f, a);
    if (y < 0) a = VO_ORB_SUB(360.0f, a);
    return a;
}

// 7-tap Gaussian, sigma 2 (cv2.getGaussianKernel(7, 2, CV_32F)).
VO_ORB_HD void gaussian_kernel(float (&k)[7]) {
    double e[7], s = 0.0;
    for (int i = 0; i < 7; ++i) { const double x = (double)(i - 3); e[i] = exp(-(x * x) / 8.0); s += e[i]; }
    for (int i = 0; i < 7; ++i) k[i] = (float)(e[i] / s);
}
// Row pass of OpenCV's float separable filter on 8-bit pixels p[0..6]: s = k0 p0, then fused multiply-adds left to right.
VO_ORB_HD float blur_row(const float (&k)[7], const uint8_t (&p)[7]) {
    float s = VO_ORB_MUL((float)p[0], k[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) s = VO_ORB_FMA((float)p[i], k[i], s);
    return s;
}
// Column pass on the row results h[0..6]: centre first, then the symmetric pairs added before the multiply; the result
// is rounded half to even and saturated.
VO_ORB_HD uint8_t blur_col(const float (&k)[7], const float (&h)[7]) {
    float s = VO_ORB_MUL(h[3], k[3]);
#pragma unroll
    for (int j = 1; j <= 3; ++j) s = VO_ORB_ADD(s, VO_ORB_MUL(VO_ORB_ADD(h[3 + j], h[3 - j]), k[3 + j]));
    const int v = VO_ORB_RINT(s);
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// rBRIEF: pattern point (px, py) rotated by (ca, sb) = (cos, sin) of the keypoint angle, rounded half to even
// (orb.cpp GET_VALUE).  ca / sb are (float)cos / (float)sin of the fp32 angle in radians, evaluated in double.
VO_ORB_HD void angle_cos_sin(float angle_deg, float &ca, float &sb) {
    const float rad = VO_ORB_MUL(angle_deg, (float)(3.141592653589793238462643383279502884 / 180.0));
    ca = (float)cos((double)rad);
    sb = (float)sin((double)rad);
}
VO_ORB_HD void rotate_pattern_point(int px, int py, float ca, float sb, int &ix, int &iy) {
    const float fx = (float)px, fy = (float)py;
    ix = VO_ORB_RINT(VO_ORB_SUB(VO_ORB_MUL(fx, ca), VO_ORB_MUL(fy, sb)));
    iy = VO_ORB_RINT(VO_ORB_ADD(VO_ORB_MUL(fx, sb), VO_ORB_MUL(fy, ca)));
}

}  // namespace orb
}  // namespace vo
