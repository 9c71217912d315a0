// R2D2 front-end: network forward + heads + non-maximum suppression + descriptor gather for one image
// (R2D2.py:202-232 -> extract_keypoints :172-188 -> extract_multiscale :104-169 at scale 1 -> NonMaxSuppression
// :82-101; network feature_extractors/r2d2/nets/patchnet.py:11-186).
//
// The convolution stack runs on the tensor cores (conv_tc.cu); only the first layer (3 input channels: K = 27) runs on
// the CUDA cores, fused with the uint8 -> normalised-float conversion.  The heads never materialise the up-sampled
// 128-channel map (239 MB at 1241 x 376 in the reference): reliability / repeatability are computed per output pixel
// from the four half-resolution neighbours, and descriptors are interpolated and normalised for the surviving
// keypoints only.
#include "common.cuh"
#include <stdlib.h>
#include <vector>
#include <cuda.h>

namespace vo {
int conv_map_act(vo_ctx *ctx, void *map_out, const float *ptr, int H, int W, int C);
int conv_map_weight(vo_ctx *ctx, void *map_out, const float *ptr, int cout, int ktot);
int conv_tc_launch(vo_ctx *ctx, const void *map_a_hi, const void *map_a_lo, const void *map_b_hi, const void *map_b_lo, int H,
                   int W, int cin, int cout, int k, int dil, int pad, int relu, const float *scale, const float *shift,
                   float *out_full, float *out_hi, float *out_lo, cudaStream_t st);
}  // namespace vo

struct vo_r2d2_stage {
    int cin, cout, k, dil, pad, relu, pool_after;
    int H, W;                     // resolution the layer runs at
    float *w_hi, *w_lo;           // [cout][k*k*cin] (layer 0: w_hi holds the fp32 weights)
    float *scale, *shift;         // [cout]
    float *out_full, *out_hi, *out_lo;
    float *pool_hi, *pool_lo;     // pooled output when pool_after
    CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
};

struct vo_r2d2 {
    vo_ctx *ctx;
    int H, W, Hf, Wf, Ho, Wo, C, upsample, max_kp;  // Hf x Wf: resolution of the last layer; Ho x Wo: output maps
    std::vector<vo_r2d2_stage> st;
    std::vector<void *> allocs;
    uint8_t *rgb;                 // device staging of the input image
    float *head_w;                // clf_w0[C] | clf_w1[C] | sal_w[C] | clf_b0, clf_b1, sal_b
    float *rel, *rep;             // [Ho][Wo]
    int32_t *row_count, *row_base;  // [Ho], [Ho + 1]
};

namespace vo {
namespace {

__host__ __device__ inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest, ties away (cvt.rna.tf32.f32), finite inputs
    float r;
    memcpy(&r, &u, 4);
    return r;
}
__device__ __forceinline__ float dev_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---------------------------------------------------------------- layer 0: uint8 RGB -> normalise -> 3x3 conv (3 -> C_out)
// norm_RGB (tools/dataloader.py): ToTensor (x / 255) then (x - mean) / std with the ImageNet statistics.
template <int COUT>
__global__ void __launch_bounds__(128)
first_conv_kernel(const uint8_t *__restrict__ rgb, int H, int W, int k, int dil, int pad, const float *__restrict__ w,
                  const float *__restrict__ scale, const float *__restrict__ shift, int relu, float *__restrict__ out_hi,
                  float *__restrict__ out_lo) {
    extern __shared__ float sw[];  // [k*k*3][COUT] (transposed for broadcast-free reads) | lut[3][256]
    const int taps = k * k * 3;
    float *lut = sw + taps * COUT;  // (v / 255 - mean) / std for every byte value: the two IEEE divisions are done once
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int i = threadIdx.x; i < taps * COUT; i += blockDim.x) {
        const int t = i / COUT, co = i % COUT;
        sw[i] = w[(size_t)co * taps + t];
    }
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int ch = i >> 8;
        lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), mean[ch]), stdv[ch]);
    }
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)H * W) return;
    const int y = (int)(p / W), x = (int)(p % W);
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
    for (int ty = 0; ty < k; ++ty) {
        const int yy = y - pad + ty * dil;
        for (int tx = 0; tx < k; ++tx) {
            const int xx = x - pad + tx * dil;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;  // zero padding of the NORMALISED image
            const uint8_t *px = rgb + ((size_t)yy * W + xx) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float v = lut[ch * 256 + px[ch]];
                const float *wr = sw + ((ty * k + tx) * 3 + ch) * COUT;
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] = __fmaf_rn(v, wr[c], acc[c]);
            }
        }
    }
    float *oh = out_hi + (size_t)p * COUT, *ol = out_lo + (size_t)p * COUT;
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
        float v[4], hv[4], lv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float yv = __fmaf_rn(acc[c + j], scale[c + j], shift[c + j]);
            v[j] = relu ? fmaxf(yv, 0.0f) : yv;
            hv[j] = dev_tf32(v[j]);
            lv[j] = dev_tf32(v[j] - hv[j]);
        }
        *reinterpret_cast<float4 *>(oh + c) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4 *>(ol + c) = make_float4(lv[0], lv[1], lv[2], lv[3]);
    }
}

// ---------------------------------------------------------------- MaxPool2d(2), NHWC, output split into tf32 hi / lo
__global__ void __launch_bounds__(256)
maxpool2_kernel(const float *__restrict__ x, int H, int W, int C, float *__restrict__ out_hi, float *__restrict__ out_lo) {
    const int Hp = H / 2, Wp = W / 2, C4 = C / 4;
    const long long total = (long long)Hp * Wp * C4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c4 = (int)(i % C4);
    const long long pp = i / C4;
    const int xo = (int)(pp % Wp), yo = (int)(pp / Wp);
    const float4 *src = reinterpret_cast<const float4 *>(x);
    auto at = [&](int yy, int xx) { return src[((size_t)yy * W + xx) * C4 + c4]; };
    const float4 a = at(2 * yo, 2 * xo), b = at(2 * yo, 2 * xo + 1), c = at(2 * yo + 1, 2 * xo), d = at(2 * yo + 1, 2 * xo + 1);
    const float4 m = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                                 fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
    const float4 h = make_float4(dev_tf32(m.x), dev_tf32(m.y), dev_tf32(m.z), dev_tf32(m.w));
    reinterpret_cast<float4 *>(out_hi)[i] = h;
    reinterpret_cast<float4 *>(out_lo)[i] =
        make_float4(dev_tf32(m.x - h.x), dev_tf32(m.y - h.y), dev_tf32(m.z - h.z), dev_tf32(m.w - h.w));
}

// ---------------------------------------------------------------- bilinear sampling of the feature map (Upsample x2)
// torch upsample_bilinear2d, align_corners=False: src = (dst + 0.5) / scale - 0.5 clamped at 0; the four taps are
// combined as h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11).
struct Bilin {
    int y0, y1, x0, x1;
    float h0, h1, w0, w1;
};
__device__ __forceinline__ Bilin bilin_setup(int yo, int xo, int Hf, int Wf, int up) {
    Bilin b;
    if (up == 1) {
        b.y0 = b.y1 = yo; b.x0 = b.x1 = xo; b.h0 = 1.f; b.h1 = 0.f; b.w0 = 1.f; b.w1 = 0.f;
        return b;
    }
    const float sy = fmaxf(((float)yo + 0.5f) * 0.5f - 0.5f, 0.0f), sx = fmaxf(((float)xo + 0.5f) * 0.5f - 0.5f, 0.0f);
    b.y0 = (int)sy; b.x0 = (int)sx;
    b.y1 = b.y0 + (b.y0 < Hf - 1 ? 1 : 0);
    b.x1 = b.x0 + (b.x0 < Wf - 1 ? 1 : 0);
    b.h1 = sy - (float)b.y0; b.h0 = 1.0f - b.h1;
    b.w1 = sx - (float)b.x0; b.w0 = 1.0f - b.w1;
    return b;
}
__device__ __forceinline__ float4 bilin_sample4(const float *__restrict__ feat, int Wf, int C, const Bilin &b, int c) {
    auto at = [&](int yy, int xx) { return __ldg(reinterpret_cast<const float4 *>(feat + ((size_t)yy * Wf + xx) * C + c)); };
    const float4 v00 = at(b.y0, b.x0);
    if (b.h1 == 0.f && b.w1 == 0.f && b.h0 == 1.f && b.w0 == 1.f && b.y0 == b.y1 && b.x0 == b.x1) return v00;
    const float4 v01 = at(b.y0, b.x1), v10 = at(b.y1, b.x0), v11 = at(b.y1, b.x1);
    auto mix = [&](float a00, float a01, float a10, float a11) {
        return b.h0 * (b.w0 * a00 + b.w1 * a01) + b.h1 * (b.w0 * a10 + b.w1 * a11);
    };
    return make_float4(mix(v00.x, v01.x, v10.x, v11.x), mix(v00.y, v01.y, v10.y, v11.y), mix(v00.z, v01.z, v10.z, v11.z),
                       mix(v00.w, v01.w, v10.w, v11.w));
}

// ---------------------------------------------------------------- heads
// reliability = softmax(clf(x^2))[1], repeatability = softplus(sal(x^2)) / (1 + softplus) (patchnet.py:16-22, :181-186).
// One warp per block of output pixels that interpolate the SAME four feature vectors: with x2 up-sampling the output
// pixels 2k-1 and 2k both read half-resolution columns (k-1, k), so a 2 x 2 output block shares its 2 x 2 neighbours
// and every feature vector is loaded once per block instead of once per pixel.  4 channels per lane (C = 128 k).
__device__ __forceinline__ void head_finish(float s0, float s1, float s2, const float *__restrict__ hw, int C, float *rel, float *rep,
                                            size_t p) {
    const float u0 = s0 + hw[3 * C], u1 = s1 + hw[3 * C + 1], us = s2 + hw[3 * C + 2];
    const float m = fmaxf(u0, u1);
    const float e0 = expf(u0 - m), e1 = expf(u1 - m);
    rel[p] = e1 / (e0 + e1);
    const float sp = us > 20.0f ? us : log1pf(expf(us));  // torch softplus, threshold 20
    rep[p] = sp / (1.0f + sp);
}

__global__ void __launch_bounds__(256)
head_maps_kernel(const float *__restrict__ feat, int Hf, int Wf, int C, int up, int Ho, int Wo, const float *__restrict__ hw,
                 float *__restrict__ rel, float *__restrict__ rep) {
    const long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int bw = up == 2 ? Wf + 1 : Wo, bh = up == 2 ? Hf + 1 : Ho;  // blocks per row / column
    if (blk >= (long long)bw * bh) return;
    const int by = (int)(blk / bw), bx = (int)(blk % bw);
    // output pixels of the block: rows {2by-1, 2by}, columns {2bx-1, 2bx} (those inside the map); up == 1: one pixel
    const int y_lo = up == 2 ? max(2 * by - 1, 0) : by, y_hi = up == 2 ? min(2 * by, Ho - 1) : by;
    const int x_lo = up == 2 ? max(2 * bx - 1, 0) : bx, x_hi = up == 2 ? min(2 * bx, Wo - 1) : bx;
    const Bilin b0 = bilin_setup(y_lo, x_lo, Hf, Wf, up);  // indices shared by the whole block
    for (int c = lane * 4; c < C; c += 128) {
        auto at = [&](int yy, int xx) { return __ldg(reinterpret_cast<const float4 *>(feat + ((size_t)yy * Wf + xx) * C + c)); };
        const float4 v00 = at(b0.y0, b0.x0), v01 = at(b0.y0, b0.x1), v10 = at(b0.y1, b0.x0), v11 = at(b0.y1, b0.x1);
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(hw + c));
        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(hw + C + c));
        const float4 a2 = __ldg(reinterpret_cast<const float4 *>(hw + 2 * C + c));
        for (int yo = y_lo; yo <= y_hi; ++yo)
            for (int xo = x_lo; xo <= x_hi; ++xo) {
                const Bilin b = bilin_setup(yo, xo, Hf, Wf, up);
                auto mix = [&](float p00, float p01, float p10, float p11) {
                    return b.h0 * (b.w0 * p00 + b.w1 * p01) + b.h1 * (b.w0 * p10 + b.w1 * p11);
                };
                const float4 v = make_float4(mix(v00.x, v01.x, v10.x, v11.x), mix(v00.y, v01.y, v10.y, v11.y),
                                             mix(v00.z, v01.z, v10.z, v11.z), mix(v00.w, v01.w, v10.w, v11.w));
                const float4 q = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
                float s0 = q.x * a0.x + q.y * a0.y + q.z * a0.z + q.w * a0.w;
                float s1 = q.x * a1.x + q.y * a1.y + q.z * a1.z + q.w * a1.w;
                float s2 = q.x * a2.x + q.y * a2.y + q.z * a2.z + q.w * a2.w;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                // C == 128: one pass per pixel, lane 0 finishes.  (C > 128 would need the partial sums carried across
                // the channel loop; vo_r2d2_create restricts the head to C == 128.)
                if (lane == 0) head_finish(s0, s1, s2, hw, C, rel, rep, (size_t)yo * Wo + xo);
            }
    }
}

// ---------------------------------------------------------------- NMS (3x3 max, -inf padding) + thresholds, ordered
__device__ __forceinline__ bool is_keypoint(const float *__restrict__ rel, const float *__restrict__ rep, int Ho, int Wo, int y,
                                            int x, float rel_thr, float rep_thr, float score_thr) {
    const float r = rep[(size_t)y * Wo + x];
    if (!(r >= rep_thr)) return false;
    const float c = rel[(size_t)y * Wo + x];
    if (!(c >= rel_thr)) return false;
    if (!(c * r > score_thr)) return false;
    float m = r;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < Ho && xx >= 0 && xx < Wo) m = fmaxf(m, rep[(size_t)yy * Wo + xx]);
        }
    return r == m;
}

__global__ void __launch_bounds__(256)
nms_count_kernel(const float *__restrict__ rel, const float *__restrict__ rep, int Ho, int Wo, float rel_thr, float rep_thr,
                 float score_thr, int32_t *__restrict__ row_count) {
    const int y = blockIdx.x;
    int cnt = 0;
    for (int x = threadIdx.x; x < Wo; x += blockDim.x) cnt += is_keypoint(rel, rep, Ho, Wo, y, x, rel_thr, rep_thr, score_thr) ? 1 : 0;
    __shared__ int s[8];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        row_count[y] = t;
    }
}

__global__ void __launch_bounds__(1024)
row_scan_kernel(const int32_t *__restrict__ row_count, int Ho, int32_t *__restrict__ row_base, int32_t *__restrict__ total) {
    // exclusive scan of <= a few thousand row counts by one CTA (chunks of 1024 with a running carry)
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int r0 = 0; r0 < Ho; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        const int v = r < Ho ? row_count[r] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (r < Ho) row_base[r] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) { row_base[Ho] = carry; *total = carry; }
}

// one CTA per row: ordered write of (x, y, 32) and the score (ballot within a warp, warp counts scanned per chunk)
constexpr int NW_THREADS = 256;
__global__ void __launch_bounds__(NW_THREADS)
nms_write_kernel(const float *__restrict__ rel, const float *__restrict__ rep, int Ho, int Wo, float rel_thr, float rep_thr,
                 float score_thr, const int32_t *__restrict__ row_base, int max_kp, float *__restrict__ xys,
                 float *__restrict__ scores) {
    const int y = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int wcnt[NW_THREADS / 32];
    int base = row_base[y];
    if (row_base[y + 1] == base) return;
    for (int x0 = 0; x0 < Wo; x0 += NW_THREADS) {
        const int x = x0 + threadIdx.x;
        const bool kp = x < Wo && is_keypoint(rel, rep, Ho, Wo, y, x, rel_thr, rep_thr, score_thr);
        const unsigned bal = __ballot_sync(0xffffffffu, kp);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int prefix = base, total = 0;
        for (int w = 0; w < NW_THREADS / 32; ++w) {
            if (w < warp) prefix += wcnt[w];
            total += wcnt[w];
        }
        if (kp) {
            const int idx = prefix + __popc(bal & ((1u << lane) - 1u));
            if (idx < max_kp) {
                xys[(size_t)idx * 3 + 0] = (float)x;   // X = x * W / nw with nw == W at scale 1 (R2D2.py:150)
                xys[(size_t)idx * 3 + 1] = (float)y;
                xys[(size_t)idx * 3 + 2] = 32.0f;      // 32 / s
                scores[idx] = rel[(size_t)y * Wo + x] * rep[(size_t)y * Wo + x];
            }
        }
        base += total;
        __syncthreads();
    }
}

// one warp per keypoint: interpolate the feature vector at its pixel, L2-normalise (F.normalize, eps 1e-12)
__global__ void __launch_bounds__(256)
desc_gather_kernel(const float *__restrict__ xys, const int32_t *__restrict__ total, int max_kp, const float *__restrict__ feat,
                   int Hf, int Wf, int C, int up, float *__restrict__ desc) {
    const int idx = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (idx >= min(*total, max_kp)) return;
    const int x = (int)xys[(size_t)idx * 3 + 0], y = (int)xys[(size_t)idx * 3 + 1];
    const Bilin b = bilin_setup(y, x, Hf, Wf, up);
    const float4 v = bilin_sample4(feat, Wf, C, b, lane * 4);
    float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    *reinterpret_cast<float4 *>(desc + (size_t)idx * C + lane * 4) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
}

void *dev_alloc(vo_r2d2 *n, size_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
    n->allocs.push_back(p);
    return p;
}

}  // namespace
}  // namespace vo

extern "C" void vo_r2d2_destroy(vo_r2d2 *net) {
    if (!net) return;
    cudaDeviceSynchronize();
    for (void *p : net->allocs) cudaFree(p);
    delete net;
}

extern "C" int vo_r2d2_out_shape(const vo_r2d2 *net, int *Ho, int *Wo) {
    using namespace vo;
    VO_REQUIRE(net && Ho && Wo, "vo_r2d2_out_shape: null argument");
    *Ho = net->Ho; *Wo = net->Wo;
    return VO_OK;
}

extern "C" int vo_r2d2_create(vo_ctx *ctx, const vo_r2d2_config *cfg, vo_r2d2 **out) {
    using namespace vo;
    VO_REQUIRE(ctx && cfg && out, "vo_r2d2_create: null argument");
    *out = nullptr;
    VO_REQUIRE(cfg->H > 0 && cfg->W > 0 && cfg->n_layers >= 2 && cfg->layers && cfg->max_kp > 0, "vo_r2d2_create: bad config");
    VO_REQUIRE(cfg->upsample == 1 || cfg->upsample == 2, "vo_r2d2_create: upsample must be 1 or 2");
    VO_REQUIRE(cfg->layers[0].cin == 3 && (cfg->layers[0].cout == 32 || cfg->layers[0].cout == 64),
               "vo_r2d2_create: the first layer must map 3 channels to 32 or 64");
    VO_REQUIRE(cfg->layers[0].pool_after == 0, "vo_r2d2_create: pooling directly after the first layer is not supported");
    VO_REQUIRE(cfg->clf_w && cfg->clf_b && cfg->sal_w && cfg->sal_b, "vo_r2d2_create: null head weights");
    vo_r2d2 *n = new vo_r2d2();
    n->ctx = ctx; n->H = cfg->H; n->W = cfg->W; n->upsample = cfg->upsample; n->max_kp = cfg->max_kp;
    int H = cfg->H, W = cfg->W, rc = VO_OK;
    const float *prev_hi = nullptr, *prev_lo = nullptr;
    auto fail = [&](int code) { vo_r2d2_destroy(n); return code; };
    for (int li = 0; li < cfg->n_layers; ++li) {
        const vo_r2d2_layer &L = cfg->layers[li];
        if (!(L.w && L.bias && (!L.bn || (L.bn_mean && L.bn_var)))) { set_error("vo_r2d2_create: layer %d has null weights", li); return fail(VO_ERR_ARG); }
        if (li > 0 && L.cin != cfg->layers[li - 1].cout) { set_error("vo_r2d2_create: layer %d C_in %d != previous C_out", li, L.cin); return fail(VO_ERR_ARG); }
        if (li > 0 && (L.cin % 32 || !(L.cout == 32 || L.cout == 64 || L.cout == 128))) { set_error("vo_r2d2_create: layer %d: unsupported channel counts %d -> %d", li, L.cin, L.cout); return fail(VO_ERR_ARG); }
        vo_r2d2_stage s{};
        s.cin = L.cin; s.cout = L.cout; s.k = L.k; s.dil = L.dil; s.pad = ((L.k - 1) * L.dil) / 2; s.relu = L.relu;
        s.pool_after = L.pool_after; s.H = H; s.W = W;
        const size_t nw = (size_t)L.cout * L.k * L.k * L.cin, npx = (size_t)H * W;
        std::vector<float> hi(nw), lo(nw), sc(L.cout), sh(L.cout);
        for (size_t i = 0; i < nw; ++i) {
            if (li == 0) { hi[i] = L.w[i]; lo[i] = 0.f; }
            else { hi[i] = tf32_rna_host(L.w[i]); lo[i] = tf32_rna_host(L.w[i] - hi[i]); }
        }
        for (int c = 0; c < L.cout; ++c) {  // y = (conv + bias - mean) / sqrt(var + eps)
            const float inv = L.bn ? 1.0f / sqrtf(L.bn_var[c] + cfg->bn_eps) : 1.0f;
            sc[c] = inv;
            sh[c] = L.bn ? (L.bias[c] - L.bn_mean[c]) * inv : L.bias[c];
        }
        s.w_hi = (float *)dev_alloc(n, nw * 4); s.w_lo = (float *)dev_alloc(n, nw * 4);
        s.scale = (float *)dev_alloc(n, L.cout * 4); s.shift = (float *)dev_alloc(n, L.cout * 4);
        const bool last = li == cfg->n_layers - 1;
        const bool need_full = last || L.pool_after;
        if (need_full) s.out_full = (float *)dev_alloc(n, npx * L.cout * 4);
        else { s.out_hi = (float *)dev_alloc(n, npx * L.cout * 4); s.out_lo = (float *)dev_alloc(n, npx * L.cout * 4); }
        if (L.pool_after) {
            if (L.pool_after != 2 || last) { set_error("vo_r2d2_create: layer %d: only MaxPool2d(2) between layers is supported", li); return fail(VO_ERR_ARG); }
            const size_t pp = (size_t)(H / 2) * (W / 2) * L.cout;
            s.pool_hi = (float *)dev_alloc(n, pp * 4); s.pool_lo = (float *)dev_alloc(n, pp * 4);
        }
        if (!s.w_hi || !s.w_lo || !s.scale || !s.shift || (need_full && !s.out_full) || (!need_full && (!s.out_hi || !s.out_lo)) ||
            (L.pool_after && (!s.pool_hi || !s.pool_lo))) { set_error("vo_r2d2_create: out of device memory"); return fail(VO_ERR_CUDA); }
        cudaMemcpy(s.w_hi, hi.data(), nw * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(s.w_lo, lo.data(), nw * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(s.scale, sc.data(), L.cout * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(s.shift, sh.data(), L.cout * 4, cudaMemcpyHostToDevice);
        if (li > 0) {
            if ((rc = conv_map_act(ctx, &s.map_a_hi, prev_hi, H, W, L.cin))) return fail(rc);
            if ((rc = conv_map_act(ctx, &s.map_a_lo, prev_lo, H, W, L.cin))) return fail(rc);
            if ((rc = conv_map_weight(ctx, &s.map_b_hi, s.w_hi, L.cout, L.k * L.k * L.cin))) return fail(rc);
            if ((rc = conv_map_weight(ctx, &s.map_b_lo, s.w_lo, L.cout, L.k * L.k * L.cin))) return fail(rc);
        }
        if (L.pool_after) { prev_hi = s.pool_hi; prev_lo = s.pool_lo; H /= 2; W /= 2; }
        else { prev_hi = s.out_hi; prev_lo = s.out_lo; }
        n->st.push_back(s);
    }
    n->Hf = H; n->Wf = W; n->C = cfg->layers[cfg->n_layers - 1].cout;
    n->Ho = H * cfg->upsample; n->Wo = W * cfg->upsample;
    if (n->C != 128) { set_error("vo_r2d2_create: the descriptor length must be 128 (got %d)", n->C); return fail(VO_ERR_ARG); }
    const int C = n->C;
    std::vector<float> hw(3 * C + 4, 0.f);
    for (int c = 0; c < C; ++c) { hw[c] = cfg->clf_w[c]; hw[C + c] = cfg->clf_w[C + c]; hw[2 * C + c] = cfg->sal_w[c]; }
    hw[3 * C] = cfg->clf_b[0]; hw[3 * C + 1] = cfg->clf_b[1]; hw[3 * C + 2] = cfg->sal_b[0];
    n->head_w = (float *)dev_alloc(n, hw.size() * 4);
    n->rgb = (uint8_t *)dev_alloc(n, (size_t)cfg->H * cfg->W * 3);
    n->rel = (float *)dev_alloc(n, (size_t)n->Ho * n->Wo * 4);
    n->rep = (float *)dev_alloc(n, (size_t)n->Ho * n->Wo * 4);
    n->row_count = (int32_t *)dev_alloc(n, (size_t)n->Ho * 4);
    n->row_base = (int32_t *)dev_alloc(n, (size_t)(n->Ho + 1) * 4);
    if (!n->head_w || !n->rgb || !n->rel || !n->rep || !n->row_count || !n->row_base) { set_error("vo_r2d2_create: out of device memory"); return fail(VO_ERR_CUDA); }
    cudaMemcpy(n->head_w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("vo_r2d2_create: %s", cudaGetErrorString(e)); return fail(VO_ERR_CUDA); }
    *out = n;
    return VO_OK;
}

extern "C" int vo_r2d2_extract(vo_r2d2 *net, const uint8_t *rgb, float rel_thr, float rep_thr, float score_thr, float *xys,
                               float *desc, float *scores, int32_t *count, float *rel_map, float *rep_map, void *stream) {
    using namespace vo;
    VO_REQUIRE(net && rgb && xys && desc && scores && count, "vo_r2d2_extract: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    vo_ctx *ctx = net->ctx;
    int rc;
    VO_CUDA(cudaMemcpyAsync(net->rgb, rgb, (size_t)net->H * net->W * 3, cudaMemcpyDefault, st));
    for (size_t li = 0; li < net->st.size(); ++li) {
        vo_r2d2_stage &s = net->st[li];
        if (li == 0) {
            const long long px = (long long)s.H * s.W;
            const size_t smem = ((size_t)s.k * s.k * 3 * s.cout + 768) * sizeof(float);
            if (s.cout == 32)
                first_conv_kernel<32><<<(unsigned)((px + 127) / 128), 128, smem, st>>>(net->rgb, s.H, s.W, s.k, s.dil, s.pad, s.w_hi, s.scale,
                                                                                        s.shift, s.relu, s.out_hi, s.out_lo);
            else
                first_conv_kernel<64><<<(unsigned)((px + 127) / 128), 128, smem, st>>>(net->rgb, s.H, s.W, s.k, s.dil, s.pad, s.w_hi, s.scale,
                                                                                        s.shift, s.relu, s.out_hi, s.out_lo);
            VO_LAUNCH_CHECK(ctx);
        } else {
            if ((rc = conv_tc_launch(ctx, &s.map_a_hi, &s.map_a_lo, &s.map_b_hi, &s.map_b_lo, s.H, s.W, s.cin, s.cout, s.k, s.dil, s.pad,
                                     s.relu, s.scale, s.shift, s.out_full, s.out_hi, s.out_lo, st)))
                return rc;
        }
        if (s.pool_after) {
            const long long total = (long long)(s.H / 2) * (s.W / 2) * (s.cout / 4);
            maxpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s.out_full, s.H, s.W, s.cout, s.pool_hi, s.pool_lo);
            VO_LAUNCH_CHECK(ctx);
        }
    }
    const float *feat = net->st.back().out_full;
    float *rel = rel_map ? rel_map : net->rel, *rep = rep_map ? rep_map : net->rep;
    const long long warps = net->upsample == 2 ? (long long)(net->Hf + 1) * (net->Wf + 1) : (long long)net->Ho * net->Wo;
    head_maps_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(feat, net->Hf, net->Wf, net->C, net->upsample, net->Ho, net->Wo,
                                                                            net->head_w, rel, rep);
    VO_LAUNCH_CHECK(ctx);
    nms_count_kernel<<<net->Ho, 256, 0, st>>>(rel, rep, net->Ho, net->Wo, rel_thr, rep_thr, score_thr, net->row_count);
    VO_LAUNCH_CHECK(ctx);
    row_scan_kernel<<<1, 1024, 0, st>>>(net->row_count, net->Ho, net->row_base, count);
    VO_LAUNCH_CHECK(ctx);
    nms_write_kernel<<<net->Ho, NW_THREADS, 0, st>>>(rel, rep, net->Ho, net->Wo, rel_thr, rep_thr, score_thr, net->row_base, net->max_kp, xys,
                                             scores);
    VO_LAUNCH_CHECK(ctx);
    // the keypoint count lives on the device: launch for the capacity, surplus warps exit on the count
    desc_gather_kernel<<<(unsigned)(((long long)net->max_kp * 32 + 255) / 256), 256, 0, st>>>(xys, count, net->max_kp, feat, net->Hf,
                                                                                               net->Wf, net->C, net->upsample, desc);
    VO_LAUNCH_CHECK(ctx);
    return VO_OK;
}
