// Depth back-projection (dense) and the fused sparse path: keypoint gather + min-flow filter +
// depth lookup at truncated pixel + range gate + order-preserving compaction.
//
// Replaces cv2.rgbd.depthTo3d (VisualOdometry_Stereo.py:96), the fancy-index gather at :97,
// the 0<Z<50 gate at :100-105 and the keypoint gather / 3 px flow filter at :257-264.
// Arithmetic follows depthTo3d for float depth: intrinsics are cast to the depth dtype (fp32),
// X = ((u - cx) * (1/fx)) * z, Y = ((v - cy) * (1/fy)) * z, Z = z, each operation rounded to
// fp32 (explicit _rn intrinsics so the compiler cannot contract them into FMAs).
#include "common.cuh"

namespace vo {
namespace {

struct Intr {
    float inv_fx, inv_fy, cx, cy;
};

__host__ Intr make_intr(const double *K) {
    Intr k;
    const float fx = (float)K[0], fy = (float)K[4];
    k.cx = (float)K[2];
    k.cy = (float)K[5];
    k.inv_fx = 1.0f / fx;
    k.inv_fy = 1.0f / fy;
    return k;
}

__device__ __forceinline__ void backproject(const Intr k, int u, int v, float z, float &X, float &Y) {
    X = __fmul_rn(__fmul_rn(__fsub_rn((float)u, k.cx), k.inv_fx), z);
    Y = __fmul_rn(__fmul_rn(__fsub_rn((float)v, k.cy), k.inv_fy), z);
}

// Dense: one thread owns 4 consecutive pixels of the flattened [B*H*W] image (one 128-bit streaming load).  Its 12
// outputs go through a warp-private shared-memory tile so that every 128-bit streaming store of a warp covers 512
// contiguous bytes (a direct store would put each lane's 16 B at a 48 B stride: half-filled sectors on every request).
// HBM-bound: 16 B per pixel.
constexpr int BD_THREADS = 256;

__global__ void __launch_bounds__(BD_THREADS)
backproject_dense_kernel(const float *__restrict__ depth, float *__restrict__ xyz, long long n_px, int H, int W,
                         Intr k) {
    __shared__ float4 tile[BD_THREADS / 32][96];  // per warp: 128 pixels x 3 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_vec = n_px >> 2;
    const long long n_vec_warp = (n_vec + 31) & ~31ll;  // whole warps iterate together (shared-memory exchange)
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int UNROLL = 4;  // loads of UNROLL grid-strided vectors are issued before the first is consumed
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec_warp; i0 += UNROLL * stride) {
        float4 zin[UNROLL];
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
            const long long i = i0 + r * stride;
            zin[r] = (i < n_vec) ? __ldcs(reinterpret_cast<const float4 *>(depth) + i) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
            const long long i = i0 + r * stride;
            const long long w0 = i - lane;  // first vector of this warp's 32
            if (w0 >= n_vec) break;         // warp-uniform
            if (i < n_vec) {
                const float z[4] = {zin[r].x, zin[r].y, zin[r].z, zin[r].w};
                float o[12];
                const long long p0 = i << 2;
                int u = (int)(p0 % W);
                int v = (int)((p0 / W) % H);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    backproject(k, u, v, z[q], o[3 * q], o[3 * q + 1]);
                    o[3 * q + 2] = z[q];
                    if (++u == W) { u = 0; if (++v == H) v = 0; }
                }
                tile[warp][3 * lane + 0] = make_float4(o[0], o[1], o[2], o[3]);
                tile[warp][3 * lane + 1] = make_float4(o[4], o[5], o[6], o[7]);
                tile[warp][3 * lane + 2] = make_float4(o[8], o[9], o[10], o[11]);
            }
            __syncwarp();
            const long long valid = min(32ll, n_vec - w0) * 3;  // float4 outputs this warp produced
            float4 *dst = reinterpret_cast<float4 *>(xyz) + w0 * 3;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int j = q * 32 + lane;
                if (j < valid) __stcs(dst + j, tile[warp][j]);
            }
            __syncwarp();
        }
    }
    // tail (n_px not a multiple of 4)
    if (blockIdx.x == 0 && threadIdx.x < (n_px & 3)) {
        const long long p = (n_vec << 2) + threadIdx.x;
        const int u = (int)(p % W), v = (int)((p / W) % H);
        const float z = depth[p];
        float X, Y;
        backproject(k, u, v, z, X, Y);
        xyz[p * 3 + 0] = X; xyz[p * 3 + 1] = Y; xyz[p * 3 + 2] = z;
    }
}

constexpr int GB_THREADS = 512;
constexpr uint32_t VO_DEPTH_OOB_BITS = 0xffc0b200u;  // quiet NaN with a payload: "keypoint truncates outside the image"

// Depth at every keypoint's truncated pixel, depth[int(y), int(x)] (VisualOdometry_Stereo.py:97), as a compact
// [B][n_stride] array.  `depth` only has to be device-ACCESSIBLE: reading a pinned host image through the mapped
// pointer moves one 32-byte sector per keypoint over PCIe instead of the whole H x W map (1.87 MB at 1241 x 376).
// Thin and persistent on purpose: a zero-copy read takes ~2 us of link latency and the link sustains only a few hundred
// reads in flight (tools/h2d_wall.py: ~366 M sector reads/s), so a few thousand resident threads with SD_ILP independent loads
// each saturate it.  64-thread CTAs with a small register footprint co-reside with the CTAs of a matcher launch that is
// running on another stream (which leaves only a few thousand free registers per SM), instead of queueing behind them:
// the e2e path overlaps the sampling of the next chunk with the matcher of the current one.
// <= 32 registers per thread: three 80-register matcher CTAs leave 1024 registers per SM sub-partition, one such warp.
constexpr int SD_THREADS = 64, SD_ILP = 4;
__global__ void __launch_bounds__(SD_THREADS, 32)
sample_depth_kernel(const float *__restrict__ kp, int n_stride, int kp_stride, const int32_t *__restrict__ n_kp,
                    const float *__restrict__ depth, int H, int W, long long total, float *__restrict__ z_kp) {
    const long long step = (long long)gridDim.x * SD_THREADS * SD_ILP;
    for (long long base = (long long)blockIdx.x * SD_THREADS * SD_ILP + threadIdx.x; base < total; base += step) {
        float z[SD_ILP];
        const float *src[SD_ILP];
#pragma unroll
        for (int q = 0; q < SD_ILP; ++q) {
            const long long gid = base + (long long)q * SD_THREADS;
            src[q] = nullptr;
            z[q] = __int_as_float(0x7fc00000);
            if (gid < total) {
                const int b = (int)(gid / n_stride), i = (int)(gid % n_stride);
                const int n = n_kp ? min(n_kp[b], n_stride) : n_stride;
                if (i < n) {
                    const float *p = kp + (size_t)gid * kp_stride;
                    const int u = (int)p[0], v = (int)p[1];
                    if (u < 0 || u >= W || v < 0 || v >= H) z[q] = __uint_as_float(VO_DEPTH_OOB_BITS);
                    else src[q] = depth + (size_t)b * H * W + (size_t)v * W + u;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SD_ILP; ++q)                       // all loads issued before the first use
            if (src[q]) z[q] = __ldg(src[q]);
#pragma unroll
        for (int q = 0; q < SD_ILP; ++q) {
            const long long gid = base + (long long)q * SD_THREADS;
            if (gid < total) z_kp[gid] = z[q];
        }
    }
}

// Sparse fused path: one CTA per frame pair, chunks of GB_THREADS matches, ballot + warp-count
// scan keeps the surviving correspondences in match order (the reference's boolean-mask
// compression is order preserving and the hypothesis table indexes that order).
__global__ void __launch_bounds__(GB_THREADS)
gather_backproject_kernel(const int32_t *__restrict__ pairs, const int32_t *__restrict__ n_pairs, int pair_cap,
                          const float *__restrict__ ref_kp, const float *__restrict__ cur_kp, int n_stride, int m_stride,
                          int kp_stride, const float *__restrict__ depth, const float *__restrict__ depth_kp, int H,
                          int W, Intr k, float min_flow, float z_min, float z_max, float *__restrict__ xyz,
                          float *__restrict__ ref_uv,
                          float *__restrict__ cur_uv, int32_t *__restrict__ src, int32_t *__restrict__ n_out,
                          int32_t *__restrict__ status) {
    const int b = blockIdx.x;
    const int n = min(n_pairs[b], pair_cap);
    __shared__ int warp_cnt[GB_THREADS / 32];
    __shared__ int base_s, oob_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { base_s = 0; oob_s = 0; }
    __syncthreads();
    const float *dimg = depth ? depth + (size_t)b * H * W : nullptr;
    const float *dkp = depth_kp ? depth_kp + (size_t)b * n_stride : nullptr;  // depth already sampled per ref keypoint
    for (int m0 = 0; m0 < n; m0 += GB_THREADS) {
        const int m = m0 + threadIdx.x;
        bool keep = false;
        float rx = 0, ry = 0, cx = 0, cy = 0, X = 0, Y = 0, Z = 0;
        if (m < n) {
            const int ir = pairs[((size_t)b * pair_cap + m) * 2 + 0];
            const int ic = pairs[((size_t)b * pair_cap + m) * 2 + 1];
            const float *pr = ref_kp + ((size_t)b * n_stride + ir) * kp_stride;
            const float *pc = cur_kp + ((size_t)b * m_stride + ic) * kp_stride;
            rx = pr[0]; ry = pr[1]; cx = pc[0]; cy = pc[1];
            // np.linalg.norm(ref - cur, axis=1) in fp32, keep diff >= 3 (:260-264)
            const float dx = __fsub_rn(rx, cx), dy = __fsub_rn(ry, cy);
            const float diff = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            if (diff >= min_flow) {
                // .astype(np.int32): truncation toward zero; index [row = y, col = x] (:97)
                const int u = (int)rx, v = (int)ry;
                const bool inside = !(u < 0 || u >= W || v < 0 || v >= H);
                if (dkp) Z = __ldg(dkp + ir);
                if (!inside || (dkp && __float_as_uint(Z) == VO_DEPTH_OOB_BITS)) {
                    oob_s = 1;  // reference: IndexError (or negative wrap) -> pair fails
                } else {
                    if (!dkp) Z = __ldg(dimg + (size_t)v * W + u);
                    backproject(k, u, v, Z, X, Y);
                    keep = (Z > z_min) && (Z < z_max);  // NaN fails both
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int prefix = base_s;
        for (int w = 0; w < warp; ++w) prefix += warp_cnt[w];
        if (keep) {
            const size_t o = (size_t)b * pair_cap + prefix + __popc(bal & ((1u << lane) - 1u));
            xyz[o * 3 + 0] = X; xyz[o * 3 + 1] = Y; xyz[o * 3 + 2] = Z;
            ref_uv[o * 2 + 0] = rx; ref_uv[o * 2 + 1] = ry;
            cur_uv[o * 2 + 0] = cx; cur_uv[o * 2 + 1] = cy;
            if (src) src[o] = m;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < GB_THREADS / 32; ++w) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        n_out[b] = base_s;
        if (status) status[b] = oob_s ? VO_ST_KP_OUT_OF_IMAGE : VO_ST_OK;
    }
}

}  // namespace
}  // namespace vo

extern "C" int vo_backproject_dense(vo_ctx *ctx, const float *depth, int B, int H, int W, const double *K_h,
                                    float *xyz, void *stream) {
    using namespace vo;
    VO_REQUIRE(ctx && depth && xyz && K_h, "vo_backproject_dense: null argument");
    VO_REQUIRE(B >= 0 && H > 0 && W > 0, "vo_backproject_dense: bad shape");
    VO_REQUIRE(((uintptr_t)depth % 16) == 0 && ((uintptr_t)xyz % 16) == 0, "vo_backproject_dense: buffers must be 16B aligned");
    VO_REQUIRE(K_h[0] != 0.0 && K_h[4] != 0.0, "vo_backproject_dense: zero focal length");
    if (B == 0) return VO_OK;
    const long long n_px = (long long)B * H * W;
    const long long n_vec = n_px >> 2;
    // 4 resident CTAs of 256 threads per SM (register-limited), 4 x 512 B loads in flight per warp, grid-stride beyond
    long long want = (n_vec + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 4;
    int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    VO_PROF(ctx, (cudaStream_t)stream, VO_STAGE_DENSE);
    backproject_dense_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(depth, xyz, n_px, H, W, make_intr(K_h));
    VO_LAUNCH_CHECK(ctx);
    VO_PROF(ctx, (cudaStream_t)stream, -1);
    return VO_OK;
}

namespace vo {
int gather_backproject_impl(vo_ctx *ctx, const int32_t *pairs, const int32_t *n_pairs, int B, int pair_cap,
                            const float *ref_kp, const float *cur_kp, int n_stride, int m_stride, int kp_stride,
                            const float *depth, const float *depth_kp, int H, int W, const double *K_h,
                            float min_flow_px, float z_min, float z_max, float *xyz, float *ref_uv, float *cur_uv,
                            int32_t *src, int32_t *n_out, int32_t *status, void *stream) {
    using namespace vo;
    VO_REQUIRE(ctx && pairs && n_pairs && ref_kp && cur_kp && (depth || depth_kp) && K_h && xyz && ref_uv && cur_uv && n_out,
               "vo_gather_backproject: null argument");
    VO_REQUIRE(B >= 0 && pair_cap >= 0 && H > 0 && W > 0 && kp_stride >= 2, "vo_gather_backproject: bad shape");
    VO_REQUIRE(K_h[0] != 0.0 && K_h[4] != 0.0, "vo_gather_backproject: zero focal length");
    if (B == 0) return VO_OK;
    VO_PROF(ctx, (cudaStream_t)stream, VO_STAGE_GATHER);
    gather_backproject_kernel<<<B, GB_THREADS, 0, (cudaStream_t)stream>>>(
        pairs, n_pairs, pair_cap, ref_kp, cur_kp, n_stride, m_stride, kp_stride, depth, depth_kp, H, W, make_intr(K_h),
        min_flow_px, z_min, z_max, xyz, ref_uv, cur_uv, src, n_out, status);
    VO_LAUNCH_CHECK(ctx);
    VO_PROF(ctx, (cudaStream_t)stream, -1);
    return VO_OK;
}
}  // namespace vo

extern "C" int vo_gather_backproject(vo_ctx *ctx, const int32_t *pairs, const int32_t *n_pairs, int B, int pair_cap,
                                     const float *ref_kp, const float *cur_kp, int n_stride, int m_stride, int kp_stride,
                                     const float *depth, int H, int W, const double *K_h, float min_flow_px, float z_min,
                                     float z_max, float *xyz, float *ref_uv, float *cur_uv, int32_t *src, int32_t *n_out,
                                     int32_t *status, void *stream) {
    return vo::gather_backproject_impl(ctx, pairs, n_pairs, B, pair_cap, ref_kp, cur_kp, n_stride, m_stride, kp_stride,
                                       depth, nullptr, H, W, K_h, min_flow_px, z_min, z_max, xyz, ref_uv, cur_uv, src,
                                       n_out, status, stream);
}

extern "C" int vo_sample_depth(vo_ctx *ctx, const float *kp, int B, int n_stride, int kp_stride, const int32_t *n_kp,
                               const float *depth, int H, int W, float *depth_kp, void *stream) {
    using namespace vo;
    VO_REQUIRE(ctx && kp && depth && depth_kp, "vo_sample_depth: null argument");
    VO_REQUIRE(B >= 0 && n_stride >= 0 && kp_stride >= 2 && H > 0 && W > 0, "vo_sample_depth: bad shape");
    const long long total = (long long)B * n_stride;
    if (total == 0) return VO_OK;
    const long long want = (total + SD_THREADS * SD_ILP - 1) / (SD_THREADS * SD_ILP);
    const unsigned grid = (unsigned)(want < 148 ? want : 148);     // persistent, one thin CTA per SM: all resident in the leftover registers at once
    sample_depth_kernel<<<grid, SD_THREADS, 0, (cudaStream_t)stream>>>(kp, n_stride, kp_stride, n_kp, depth, H, W, total, depth_kp);
    VO_LAUNCH_CHECK(ctx);
    return VO_OK;
}
