// Float-descriptor matcher, CUDA-core FP32 validation kernel (VO_PREC_FP32_SIMT).
//
// Computes the reference's own arithmetic form — direct sum of (a-b)^2 for L2
// (cv2.BFMatcher(NORM_L2).knnMatch, feature_extractors/SIFT.py:11,27) or the plain dot product
// for cosine similarity (R2D2.py:56) — in FP32 with a fixed k-ascending accumulation order,
// fused with the row top-2 / column arg-min reductions.  It exists to cross-check the tcgen05
// kernel on the device and to serve descriptor shapes the tensor-core path does not take;
// it is not tuned beyond shared-memory tiling.
#include "common.cuh"

namespace vo {
namespace {

constexpr int FS_TM = 64, FS_TN = 64, FS_KC = 32, FS_THREADS = 256;

template <int METRIC>
__global__ void __launch_bounds__(FS_THREADS)
match_f32_simt_kernel(const float *__restrict__ ref, const float *__restrict__ cur, int n_stride, int m_stride,
                      int dim, const int32_t *__restrict__ n_ref, const int32_t *__restrict__ n_cur, int n_split,
                      vo_row_partial *__restrict__ part, unsigned long long *__restrict__ colkey) {
    __shared__ float As[FS_TM][FS_KC + 1];
    __shared__ float Bs[FS_TN][FS_KC + 1];
    __shared__ float Ss[FS_TM][FS_TN + 1];

    const int b = blockIdx.z, split = blockIdx.y;
    const int N = n_ref ? min(n_ref[b], n_stride) : n_stride;
    const int M = n_cur ? min(n_cur[b], m_stride) : m_stride;
    const int row0 = blockIdx.x * FS_TM;
    const int tiles_total = (M + FS_TN - 1) / FS_TN;
    const int tiles_per_split = (tiles_total + n_split - 1) / n_split;
    const int t_begin = split * tiles_per_split;
    const int t_end = min(tiles_total, t_begin + tiles_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float *A = ref + (size_t)b * n_stride * dim;
    const float *Bm = cur + (size_t)b * m_stride * dim;

    float s1 = INFINITY, s2 = INFINITY;  // row state lives in threads 0..63
    int32_t i1 = -1, i2 = -1;

    for (int t = t_begin; t < t_end; ++t) {
        const int col0 = t * FS_TN;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < dim; k0 += FS_KC) {
            __syncthreads();
            for (int e = tid; e < FS_TM * FS_KC; e += FS_THREADS) {
                const int r = e / FS_KC, k = e % FS_KC;
                const int gr = row0 + r, gc = col0 + r, gk = k0 + k;
                As[r][k] = (gr < N && gk < dim) ? A[(size_t)gr * dim + gk] : 0.f;
                Bs[r][k] = (gc < M && gk < dim) ? Bm[(size_t)gc * dim + gk] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < FS_KC; ++k) {
                float av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][k];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = Bs[tx * 4 + j][k];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (METRIC == VO_METRIC_L2) {
                            const float d = __fsub_rn(av[i], bv[j]);
                            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(d, d));
                        } else {
                            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
                        }
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                Ss[ty * 4 + i][tx * 4 + j] = (METRIC == VO_METRIC_L2) ? acc[i][j] : -acc[i][j];
        __syncthreads();
        if (tid < FS_TM) {  // row scan: strict '<' in ascending column order = lowest index on ties
            const int row = row0 + tid;
            if (row < N) {
                const int lim = min(FS_TN, M - col0);
                for (int j = 0; j < lim; ++j) {
                    const float s = Ss[tid][j];
                    if (s < s1) {
                        s2 = s1; i2 = i1; s1 = s; i1 = col0 + j;
                    } else if (s < s2) {
                        s2 = s; i2 = col0 + j;
                    }
                }
            }
        } else if (tid < FS_TM + FS_TN) {  // column scan
            const int j = tid - FS_TM;
            const int col = col0 + j;
            if (col < M) {
                const int lim = min(FS_TM, N - row0);
                float best = INFINITY;
                int brow = -1;
                for (int r = 0; r < lim; ++r) {
                    const float s = Ss[r][j];
                    if (s < best || brow < 0) { best = s; brow = r; }
                }
                if (brow >= 0) {
                    const unsigned long long key = ((unsigned long long)float_to_ordered(best) << 32) |
                                                   (unsigned long long)(uint32_t)(row0 + brow);
                    atomicMin(&colkey[(size_t)b * m_stride + col], key);
                }
            }
        }
    }
    if (tid < FS_TM) {
        const int row = row0 + tid;
        if (row < n_stride) {
            vo_row_partial p;
            p.s1 = float_to_ordered(s1); p.s2 = float_to_ordered(s2);
            p.i1 = i1; p.i2 = i2;
            part[((size_t)b * n_split + split) * n_stride + row] = p;
        }
    }
}

}  // namespace

int match_f32_simt(vo_ctx *ctx, const float *ref, const float *cur, int B, int n_stride, int m_stride, int dim,
                   const int32_t *n_ref, const int32_t *n_cur, int metric, vo_row_partial *part, int n_split,
                   unsigned long long *colkey, cudaStream_t st) {
    dim3 grid(ceil_div(n_stride, FS_TM), n_split, B);
    if (metric == VO_METRIC_L2)
        match_f32_simt_kernel<VO_METRIC_L2><<<grid, FS_THREADS, 0, st>>>(ref, cur, n_stride, m_stride, dim, n_ref, n_cur,
                                                                         n_split, part, colkey);
    else
        match_f32_simt_kernel<VO_METRIC_COSINE><<<grid, FS_THREADS, 0, st>>>(ref, cur, n_stride, m_stride, dim, n_ref,
                                                                             n_cur, n_split, part, colkey);
    VO_LAUNCH_CHECK(ctx);
    return VO_OK;
}

}  // namespace vo
