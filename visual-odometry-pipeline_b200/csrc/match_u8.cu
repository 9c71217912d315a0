// 256-bit binary-descriptor matcher (ORB): XOR + POPC (Hamming) or byte-wise squared L2.
//
// Replaces cv2.BFMatcher.knnMatch on uint8 descriptors (feature_extractors/ORB.py:8,25;
// the reference's BFMatcher() is NORM_L2 over byte values) and
// cv2.BFMatcher(NORM_HAMMING, crossCheck=True) (the north-star semantics).
//
// Layout: every thread keeps ROWS_PT reference descriptors in registers (8 words each) and
// walks the current-frame descriptors, which the CTA stages in shared memory; all lanes of a
// warp read the same staged descriptor (smem broadcast, 2 x LDS.128 per 32 x ROWS_PT
// distances).  Row top-2 is therefore thread-private; the column arg-min is a warp REDUX.MIN
// over a packed (distance << 10 | row-in-CTA) key, merged per CTA in shared memory and per
// grid with one 64-bit atomicMin per (CTA, column).  The N x M distance matrix never exists.
//
// Hamming arithmetic: the POPC pipe issues 16 lanes/clk/SM, a quarter of the logic pipe, so a
// plain 8 x (XOR, POPC) distance is POPC-bound.  The eight XOR words are first compressed by a
// carry-save adder tree into two words of weight 1, one of weight 2 and one of weight 4, which
// leaves 4 POPC per distance (the floor: a bit column of eight words takes nine values).  On
// descriptors in prefix-XOR form (hamming_math.cuh) the tree costs 13 LOP3 instead of 16.
#include "common.cuh"
#include "hamming_math.cuh"
#include <stdlib.h>

namespace vo {
namespace {

constexpr int U8_THREADS = 256;
constexpr int U8_ROWS_PT = 4;
constexpr int U8_ROW_BITS = 10;
constexpr int U8_ROWS_CTA = U8_THREADS * U8_ROWS_PT;  // 1024 -> row-in-CTA fits 10 bits
constexpr int U8_CHUNK = 512;                         // staged columns per pass (16 KB + 16 KB of per-warp column minima)
constexpr uint32_t U8_INVALID_ROW = 0x80000000u;      // added to the column key of rows >= N: never the minimum
constexpr int U8_COL_BITS = 20;                       // Hamming row key = dist << 20 | column
static_assert(U8_ROWS_CTA == (1 << U8_ROW_BITS), "row-in-CTA must fit the key");

// Multipliers the compiler cannot see through: `x * opaque(4) + y` stays an IMAD (FMA pipe) instead of being strength-
// reduced to LEA / IADD3 on the logic pipe, which is the pipe this kernel is bound by.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
struct U8Mul {
    uint32_t m1, m2, m4, mrow, mcol;
};

template <int NORM>
__device__ __forceinline__ uint32_t dist256(const uint32_t (&a)[8], const uint4 b0, const uint4 b1, const U8Mul &mu) {
    const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if (NORM == VO_NORM_HAMMING) {
        // a, b in prefix-XOR form (hamming_math.cuh): 13 LOP3 + 4 POPC; the weighted sum is three IMADs (FMA pipe)
        const HammingPlanes p = hamming_planes(a, b);
        const uint32_t t1 = (uint32_t)__popc(p.twos) * mu.m2 + (uint32_t)__popc(p.ones_a);
        const uint32_t t2 = (uint32_t)__popc(p.fours) * mu.m4 + (uint32_t)__popc(p.ones_b);
        return t1 * mu.m1 + t2;
    } else {
        uint32_t d = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t ad = __vabsdiffu4(a[w], b[w]);
            d = __dp4a(ad, ad, d);  // sum of squared byte differences, exact in u32
        }
        return d;
    }
}

template <int NORM, bool SECOND>
__global__ void __launch_bounds__(U8_THREADS, 3)
match_u8_kernel(const uint8_t *__restrict__ ref, const uint8_t *__restrict__ cur, int n_stride, int m_stride,
                const int32_t *__restrict__ n_ref, const int32_t *__restrict__ n_cur, int n_split,
                vo_row_partial *__restrict__ part, unsigned long long *__restrict__ colkey) {
    __shared__ uint4 sdesc[U8_CHUNK * 2];
    __shared__ uint32_t swcol[U8_THREADS / 32][U8_CHUNK];  // per-warp column minima of the staged chunk
    // Hamming distances are <= 256, so (dist << 20 | column) orders rows by distance, then by lowest column:
    // the row top-2 is two integer minima per distance instead of compare/select chains.
    constexpr bool PACKED = (NORM == VO_NORM_HAMMING);

    const int b = blockIdx.z, split = blockIdx.y;
    const int N = n_ref ? min(n_ref[b], n_stride) : n_stride;
    const int M = n_cur ? min(n_cur[b], m_stride) : m_stride;
    const int row_base = blockIdx.x * U8_ROWS_CTA;
    const int cols_per_split = (M + n_split - 1) / n_split;
    const int c_begin = split * cols_per_split;
    const int c_end = min(M, c_begin + cols_per_split);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t a[U8_ROWS_PT][8];
    uint32_t rowid[U8_ROWS_PT];                     // row-in-CTA, or that + INVALID for rows past N
    uint32_t s1[U8_ROWS_PT], s2[U8_ROWS_PT];        // PACKED: keys; else distances
    int32_t i1[U8_ROWS_PT], i2[U8_ROWS_PT];
#pragma unroll
    for (int r = 0; r < U8_ROWS_PT; ++r) {
        const int row = row_base + r * U8_THREADS + tid;
        const bool valid = row < N;
        const uint4 *p = reinterpret_cast<const uint4 *>(ref + ((size_t)b * n_stride + (valid ? row : 0)) * 32);
        uint4 q0 = valid ? p[0] : make_uint4(0, 0, 0, 0), q1 = valid ? p[1] : make_uint4(0, 0, 0, 0);
        a[r][0] = q0.x; a[r][1] = q0.y; a[r][2] = q0.z; a[r][3] = q0.w;
        a[r][4] = q1.x; a[r][5] = q1.y; a[r][6] = q1.z; a[r][7] = q1.w;
        if (NORM == VO_NORM_HAMMING) hamming_prefix_form(a[r]);
        rowid[r] = (uint32_t)(r * U8_THREADS + tid) + (valid ? 0u : U8_INVALID_ROW);
        s1[r] = s2[r] = 0xffffffffu;
        i1[r] = i2[r] = -1;
    }

    const uint4 *cur4 = reinterpret_cast<const uint4 *>(cur + (size_t)b * m_stride * 32);
    const U8Mul mu = {opaque_u32(1u), opaque_u32(2u), opaque_u32(4u), opaque_u32(1u << U8_ROW_BITS), opaque_u32(1u << U8_COL_BITS)};
    if (row_base < N) {
        for (int c0 = c_begin; c0 < c_end; c0 += U8_CHUNK) {
            const int cnt = min(U8_CHUNK, c_end - c0);
            __syncthreads();  // previous chunk fully consumed
            if (NORM == VO_NORM_HAMMING) {
                for (int t = tid; t < cnt; t += U8_THREADS) {
                    const uint4 q0 = cur4[(size_t)(c0 + t) * 2], q1 = cur4[(size_t)(c0 + t) * 2 + 1];
                    uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                    hamming_prefix_form(w);
                    sdesc[2 * t] = make_uint4(w[0], w[1], w[2], w[3]);
                    sdesc[2 * t + 1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
            } else {
                for (int t = tid; t < cnt * 2; t += U8_THREADS) sdesc[t] = cur4[(size_t)c0 * 2 + t];
            }
            __syncthreads();

            int j = 0;
            if (PACKED && !SECOND) {
                // two columns per iteration: the row minimum of both keys and the running one is a single 3-input
                // VIMNMX3, the column minimum over the thread's four rows is VIMNMX3 + VIMNMX, and lane 0 stores both
                // warp minima with one 64-bit STS
                for (; j + 1 < cnt; j += 2) {
                    const uint4 b00 = sdesc[2 * j], b01 = sdesc[2 * j + 1], b10 = sdesc[2 * j + 2], b11 = sdesc[2 * j + 3];
                    const uint32_t col = (uint32_t)(c0 + j);
                    uint32_t ck0[U8_ROWS_PT], ck1[U8_ROWS_PT];
#pragma unroll
                    for (int r = 0; r < U8_ROWS_PT; ++r) {
                        const uint32_t d0 = dist256<NORM>(a[r], b00, b01, mu), d1 = dist256<NORM>(a[r], b10, b11, mu);
                        s1[r] = min(min(d0 * mu.mcol + col, d1 * mu.mcol + (col + 1u)), s1[r]);
                        ck0[r] = d0 * mu.mrow + rowid[r];
                        ck1[r] = d1 * mu.mrow + rowid[r];
                    }
                    static_assert(U8_ROWS_PT == 4, "column reduction below is written for 4 rows per thread");
                    const uint32_t m0 = min(min(min(ck0[0], ck0[1]), ck0[2]), ck0[3]);
                    const uint32_t m1 = min(min(min(ck1[0], ck1[1]), ck1[2]), ck1[3]);
                    const uint32_t w0 = __reduce_min_sync(0xffffffffu, m0), w1 = __reduce_min_sync(0xffffffffu, m1);
                    if (lane == 0) *reinterpret_cast<uint2 *>(&swcol[warp][j]) = make_uint2(w0, w1);
                }
            }
#pragma unroll 2
            for (; j < cnt; ++j) {
                const uint4 b0 = sdesc[2 * j], b1 = sdesc[2 * j + 1];
                const uint32_t col = (uint32_t)(c0 + j);
                uint32_t ckey = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < U8_ROWS_PT; ++r) {
                    const uint32_t d = dist256<NORM>(a[r], b0, b1, mu);
                    if (PACKED) {
                        const uint32_t k = (d << U8_COL_BITS) + col;
                        if (SECOND) s2[r] = min(s2[r], max(s1[r], k));
                        s1[r] = min(s1[r], k);
                    } else {
                        if (d < s1[r]) {
                            if (SECOND) { s2[r] = s1[r]; i2[r] = i1[r]; }
                            s1[r] = d; i1[r] = (int32_t)col;
                        } else if (SECOND && d < s2[r]) {
                            s2[r] = d; i2[r] = (int32_t)col;
                        }
                    }
                    ckey = min(ckey, (d << U8_ROW_BITS) + rowid[r]);
                }
                const uint32_t wmin = __reduce_min_sync(0xffffffffu, ckey);
                if (lane == 0) swcol[warp][j] = wmin;  // warp-private slot: no atomics in the inner loop
            }
            __syncthreads();
            for (int t = tid; t < cnt; t += U8_THREADS) {
                uint32_t k = swcol[0][t];
#pragma unroll
                for (int w = 1; w < U8_THREADS / 32; ++w) k = min(k, swcol[w][t]);
                if (k < U8_INVALID_ROW) {
                    const unsigned long long g = ((unsigned long long)(k >> U8_ROW_BITS) << 32) |
                                                 (unsigned long long)(uint32_t)(row_base + (int)(k & (U8_ROWS_CTA - 1)));
                    atomicMin(&colkey[(size_t)b * m_stride + c0 + t], g);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < U8_ROWS_PT; ++r) {
        const int row = row_base + r * U8_THREADS + tid;
        if (row < n_stride) {
            vo_row_partial p;
            if (PACKED) {
                const bool v1 = row < N && s1[r] != 0xffffffffu, v2 = row < N && SECOND && s2[r] != 0xffffffffu;
                p.s1 = v1 ? (s1[r] >> U8_COL_BITS) : 0xffffffffu;
                p.i1 = v1 ? (int32_t)(s1[r] & ((1u << U8_COL_BITS) - 1)) : -1;
                p.s2 = v2 ? (s2[r] >> U8_COL_BITS) : 0xffffffffu;
                p.i2 = v2 ? (int32_t)(s2[r] & ((1u << U8_COL_BITS) - 1)) : -1;
            } else {
                const bool v = row < N;
                p.s1 = v ? s1[r] : 0xffffffffu; p.s2 = v ? s2[r] : 0xffffffffu;
                p.i1 = v ? i1[r] : -1; p.i2 = v ? i2[r] : -1;
            }
            const size_t at = ((size_t)b * n_split + split) * n_stride + row;
            if (SECOND) {
                part[at] = p;
            } else {  // nothing reads the second best: 8-byte partials (SCORE_COMPACT_PARTIALS)
                vo_row_best q;
                q.s1 = p.s1; q.i1 = p.i1;
                reinterpret_cast<vo_row_best *>(part)[at] = q;
            }
        }
    }
}

__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

}  // namespace

int fill_u64(vo_ctx *ctx, unsigned long long *p, size_t n, unsigned long long v, cudaStream_t st) {
    if (n == 0) return VO_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    fill_u64_kernel<<<blocks, 256, 0, st>>>(p, n, v);
    VO_LAUNCH_CHECK(ctx);
    return VO_OK;
}

}  // namespace vo

extern "C" int vo_match_u8(vo_ctx *ctx, const uint8_t *ref, const uint8_t *cur, int B, int n_stride, int m_stride,
                           const int32_t *n_ref, const int32_t *n_cur, int bytes, int norm, int mode, double ratio,
                           int32_t *out_pairs, float *out_dist, int32_t *out_count, const vo_knn_out *knn,
                           void *stream) {
    using namespace vo;
    VO_REQUIRE(ctx, "vo_match_u8: null ctx");
    VO_REQUIRE(bytes == 32 || (bytes == 128 && norm == VO_NORM_L2_U8),
               "vo_match_u8: 32-byte (256-bit) descriptors, or 128-byte descriptors under VO_NORM_L2_U8 (SIFT values as uint8), got %d", bytes);
    VO_REQUIRE(norm == VO_NORM_HAMMING || norm == VO_NORM_L2_U8 || norm == VO_NORM_HAMMING_TC, "vo_match_u8: bad norm %d", norm);
    VO_REQUIRE(mode >= VO_MODE_RATIO && mode <= VO_MODE_NN, "vo_match_u8: bad mode %d", mode);
    VO_REQUIRE(mode != VO_MODE_THRESH && mode != VO_MODE_THRESH_MUTUAL && mode != VO_MODE_RATIO_MUTUAL,
               "vo_match_u8: similarity modes need float descriptors");
    VO_REQUIRE(B >= 0 && n_stride >= 0 && m_stride >= 0, "vo_match_u8: negative size");
    VO_REQUIRE(out_pairs && out_count, "vo_match_u8: null output");
    VO_REQUIRE(((uintptr_t)ref % 16) == 0 && ((uintptr_t)cur % 16) == 0, "vo_match_u8: descriptors must be 16B aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return VO_OK;
    if (n_stride == 0 || m_stride == 0) {
        VO_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t) * B, st));
        return VO_OK;
    }
    VO_REQUIRE(norm != VO_NORM_HAMMING || m_stride < (1 << U8_COL_BITS),
               "vo_match_u8: Hamming matcher supports at most %d descriptors per frame", (1 << U8_COL_BITS) - 1);
    // The reference's ORB semantics (BFMatcher NORM_L2 over byte values) is a float-style L2 match of
    // integer-valued vectors: without a column arg-min (ratio / plain-NN rules) it runs as the exact fp16 tensor-core
    // pass SIFT uses, bytes widened to fp16 and zero-padded to 128 dimensions (3x the CUDA-core kernel: the pass is bound
    // by its row top-2 epilogue, not by the padded GEMM).  Same scores (exact integers), same tie rule, same finalize.
    const bool need_cols = mode == VO_MODE_MUTUAL || (knn && knn->col_idx);
    VO_REQUIRE(bytes == 32 || !need_cols, "vo_match_u8: 128-byte descriptors support the rules without a column side (ratio, NN)");
    if (norm == VO_NORM_L2_U8 && !need_cols && (bytes == 128 || !getenv("VO_U8_L2_SIMT"))) {
        int rc;
        unsigned long long *colkey_tc;
        if ((rc = ws_get(ctx, WS_COLKEY, sizeof(unsigned long long) * (size_t)B * m_stride, (void **)&colkey_tc))) return rc;
        vo_row_partial *part_tc;
        const float *row_norm_tc = nullptr;
        int n_split_tc;
        if ((rc = match_f32_tc(ctx, reinterpret_cast<const float *>(ref), reinterpret_cast<const float *>(cur), B, n_stride,
                               m_stride, n_ref, n_cur, VO_METRIC_L2, 16, 0, &part_tc, &n_split_tc, colkey_tc, &row_norm_tc, st, bytes)))
            return rc;
        return match_finalize(ctx, part_tc, n_split_tc, colkey_tc, B, n_stride, m_stride, n_ref, n_cur, SCORE_L2SQ_F32, mode,
                              ratio, row_norm_tc, out_pairs, out_dist, out_count, knn, nullptr, st);
    }
    // Opt-in tensor-core Hamming (VO_NORM_HAMMING_TC; VO_NORM_HAMMING stays XOR + POPC, what the north-star prescribes): with the
    // 256 bits as e4m3 -1 / +1, a.b = 256 - 2 popcount(a xor b), so the single pass of the tcgen05 matcher over K = 256 orders
    // by Hamming distance (match_bits_tc, csrc/match_f32_tc.cu).  Exact integers throughout, same tie rules, same finalize:
    // results are bit-identical to the XOR + POPC kernel (tests/test_gpu_match_u8.py).
    if (norm == VO_NORM_HAMMING_TC) {
        int rc;
        unsigned long long *colkey_tc;
        if ((rc = ws_get(ctx, WS_COLKEY, sizeof(unsigned long long) * (size_t)B * m_stride, (void **)&colkey_tc))) return rc;
        vo_row_partial *part_tc;
        int n_split_tc;
        // the second best of a row is read by the ratio rule and by the k-NN output only
        const bool need_second = mode == VO_MODE_RATIO || (knn && (knn->row_idx || knn->row_val));
        if ((rc = match_bits_tc(ctx, ref, cur, B, n_stride, m_stride, n_ref, n_cur, need_cols ? 1 : 0, need_second ? 1 : 0, &part_tc,
                                &n_split_tc, colkey_tc, st)))
            return rc;
        return match_finalize(ctx, part_tc, n_split_tc, colkey_tc, B, n_stride, m_stride, n_ref, n_cur,
                              SCORE_HAMMING_F32 | (need_second ? 0 : SCORE_COMPACT_PARTIALS), mode, ratio, nullptr, out_pairs, out_dist,
                              out_count, knn, nullptr, st);
    }
    const int row_blocks = ceil_div(n_stride, U8_ROWS_CTA);
    // Column splits: every CTA costs the same, so pick the smallest split count whose CTA total fills the
    // resident slots (3 CTAs per SM) to >= 94 % in its last wave, with >= 256 columns per split.
    int n_split = 1;
    {
        const int slots = 3 * ctx->sm_count;
        const long long have = (long long)B * row_blocks;
        int max_split = m_stride / 256 > 0 ? m_stride / 256 : 1;
        if (max_split > 64) max_split = 64;
        double best_eff = -1.0;
        for (int s = 1; s <= max_split; ++s) {
            const long long total = have * s;
            const long long waves = (total + slots - 1) / slots;
            const double eff = (double)total / (double)(waves * slots);
            if (eff > best_eff + 1e-9) { best_eff = eff; n_split = s; }
            if (eff >= 0.94) { n_split = s; break; }
        }
    }
    vo_row_partial *part;
    unsigned long long *colkey;
    int rc;
    if ((rc = ws_get(ctx, WS_ROWPART, sizeof(vo_row_partial) * (size_t)B * n_split * n_stride, (void **)&part))) return rc;
    if ((rc = ws_get(ctx, WS_COLKEY, sizeof(unsigned long long) * (size_t)B * m_stride, (void **)&colkey))) return rc;
    VO_PROF(ctx, st, VO_STAGE_FILL);
    if ((rc = fill_u64(ctx, colkey, (size_t)B * m_stride, ~0ull, st))) return rc;

    const bool second = (mode == VO_MODE_RATIO) || (knn && (knn->row_idx || knn->row_val));
    dim3 grid(row_blocks, n_split, B);
    VO_PROF(ctx, st, VO_STAGE_MATCH);
#define LAUNCH_U8(NORM, SEC) \
    match_u8_kernel<NORM, SEC><<<grid, U8_THREADS, 0, st>>>(ref, cur, n_stride, m_stride, n_ref, n_cur, n_split, part, colkey)
    if (norm == VO_NORM_HAMMING) {
        if (second) LAUNCH_U8(VO_NORM_HAMMING, true); else LAUNCH_U8(VO_NORM_HAMMING, false);
    } else {
        if (second) LAUNCH_U8(VO_NORM_L2_U8, true); else LAUNCH_U8(VO_NORM_L2_U8, false);
    }
#undef LAUNCH_U8
    VO_LAUNCH_CHECK(ctx);
    return match_finalize(ctx, part, n_split, colkey, B, n_stride, m_stride, n_ref, n_cur,
                          (norm == VO_NORM_HAMMING ? SCORE_HAMMING : SCORE_L2SQ_U32) | (second ? 0 : SCORE_COMPACT_PARTIALS), mode, ratio,
                          nullptr, out_pairs, out_dist, out_count, knn, nullptr, st);
}
