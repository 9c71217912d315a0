// Matcher finalize: merge per-split row partials, apply the acceptance rule, compact.
//
// Two launches.  `finalize_kernel`: one CTA per (2048-row chunk, frame pair) merges, decides and compacts its chunk in
// place (accepted pairs of chunk c sit at out_pairs[c * 2048 ...], their number in chunk_count).  `finalize_pack_kernel`:
// one CTA per pair scans the chunk counts and slides every chunk's segment left to its final offset (a segment only
// ever moves towards lower addresses and never into the source of a later chunk: read all -> barrier -> write all).
// With one CTA per pair doing everything, 32 pairs of 20k rows kept 32 of 148 SMs busy for 40 barrier-separated rounds
// (0.176 ms at c4).  Input is what the distance kernels leave behind:
//   part   [B][n_split][n_stride]  best / second best score + column of every ref row
//   colkey [B][m_stride]           (score << 32 | ref row) minimum of every cur column
// Output is the reference's `get_matches` result (feature_extractors/SIFT.py:25-34,
// ORB.py:23-32, R2D2.py:29-66): accepted (ref, cur) index pairs in ascending ref order.
#include "common.cuh"

namespace vo {

namespace {

constexpr int FIN_THREADS = 512;
constexpr int FIN_ROUNDS = 4;                          // rows per thread, at most
constexpr int FIN_CHUNK = FIN_THREADS * FIN_ROUNDS;    // rows per CTA, at most (the host picks 512 .. 2048: chunk_rows)

__device__ __forceinline__ void top2_insert(uint32_t s, int32_t i, uint32_t &s1, int32_t &i1, uint32_t &s2,
                                            int32_t &i2) {
    // ties resolve to the lower column index (cv2 knnMatch order), whatever order the partials arrive in:
    // the tcgen05 kernel's two epilogue groups own interleaved column tiles
    if (i < 0) return;
    if (s < s1 || (s == s1 && i < i1)) {
        s2 = s1; i2 = i1; s1 = s; i1 = i;
    } else if (s < s2 || (s == s2 && (i2 < 0 || i < i2))) {
        s2 = s; i2 = i;
    }
}

// value reported for a score: distance (u8 / L2) or similarity (cosine)
__device__ __forceinline__ float score_value(int kind, uint32_t s, float rn) {
    switch (kind) {
        case SCORE_HAMMING: return (float)s;
        case SCORE_L2SQ_U32: return __fsqrt_rn((float)s);
        case SCORE_L2SQ_F32: {
            float d2 = __fadd_rn(ordered_to_float(s), rn);
            return __fsqrt_rn(fmaxf(d2, 0.0f));
        }
        case SCORE_HAMMING_F32: return __fmul_rn(0.5f, __fadd_rn(256.0f, ordered_to_float(s)));   // key = -(a.b) = 2 Hamming - 256
        default: return -ordered_to_float(s);  // SCORE_NEGSIM_F32 -> similarity
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_kernel(const vo_row_partial *__restrict__ part, int n_split, const unsigned long long *__restrict__ colkey,
                int n_stride, int m_stride, const int32_t *__restrict__ n_ref, const int32_t *__restrict__ n_cur,
                int kind, int mode, double param, const float *__restrict__ row_norm, int32_t *__restrict__ out_pairs,
                float *__restrict__ out_dist, int32_t *__restrict__ chunk_count, int32_t *__restrict__ knn_row_idx,
                float *__restrict__ knn_row_val, int32_t *__restrict__ knn_col_idx, uint8_t *__restrict__ near_tie,
                int chunk_rows) {
    const bool compact = (kind & SCORE_COMPACT_PARTIALS) != 0;
    kind &= ~SCORE_COMPACT_PARTIALS;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int chunk0 = chunk * chunk_rows;
    const int N = n_ref ? min(n_ref[b], n_stride) : n_stride;
    const int M = n_cur ? min(n_cur[b], m_stride) : m_stride;
    const unsigned long long *ck = colkey + (size_t)b * m_stride;
    __shared__ int warp_cnt[FIN_THREADS / 32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;

    if (knn_col_idx) {
        for (int j = chunk * FIN_THREADS + threadIdx.x; j < m_stride; j += gridDim.x * FIN_THREADS) {
            unsigned long long k = (j < M) ? ck[j] : ~0ull;
            knn_col_idx[(size_t)b * m_stride + j] = (k == ~0ull) ? -1 : (int32_t)(uint32_t)(k & 0xffffffffull);
        }
    }
    __syncthreads();

    for (int r0 = chunk0; r0 < min(n_stride, chunk0 + chunk_rows); r0 += FIN_THREADS) {
        const int row = r0 + threadIdx.x;
        bool keep = false;
        int32_t i1 = -1, i2 = -1;
        float v1 = 0.f, v2 = 0.f, dist1 = 0.f;
        if (row < N) {
            uint32_t s1 = 0xffffffffu, s2 = 0xffffffffu;
            for (int sp = 0; sp < n_split; ++sp) {
                const size_t at = ((size_t)b * n_split + sp) * n_stride + row;
                if (compact) {
                    const vo_row_best q = reinterpret_cast<const vo_row_best *>(part)[at];
                    top2_insert(q.s1, q.i1, s1, i1, s2, i2);
                } else {
                    const vo_row_partial p = part[at];
                    top2_insert(p.s1, p.i1, s1, i1, s2, i2);
                    top2_insert(p.s2, p.i2, s1, i1, s2, i2);
                }
            }
            const float rn = row_norm ? row_norm[(size_t)b * n_stride + row] : 0.f;
            if (i1 >= 0) v1 = score_value(kind, s1, rn);
            if (i2 >= 0) v2 = score_value(kind, s2, rn);
            float dist2 = v2;
            dist1 = v1;
            if (kind == SCORE_NEGSIM_F32) {  // torch.sqrt(2 - 2*sim), fp32 (R2D2.py:59); NaN when sim > 1
                dist1 = __fsqrt_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, v1)));
                dist2 = __fsqrt_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, v2)));
            }
            bool mutual = false;
            if (i1 >= 0 && i1 < M) mutual = ((uint32_t)(ck[i1] & 0xffffffffull) == (uint32_t)row);
            switch (mode) {
                case VO_MODE_RATIO:  // m.distance < 0.85*n.distance, Python doubles
                    keep = (i2 >= 0) && ((double)dist1 < param * (double)dist2);
                    break;
                case VO_MODE_MUTUAL: keep = mutual; break;
                case VO_MODE_RATIO_MUTUAL: {
                    float ratio = __fdiv_rn(dist1, __fadd_rn(dist2, 1e-8f));
                    keep = (i2 >= 0) && mutual && ((double)ratio <= param);
                    break;
                }
                case VO_MODE_THRESH_MUTUAL: keep = (i1 >= 0) && mutual && ((double)v1 >= param); break;
                case VO_MODE_THRESH: keep = (i1 >= 0) && ((double)v1 >= param); break;
                default: keep = (i1 >= 0); break;  // VO_MODE_NN
            }
            if (near_tie) {
                bool nt = (i2 >= 0) && !(__fsub_rn(dist2, dist1) > 1e-5f * dist2);
                near_tie[(size_t)b * n_stride + row] = nt ? 1 : 0;
            }
        } else if (row < n_stride && near_tie) {
            near_tie[(size_t)b * n_stride + row] = 0;
        }
        if (row < n_stride) {
            if (knn_row_idx) {
                knn_row_idx[((size_t)b * n_stride + row) * 2 + 0] = i1;
                knn_row_idx[((size_t)b * n_stride + row) * 2 + 1] = i2;
            }
            if (knn_row_val) {
                knn_row_val[((size_t)b * n_stride + row) * 2 + 0] = v1;
                knn_row_val[((size_t)b * n_stride + row) * 2 + 1] = v2;
            }
        }
        // order-preserving compaction of this chunk
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int prefix = base_s;
        for (int w = 0; w < warp; ++w) prefix += warp_cnt[w];
        if (keep) {
            int pos = chunk0 + prefix + __popc(bal & ((1u << lane) - 1u));  // chunk-local compaction, in place
            out_pairs[((size_t)b * n_stride + pos) * 2 + 0] = row;
            out_pairs[((size_t)b * n_stride + pos) * 2 + 1] = i1;
            if (out_dist) out_dist[(size_t)b * n_stride + pos] = dist1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < FIN_THREADS / 32; ++w) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_count[b * gridDim.x + chunk] = base_s;
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_pack_kernel(const int32_t *__restrict__ chunk_count, int n_chunks, int chunk_rows, int n_stride,
                     int32_t *__restrict__ out_pairs, float *__restrict__ out_dist, int32_t *__restrict__ out_count) {
    const int b = blockIdx.x;
    int2 *pairs = reinterpret_cast<int2 *>(out_pairs) + (size_t)b * n_stride;
    float *dist = out_dist ? out_dist + (size_t)b * n_stride : nullptr;
    int off = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int cnt = chunk_count[b * n_chunks + c];
        const int src = c * chunk_rows;
        if (off != src && cnt > 0) {
            int2 p[FIN_ROUNDS];
            float d[FIN_ROUNDS];
#pragma unroll
            for (int r = 0; r < FIN_ROUNDS; ++r) {
                const int i = r * FIN_THREADS + threadIdx.x;
                if (i < cnt) { p[r] = pairs[src + i]; if (dist) d[r] = dist[src + i]; }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FIN_ROUNDS; ++r) {
                const int i = r * FIN_THREADS + threadIdx.x;
                if (i < cnt) { pairs[off + i] = p[r]; if (dist) dist[off + i] = d[r]; }
            }
            __syncthreads();
        }
        off += cnt;
    }
    if (threadIdx.x == 0) out_count[b] = off;
}

}  // namespace

int match_finalize(vo_ctx *ctx, const vo_row_partial *part, int n_split, const unsigned long long *colkey, int B,
                   int n_stride, int m_stride, const int32_t *n_ref, const int32_t *n_cur, int score_kind, int mode,
                   double param, const float *row_norm, int32_t *out_pairs, float *out_dist, int32_t *out_count,
                   const vo_knn_out *knn, uint8_t *near_tie, cudaStream_t st) {
    // 2048-row chunks when there are pairs enough to fill the GPU, 512-row chunks for the single-pair calls of the keyframe
    // loop (one CTA per 512 rows: the merge of the per-split partials is a latency chain per row)
    const int chunk_rows = ((long long)B * ceil_div(n_stride, FIN_CHUNK) >= ctx->sm_count) ? FIN_CHUNK : FIN_THREADS;
    const int n_chunks = max(1, ceil_div(n_stride, chunk_rows));
    int32_t *chunk_count = out_count;  // a single chunk per pair is already the final layout: no pack launch
    int rc;
    if (n_chunks > 1 && (rc = ws_get(ctx, WS_FIN, sizeof(int32_t) * (size_t)B * n_chunks, (void **)&chunk_count))) return rc;
    VO_PROF(ctx, st, VO_STAGE_FINALIZE);
    finalize_kernel<<<dim3(n_chunks, B), FIN_THREADS, 0, st>>>(part, n_split, colkey, n_stride, m_stride, n_ref, n_cur, score_kind,
                                                               mode, param, row_norm, out_pairs, out_dist, chunk_count,
                                                               knn ? knn->row_idx : nullptr, knn ? knn->row_val : nullptr,
                                                               knn ? knn->col_idx : nullptr, near_tie, chunk_rows);
    VO_LAUNCH_CHECK(ctx);
    if (n_chunks > 1) {
        finalize_pack_kernel<<<B, FIN_THREADS, 0, st>>>(chunk_count, n_chunks, chunk_rows, n_stride, out_pairs, out_dist, out_count);
        VO_LAUNCH_CHECK(ctx);
    }
    VO_PROF(ctx, st, -1);
    return VO_OK;
}

}  // namespace vo
