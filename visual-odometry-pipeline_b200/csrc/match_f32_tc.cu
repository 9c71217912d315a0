// Float-descriptor matcher on the 5th-generation tensor cores (VO_PREC_TF32X3 / VO_PREC_TF32X1).
//
// Replaces `sim = d1 @ d2.t(); topk(sim, 2, dim=1); max(sim, dim=0)` of the reference's torch matchers
// (R2D2.py:56-60: cuBLAS SGEMM + two more passes over a materialised N x M matrix) and the brute-force
// distance matrix of cv2.BFMatcher.knnMatch (feature_extractors/SIFT.py:27).  Here the N x M similarity
// matrix lives only in tensor memory:
//
//   * one CTA owns a 128-row block of the reference descriptors (A, resident in shared memory for the whole
//     CTA lifetime, loaded once by TMA) and streams 128-column tiles of the current-frame descriptors (B)
//     through a 5-stage TMA/mbarrier ring;
//   * a single thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) into one of two
//     128-column fp32 accumulators in TMEM.  3xTF32: operands are pre-split into tf32 hi/lo parts and every
//     k-step issues lo*hi, hi*hi and hi*lo, which restores fp32-grade products; for integer-valued SIFT
//     descriptors one pass is already exact (every product and partial sum is an integer < 2^24);
//   * two groups of four epilogue warps (one group per accumulator) read the tile with tcgen05.ld and fold it
//     into a running top-2 per row (thread-private, one row per thread) and a per-column arg-max
//     (redux.sync.max.f32 across the 32 rows of a warp, 4 warps merged in shared memory, one 64-bit
//     atomicMin per column and CTA), while the MMA thread already fills the other accumulator.
//
// L2 mode uses the GEMM form -|a_i - b_j|^2 = (2 a_i.b_j - |b_j|^2) - |a_i|^2, evaluated per element in the
// epilogue (the row norm matters for the column arg-min, the column norm for the row arg-min).
#include "common.cuh"
#include <cuda.h>  // CUtensorMap & enums only; cuTensorMapEncodeTiled is resolved at run time

namespace vo {
namespace {

constexpr int TC_BM = 128;      // rows of A per CTA (= TMEM lanes)
constexpr int TC_BN = 128;      // columns per B tile (= accumulator columns)
constexpr int TC_D = 128;       // descriptor length
constexpr int TC_KB = 32;       // k elements per swizzle-128B row (32 fp32 = 128 B)
constexpr int TC_NKB = TC_D / TC_KB;
constexpr int TC_STAGES = 5;
constexpr int TC_BLOCK_BYTES = TC_BN * 128;  // one [128 rows x 32 k] fp32 box = 16 KB
constexpr int TC_THREADS = 320;              // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_TMEM_COLS = 256;            // 2 accumulators x 128 columns

struct TcSmem {
    // offsets into the 1024-aligned dynamic shared memory
    static constexpr int a_hi = 0;
    static constexpr int a_lo = a_hi + TC_NKB * TC_BLOCK_BYTES;
    static __host__ __device__ constexpr int b_stages(int passes) { return passes == 3 ? a_lo + TC_NKB * TC_BLOCK_BYTES : a_lo; }
    static __host__ __device__ constexpr int scol(int passes) { return b_stages(passes) + TC_STAGES * TC_BLOCK_BYTES; }
    static __host__ __device__ constexpr int bars(int passes) { return scol(passes) + 2 * 4 * TC_BN * 8; }
    static __host__ __device__ constexpr int total(int passes) { return bars(passes) + 256 + 1024 /* alignment slack */; }
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (sticky error the host reports), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float warp_max_f32(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B (8 rows x 128 B)
// | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// A,B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// ---------------------------------------------------------------- pre-pass: tf32 split + squared norms
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// one warp per descriptor row (128 floats = 32 lanes x float4)
__global__ void __launch_bounds__(256)
prep_kernel(const float *__restrict__ x, long long rows, float *__restrict__ hi, float *__restrict__ lo,
            float *__restrict__ norm2) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4 v = reinterpret_cast<const float4 *>(x)[row * 32 + lane];
    float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    reinterpret_cast<float4 *>(hi)[row * 32 + lane] = h;
    if (lo) {
        float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
        reinterpret_cast<float4 *>(lo)[row * 32 + lane] = l;
    }
    if (norm2) {
        float s = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fadd_rn(__fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (lane == 0) norm2[row] = s;
    }
}

// ---------------------------------------------------------------- main kernel
template <int PASSES, int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
match_f32_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    int n_stride, int m_stride, const int32_t *__restrict__ n_ref, const int32_t *__restrict__ n_cur,
                    const float *__restrict__ row_norm, const float *__restrict__ col_norm, int n_split,
                    vo_row_partial *__restrict__ part, unsigned long long *__restrict__ colkey) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - smem_u32(smem_raw));

    const int b = blockIdx.z, split = blockIdx.y;
    const int N = n_ref ? min(n_ref[b], n_stride) : n_stride;
    const int M = n_cur ? min(n_cur[b], m_stride) : m_stride;
    const int row0 = blockIdx.x * TC_BM;
    const int tiles_total = (M + TC_BN - 1) / TC_BN;
    const int tiles_per_split = (tiles_total + n_split - 1) / n_split;
    const int t_begin = min(tiles_total, split * tiles_per_split);
    const int t_end = min(tiles_total, t_begin + tiles_per_split);
    const int n_tiles = t_end - t_begin;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    constexpr int ITEMS = (PASSES == 3) ? 2 * TC_NKB : TC_NKB;  // B boxes streamed per tile
    const uint32_t s_a_hi = base + TcSmem::a_hi, s_a_lo = base + TcSmem::a_lo;
    const uint32_t s_b = base + TcSmem::b_stages(PASSES);
    float *scol_v = reinterpret_cast<float *>(smem + TcSmem::scol(PASSES));          // [2][4][128]
    int *scol_r = reinterpret_cast<int *>(smem + TcSmem::scol(PASSES) + 2 * 4 * TC_BN * 4);
    const uint32_t s_bar = base + TcSmem::bars(PASSES);
    // barrier slots (8 B each): full[5] empty[5] a_full tmem_full[2] tmem_empty[2]; then the TMEM base word
    auto bar_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_empty = [&](int s) { return s_bar + 8u * (TC_STAGES + s); };
    const uint32_t bar_a = s_bar + 8u * (2 * TC_STAGES);
    auto bar_tfull = [&](int g) { return s_bar + 8u * (2 * TC_STAGES + 1 + g); };
    auto bar_tempty = [&](int g) { return s_bar + 8u * (2 * TC_STAGES + 3 + g); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::bars(PASSES) + 8 * (2 * TC_STAGES + 5));

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_a, 1);
        for (int g = 0; g < 2; ++g) { mbar_init(bar_tfull(g), 1); mbar_init(bar_tempty(g), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-collective; this warp also frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0 && n_tiles > 0) {
            const int arow = b * n_stride + row0;
            mbar_expect_tx(bar_a, (PASSES == 3 ? 2 : 1) * TC_NKB * TC_BLOCK_BYTES);
            for (int kb = 0; kb < TC_NKB; ++kb) {
                tma_load_2d(s_a_hi + kb * TC_BLOCK_BYTES, &map_a_hi, kb * TC_KB, arow, bar_a);
                if (PASSES == 3) tma_load_2d(s_a_lo + kb * TC_BLOCK_BYTES, &map_a_lo, kb * TC_KB, arow, bar_a);
            }
            int it = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int brow = b * m_stride + t * TC_BN;
                for (int item = 0; item < ITEMS; ++item, ++it) {
                    const int stage = it % TC_STAGES;
                    const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(bar_empty(stage), phase ^ 1u);
                    mbar_expect_tx(bar_full(stage), TC_BLOCK_BYTES);
                    const int kb = (PASSES == 3) ? (item >> 1) : item;
                    const bool is_lo = (PASSES == 3) && (item & 1);
                    tma_load_2d(s_b + stage * TC_BLOCK_BYTES, is_lo ? &map_b_lo : &map_b_hi, kb * TC_KB, brow,
                                bar_full(stage));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0 && n_tiles > 0) {
            mbar_wait(bar_a, 0);
            tc_fence_after();
            int it = 0;
            for (int lt = 0; lt < n_tiles; ++lt) {
                const int buf = lt & 1;
                const uint32_t use = (uint32_t)(lt >> 1);
                mbar_wait(bar_tempty(buf), (use & 1u) ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TC_BN);
                uint32_t acc = 0;
                for (int item = 0; item < ITEMS; ++item, ++it) {
                    const int stage = it % TC_STAGES;
                    const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
                    mbar_wait(bar_full(stage), phase);
                    tc_fence_after();
                    const int kb = (PASSES == 3) ? (item >> 1) : item;
                    const bool is_lo = (PASSES == 3) && (item & 1);
                    const uint32_t sb = s_b + stage * TC_BLOCK_BYTES;
#pragma unroll
                    for (int k8 = 0; k8 < TC_KB / 8; ++k8) {
                        const uint64_t bdesc = make_sdesc(sb + k8 * 32);
                        const uint64_t ahi = make_sdesc(s_a_hi + kb * TC_BLOCK_BYTES + k8 * 32);
                        if (is_lo) {  // a_hi * b_lo
                            tc_mma_tf32(d_tmem, ahi, bdesc, TC_IDESC, acc);
                            acc = 1;
                        } else {
                            if (PASSES == 3) {  // a_lo * b_hi first (small term), then a_hi * b_hi
                                const uint64_t alo = make_sdesc(s_a_lo + kb * TC_BLOCK_BYTES + k8 * 32);
                                tc_mma_tf32(d_tmem, alo, bdesc, TC_IDESC, acc);
                                acc = 1;
                            }
                            tc_mma_tf32(d_tmem, ahi, bdesc, TC_IDESC, acc);
                            acc = 1;
                        }
                    }
                    tc_commit(bar_empty(stage));  // smem slot reusable once these MMAs retire
                }
                tc_commit(bar_tfull(buf));  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: 2 groups x 4 warps =====================
        const int e = warp - 2;
        const int g = e >> 2;       // accumulator / tile parity served by this group
        const int q = warp & 3;     // TMEM lane quarter this warp may read
        const int row = row0 + q * 32 + lane;
        const bool row_ok = row < N;
        float s1 = -INFINITY, s2 = -INFINITY;  // running top-2 of -|a-b|^2 or a.b: larger is better
        int32_t i1 = -1, i2 = -1;
        float *my_cv = scol_v + (g * 4 + q) * TC_BN;
        int *my_cr = scol_r + (g * 4 + q) * TC_BN;
        const float *cn = (METRIC == VO_METRIC_L2) ? col_norm + (size_t)b * m_stride : nullptr;
        const float na = (METRIC == VO_METRIC_L2 && row_ok) ? row_norm[(size_t)b * n_stride + row] : 0.0f;

        for (int lt = g; lt < n_tiles; lt += 2) {
            const int col0 = (t_begin + lt) * TC_BN;
            const uint32_t use = (uint32_t)(lt >> 1);
            mbar_wait(bar_tfull(g), use & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * TC_BN);
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                float v[32];
                tc_ld32(taddr + c * 32, v);
                if (c == TC_BN / 32 - 1) {  // whole accumulator is in registers / consumed: hand it back
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(bar_tempty(g));
                }
                const int cbase = col0 + c * 32;
                float cv = -INFINITY;
                int cr = 0;
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    float sc[8];
                    float m8 = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = cbase + j0 + j;
                        float s = v[j0 + j];
                        if (METRIC == VO_METRIC_L2)
                            s = __fsub_rn(__fmaf_rn(2.0f, s, -__ldg(cn + min(col, m_stride - 1))), na);
                        s = (col < M) ? s : -INFINITY;
                        sc[j] = s;
                        m8 = fmaxf(m8, s);
                        // column arg-max over the 32 rows of this warp (lowest row on ties)
                        const float sr = row_ok ? s : -INFINITY;
                        const float wm = warp_max_f32(sr);
                        const unsigned bal = __ballot_sync(0xffffffffu, sr == wm);
                        if (lane == j0 + j) { cv = wm; cr = __ffs(bal) - 1; }
                    }
                    if (row_ok && m8 > s2) {  // rare after the first tiles: sequential update keeps lowest index on ties
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float s = sc[j];
                            const int col = cbase + j0 + j;
                            if (s > s1) {
                                s2 = s1; i2 = i1; s1 = s; i1 = col;
                            } else if (s > s2) {
                                s2 = s; i2 = col;
                            }
                        }
                    }
                }
                my_cv[c * 32 + lane] = cv;
                my_cr[c * 32 + lane] = cr;
            }
            group_bar(1 + g);
            {   // 128 threads of the group: one column each, merge the 4 lane quarters (ascending rows)
                const int j = (e & 3) * 32 + lane;
                const int col = col0 + j;
                float best = -INFINITY;
                int brow = -1;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const float val = scol_v[(g * 4 + qq) * TC_BN + j];
                    const int r = scol_r[(g * 4 + qq) * TC_BN + j];
                    if (r >= 0 && val > best) { best = val; brow = qq * 32 + r; }
                }
                if (col < M && brow >= 0 && row0 + brow < N) {
                    const unsigned long long key =
                        ((unsigned long long)float_to_ordered(-best) << 32) | (unsigned long long)(uint32_t)(row0 + brow);
                    unsigned long long *dst = colkey + (size_t)b * m_stride + col;
                    if (key < *reinterpret_cast<volatile unsigned long long *>(dst)) atomicMin(dst, key);
                }
            }
            group_bar(1 + g);
        }
        // the two groups saw disjoint tiles: each writes its own partial (finalize merges tie-aware)
        if (row < n_stride) {
            vo_row_partial p;
            p.s1 = float_to_ordered(-s1); p.s2 = float_to_ordered(-s2);
            p.i1 = i1; p.i2 = i2;
            part[((size_t)b * (n_split * 2) + split * 2 + g) * n_stride + row] = p;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(vo_ctx *ctx, CUtensorMap *map, const float *ptr, long long rows) {
    PFN_encodeTiled fn = (PFN_encodeTiled)ctx->encode_tiled;
    cuuint64_t dims[2] = {(cuuint64_t)TC_D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_D * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return VO_ERR_CUDA;
    }
    return VO_OK;
}

template <int PASSES, int METRIC>
int launch_tc(vo_ctx *ctx, dim3 grid, const CUtensorMap &ah, const CUtensorMap &al, const CUtensorMap &bh,
              const CUtensorMap &bl, int n_stride, int m_stride, const int32_t *n_ref, const int32_t *n_cur,
              const float *row_norm, const float *col_norm, int n_split, vo_row_partial *part,
              unsigned long long *colkey, cudaStream_t st) {
    auto kern = match_f32_tc_kernel<PASSES, METRIC>;
    const int smem = TcSmem::total(PASSES);
    VO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, TC_THREADS, smem, st>>>(ah, al, bh, bl, n_stride, m_stride, n_ref, n_cur, row_norm, col_norm, n_split, part,
                                         colkey);
    VO_LAUNCH_CHECK(ctx);
    return VO_OK;
}

}  // namespace

int match_f32_tc(vo_ctx *ctx, const float *ref, const float *cur, int B, int n_stride, int m_stride,
                 const int32_t *n_ref, const int32_t *n_cur, int metric, int passes, vo_row_partial **part_out,
                 int *n_split_out, unsigned long long *colkey, const float **row_norm_out, cudaStream_t st) {
    if (!ctx->tc_ready) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return VO_ERR_UNSUPPORTED;
        }
        ctx->encode_tiled = fn;
        ctx->tc_ready = 1;
    }
    const long long rows_a = (long long)B * n_stride, rows_b = (long long)B * m_stride;
    const bool l2 = metric == VO_METRIC_L2;
    // workspace: A_hi | A_lo | B_hi | B_lo  and  row norms | column norms
    float *split_a, *split_b, *norms;
    int rc;
    const size_t per_a = (size_t)rows_a * TC_D * sizeof(float), per_b = (size_t)rows_b * TC_D * sizeof(float);
    if ((rc = ws_get(ctx, WS_SPLIT_A, per_a * (passes == 3 ? 2 : 1), (void **)&split_a))) return rc;
    if ((rc = ws_get(ctx, WS_SPLIT_B, per_b * (passes == 3 ? 2 : 1), (void **)&split_b))) return rc;
    if ((rc = ws_get(ctx, WS_NORMS, sizeof(float) * (size_t)(rows_a + rows_b), (void **)&norms))) return rc;
    float *a_hi = split_a, *a_lo = passes == 3 ? split_a + (size_t)rows_a * TC_D : nullptr;
    float *b_hi = split_b, *b_lo = passes == 3 ? split_b + (size_t)rows_b * TC_D : nullptr;
    float *row_norm = norms, *col_norm = norms + rows_a;

    VO_PROF(ctx, st, VO_STAGE_PREP);
    prep_kernel<<<(unsigned)((rows_a + 7) / 8), 256, 0, st>>>(ref, rows_a, a_hi, a_lo, l2 ? row_norm : nullptr);
    VO_LAUNCH_CHECK(ctx);
    prep_kernel<<<(unsigned)((rows_b + 7) / 8), 256, 0, st>>>(cur, rows_b, b_hi, b_lo, l2 ? col_norm : nullptr);
    VO_LAUNCH_CHECK(ctx);

    CUtensorMap mah, mal, mbh, mbl;
    if ((rc = make_map(ctx, &mah, a_hi, rows_a))) return rc;
    if ((rc = make_map(ctx, &mbh, b_hi, rows_b))) return rc;
    if ((rc = make_map(ctx, &mal, passes == 3 ? a_lo : a_hi, rows_a))) return rc;
    if ((rc = make_map(ctx, &mbl, passes == 3 ? b_lo : b_hi, rows_b))) return rc;

    const int row_blocks = ceil_div(n_stride, TC_BM);
    const int n_split = pick_split(ctx, B, row_blocks, ceil_div(m_stride, TC_BN), 4);
    vo_row_partial *part;
    if ((rc = ws_get(ctx, WS_ROWPART, sizeof(vo_row_partial) * (size_t)B * n_split * 2 * n_stride, (void **)&part))) return rc;
    dim3 grid(row_blocks, n_split, B);
    VO_PROF(ctx, st, VO_STAGE_MATCH);
    if (passes == 3) {
        rc = l2 ? launch_tc<3, VO_METRIC_L2>(ctx, grid, mah, mal, mbh, mbl, n_stride, m_stride, n_ref, n_cur, row_norm, col_norm, n_split, part, colkey, st)
                : launch_tc<3, VO_METRIC_COSINE>(ctx, grid, mah, mal, mbh, mbl, n_stride, m_stride, n_ref, n_cur, row_norm, col_norm, n_split, part, colkey, st);
    } else {
        rc = l2 ? launch_tc<1, VO_METRIC_L2>(ctx, grid, mah, mal, mbh, mbl, n_stride, m_stride, n_ref, n_cur, row_norm, col_norm, n_split, part, colkey, st)
                : launch_tc<1, VO_METRIC_COSINE>(ctx, grid, mah, mal, mbh, mbl, n_stride, m_stride, n_ref, n_cur, row_norm, col_norm, n_split, part, colkey, st);
    }
    if (rc) return rc;
    *part_out = part;
    *n_split_out = n_split * 2;  // two epilogue groups -> two partials per (row, split)
    *row_norm_out = nullptr;  // scores already carry -|a-b|^2 in full
    return VO_OK;
}

}  // namespace vo
