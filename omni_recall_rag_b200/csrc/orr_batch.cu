// orr_batch.cu — batched queries: split-precision tcgen05 contraction + fused candidate
// selection (K3/K4 of SURVEY.md §2.1).
//
// For a batch of B queries the scoring loop of RecallSearchService.SearchAsync
// (src/OmniRecall.Api/Services/RecallSearchService.cs:28-33) is a contraction
// S[b,n] = q_b . e_n over the whole store — tensor-core work (2*N*D*B flop against 4*N*D
// bytes).  The reference multiplies in fp32; tensor cores take bf16, so both operands are
// split into bf16 planes x = hi + mid (+ lo, dropped) and three MMAs per k-step
//     hi.hi + hi.mid + mid.hi
// accumulate in fp32 in TMEM: relative error <= 3 * 2^-16 of sum|q_i e_i|, i.e. fp32-grade for
// SELECTION; the survivors are then re-scored with the reference's exact fp64 arithmetic
// (exact_row, orr_rescore.cu) and the same bound check as the single-query path proves the
// selection safe.
//
// Kernel shape (sm_100a, cta_group::1):
//   unit      = 128 queries (UMMA M = TMEM lanes) x 128 corpus rows (UMMA N = TMEM columns)
//   k-block   = 64 bf16 (one 128-byte swizzle row); 4 planes/stage (Q_hi,Q_mid,E_hi,E_mid) = 64 KB,
//               3 stages, filled by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) on mbarriers
//   warp 0    TMA producer        warp 1  TMEM alloc + single-thread tcgen05.mma issue
//   warps 2-5 epilogue: tcgen05.ld of the accumulator (one query per thread), fused score
//             acc * inv|q| * inv|e| * w_cos + w_rec*rec (+ keyword), compare with the query's
//             threshold, append (row, score) to the query's candidate list
//   two TMEM accumulator buffers so the epilogue of unit u overlaps the MMAs of unit u+1
//   persistent grid; a CTA walks row tiles and, inside a row tile, all query blocks, so the
//   E tile is re-read from L2, not HBM.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cfloat>
#include <cstdlib>

#include "orr_internal.h"

namespace {

constexpr int BM = 128;                       // queries per unit
constexpr int BN = 128;                       // corpus rows per unit
constexpr int BK = 64;                        // bf16 per k-block (128 B)
constexpr int STAGES = 3;
constexpr int PLANE_BYTES = BM * BK * 2;      // 16 KB
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;  // 64 KB
constexpr int AUX_BYTES = BN * 8;             // float2 per corpus row of the unit
constexpr int TMEM_COLS = 2 * BN;             // two accumulator buffers
constexpr int BATCH_THREADS = 192;            // 6 warps

constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, 128-byte swizzle: 8-row groups 1024 B apart (SBO), version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int ORR_BATCH_MAX_TERMS = ORR_BATCH_TERMS;

struct BatchArgs {
    int32_t  n_row_tiles;        // row tiles this launch walks
    int32_t  row_tile_stride;    // 1 = every tile; >1 = sampling pass (tile t -> t * stride)
    int32_t  n_qblocks;          // padded batch / 128
    int32_t  k_blocks;           // dim / 64
    int64_t  rows;               // rows in the shard
    const float2* rowaux;        // [rows padded to 128] {w_cos * inv|e|, w_rec * rec or -inf}
    const float*  qscale;        // [B padded] inv|q| (0 for padding / zero queries)
    const float*  thr;           // [B padded] candidate threshold (main pass)
    // main pass output
    uint2*    cand;              // [B][cand_cap] (row, score bits)
    uint32_t* cand_count;        // [B]
    int32_t   cand_cap;
    // dense output (sampling pass / debug): scores[b][dense_ld]
    float*    dense;
    int64_t   dense_ld;
    int32_t   mode;              // 0 = main pass (threshold + append), 1 = dense store
    // keyword side: per query up to ORR_BATCH_MAX_TERMS term bitmaps over rows
    const uint32_t* term_bits;   // [n_batch_terms][row_words] bit r%32 of word r/32 = row r has the term
    int64_t   row_words;
    const int32_t* q_term_ids;   // [B padded][ORR_BATCH_MAX_TERMS] batch-term index or -1
    const float*   q_kw_w;       // [B padded] w_kw / |terms_b| (0 if none)
    int32_t   max_terms;
};

__global__ void __launch_bounds__(BATCH_THREADS, 1)
orr_batch_gemm_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qmid,
                      const __grid_constant__ CUtensorMap map_ehi, const __grid_constant__ CUtensorMap map_emid,
                      const BatchArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_mem = smem;                                         // STAGES x 64 KB, 1024-B aligned
    float2* aux = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES);   // 2 x BN
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 2 * AUX_BYTES);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // [2]
    uint64_t* aux_bar = bars + 2 * STAGES + 4;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&tfull_bar[b]), 1);
            mbar_init(smem_u32(&tempty_bar[b]), 4);
            mbar_init(smem_u32(&aux_bar[b]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int units_per_tile = a.n_qblocks;
    const int my_tiles = (a.n_row_tiles > (int)blockIdx.x) ? (a.n_row_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int my_units = my_tiles * units_per_tile;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = 0; u < my_units; ++u) {
                const int t = blockIdx.x + (u / units_per_tile) * gridDim.x;
                const int row_tile = t * a.row_tile_stride;
                const int qb = u % units_per_tile;
                const int buf = u & 1;
                // aux[buf] is free once the epilogue of unit u-2 has released its accumulator
                mbar_wait(smem_u32(&tempty_bar[buf]), ((uint32_t)(u >> 1) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&aux_bar[buf]), AUX_BYTES);
                bulk_g2s(smem_u32(aux + buf * BN), a.rowaux + (int64_t)row_tile * BN, AUX_BYTES, smem_u32(&aux_bar[buf]));
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    const uint32_t base = smem_u32(stage_mem + stage * STAGE_BYTES);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(base + 0 * PLANE_BYTES, &map_qhi, kb * BK, qb * BM, fb);
                    tma_load_2d(base + 1 * PLANE_BYTES, &map_qmid, kb * BK, qb * BM, fb);
                    tma_load_2d(base + 2 * PLANE_BYTES, &map_ehi, kb * BK, row_tile * BN, fb);
                    tma_load_2d(base + 3 * PLANE_BYTES, &map_emid, kb * BK, row_tile * BN, fb);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0; uint32_t phase = 0;
        for (int u = 0; u < my_units; ++u) {
            const int buf = u & 1;
            mbar_wait(smem_u32(&tempty_bar[buf]), ((uint32_t)(u >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
            for (int kb = 0; kb < a.k_blocks; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t base = smem_u32(stage_mem + stage * STAGE_BYTES);
                    const uint64_t qhi = umma_desc(base), qmid = umma_desc(base + PLANE_BYTES);
                    const uint64_t ehi = umma_desc(base + 2 * PLANE_BYTES), emid = umma_desc(base + 3 * PLANE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);          // 16 bf16 = 32 B = 2 x 16-B units
                        umma_bf16(tmem_d, qhi + adv, ehi + adv, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_d, qhi + adv, emid + adv, 1u);
                        umma_bf16(tmem_d, qmid + adv, ehi + adv, 1u);
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));           // frees the smem stage when the MMAs retire
                    if (kb == a.k_blocks - 1) umma_commit(smem_u32(&tfull_bar[buf]));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may read
        for (int u = 0; u < my_units; ++u) {
            const int t = blockIdx.x + (u / units_per_tile) * gridDim.x;
            const int row_tile = t * a.row_tile_stride;
            const int qb = u % units_per_tile;
            const int buf = u & 1;
            const int b = qb * BM + quarter * 32 + lane;               // this thread's query
            const float qs = a.qscale[b];
            const float thr = a.mode == 0 ? a.thr[b] : 0.f;
            const float kww = a.q_kw_w ? a.q_kw_w[b] : 0.f;
            mbar_wait(smem_u32(&aux_bar[buf]), (uint32_t)(u >> 1) & 1u);
            mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)(u >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float2* ax = aux + buf * BN;
            const int64_t row0 = (int64_t)row_tile * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c * 32), acc);
                // keyword counts of this query for the 32 rows of the chunk: the query's term
                // bitmaps (bit j = row j has the term) are added bit-sliced into 5 count planes
                uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
                if (kww != 0.f) {
                    const int64_t word = (row0 >> 5) + c;
#pragma unroll 1
                    for (int ti = 0; ti < ORR_BATCH_MAX_TERMS; ++ti) {
                        const int id = a.q_term_ids[(int64_t)b * ORR_BATCH_MAX_TERMS + ti];
                        if (id < 0) break;
                        uint32_t w = __ldg(a.term_bits + (int64_t)id * a.row_words + word), cy;
                        cy = c0 & w; c0 ^= w; w = cy;
                        cy = c1 & w; c1 ^= w; w = cy;
                        cy = c2 & w; c2 ^= w; w = cy;
                        cy = c3 & w; c3 ^= w; w = cy;
                        c4 ^= w;
                    }
                }
                const uint32_t anykw = c0 | c1 | c2 | c3 | c4;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 ra = ax[c * 32 + j];                   // smem broadcast
                    float s = __uint_as_float(acc[j]) * qs * ra.x + ra.y;
                    if ((anykw >> j) & 1u) {
                        const int cnt = (int)((c0 >> j) & 1u) + 2 * (int)((c1 >> j) & 1u) + 4 * (int)((c2 >> j) & 1u) +
                                        8 * (int)((c3 >> j) & 1u) + 16 * (int)((c4 >> j) & 1u);
                        s += kww * (float)cnt;
                    }
                    const int64_t row = row0 + c * 32 + j;
                    if (a.mode == 0) {
                        if (s > thr) {
                            const uint32_t slot = atomicAdd(a.cand_count + b, 1u);
                            if (slot < (uint32_t)a.cand_cap)
                                a.cand[(int64_t)b * a.cand_cap + slot] = make_uint2((uint32_t)row, __float_as_uint(s));
                        }
                    } else {
                        a.dense[(int64_t)b * a.dense_ld + ((int64_t)(t * BN) + c * 32 + j)] = s;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- split planes of the store: x = hi + mid (+ dropped lo), bf16 each; 1/|e| per row ------------
__global__ void __launch_bounds__(256) orr_build_planes_kernel(const float* emb, __nv_bfloat16* hi, __nv_bfloat16* mid,
                                                               float* inv_norm, int64_t first, int64_t n, int dim) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gw; i < n; i += W) {
        const int64_t row = first + i;
        const float* x = emb + row * dim;
        float ss = 0.f;
        for (int c = lane; c < dim; c += 32) {
            const float v = x[c];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 m = __float2bfloat16_rn(v - __bfloat162float(h));
            hi[row * dim + c] = h;
            mid[row * dim + c] = m;
            ss = fmaf(v, v, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) inv_norm[row] = (ss > 0.f && ss < 3e38f) ? rsqrtf(ss) : 0.f;
    }
}

// queries: planes (zero-padded to a multiple of 128 queries) and inv|q|
__global__ void __launch_bounds__(128) orr_prep_queries_kernel(const float* q, __nv_bfloat16* qhi, __nv_bfloat16* qmid,
                                                               float* qscale, int batch, int dim) {
    const int b = blockIdx.x;
    __shared__ float red[4];
    float ss = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        const float v = b < batch ? q[(int64_t)b * dim + c] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        qhi[(int64_t)b * dim + c] = h;
        qmid[(int64_t)b * dim + c] = __float2bfloat16_rn(v - __bfloat162float(h));
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float t = red[0] + red[1] + red[2] + red[3];
        qscale[b] = (t > 0.f && t < 3e38f) ? rsqrtf(t) : 0.f;
    }
}

// per-batch row side of the fused score: {w_cos/|e|, w_rec * exp(-age/30d)}; tombstones -> -inf
__global__ void orr_build_rowaux_kernel(const int64_t* ticks, const float* inv_norm, float2* rowaux, int64_t rows,
                                        int64_t rows_padded, int64_t now_ticks, float w_cos, float w_rec,
                                        float decay_per_2p20) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_padded) return;
    float2 v = make_float2(0.f, -INFINITY);
    if (i < rows) {
        const int64_t tk = ticks[i];
        if (tk != ORR_DEAD_TICKS) {
            int64_t age20 = (now_ticks - tk) >> 20;
            age20 = age20 < 0 ? 0 : (age20 > 0x7fffffffLL ? 0x7fffffffLL : age20);
            v.x = w_cos * inv_norm[i];
            v.y = w_rec * __expf(-(float)(int32_t)age20 * decay_per_2p20);
        }
    }
    rowaux[i] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ORR_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) { orr_set_error("cuTensorMapEncodeTiled unavailable"); return ORR_E_CUDA; }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { orr_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ORR_E_CUDA; }
    return ORR_OK;
}

constexpr int BATCH_SMEM = STAGES * STAGE_BYTES + 2 * AUX_BYTES + 256;

}  // namespace

// ---- host-side launchers ---------------------------------------------------------------------------
int orr_batch_build_planes(const float* emb, void* hi, void* mid, float* inv_norm, int64_t first, int64_t n, int dim,
                           cudaStream_t st) {
    if (n <= 0) return ORR_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    orr_build_planes_kernel<<<sms * 8, 256, 0, st>>>(emb, (__nv_bfloat16*)hi, (__nv_bfloat16*)mid, inv_norm, first, n, dim);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_prep_queries(const float* q_dev, void* qhi, void* qmid, float* qscale, int batch, int batch_padded,
                           int dim, cudaStream_t st) {
    orr_prep_queries_kernel<<<batch_padded, 128, 0, st>>>(q_dev, (__nv_bfloat16*)qhi, (__nv_bfloat16*)qmid, qscale, batch, dim);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_build_rowaux(const int64_t* ticks, const float* inv_norm, void* rowaux, int64_t rows, int64_t rows_padded,
                           int64_t now_ticks, const OrrWeights& w, cudaStream_t st) {
    const float decay = (float)(1048576.0 / ((double)ORR_TICKS_PER_DAY * w.recency_days));
    orr_build_rowaux_kernel<<<(unsigned)((rows_padded + 255) / 256), 256, 0, st>>>(
        ticks, inv_norm, (float2*)rowaux, rows, rows_padded, now_ticks, (float)w.w_cos, (float)w.w_rec, decay);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_launch_gemm(const OrrBatchGemm& g, cudaStream_t st) {
    if (g.dim % BK != 0) { orr_set_error("batch path needs dim %% 64 == 0 (dim=%d)", g.dim); return ORR_E_UNSUPPORTED; }
    static bool configured = false;
    if (!configured) {
        ORR_CUDA_OK(cudaFuncSetAttribute(orr_batch_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BATCH_SMEM));
        configured = true;
    }
    CUtensorMap mqh, mqm, meh, mem;
    int rc;
    if ((rc = make_plane_map(&mqh, g.qhi, g.batch_padded, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&mqm, g.qmid, g.batch_padded, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&meh, g.ehi, g.rows, g.dim, BN)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&mem, g.emid, g.rows, g.dim, BN)) != ORR_OK) return rc;
    BatchArgs a{};
    const int64_t all_tiles = (g.rows + BN - 1) / BN;
    a.row_tile_stride = g.tile_stride < 1 ? 1 : g.tile_stride;
    a.n_row_tiles = (int32_t)((all_tiles + a.row_tile_stride - 1) / a.row_tile_stride);
    a.n_qblocks = g.batch_padded / BM;
    a.k_blocks = g.dim / BK;
    a.rows = g.rows;
    a.rowaux = (const float2*)g.rowaux;
    a.qscale = g.qscale;
    a.thr = g.thr;
    a.cand = (uint2*)g.cand;
    a.cand_count = g.cand_count;
    a.cand_cap = g.cand_cap;
    a.dense = g.dense;
    a.dense_ld = g.dense_ld;
    a.mode = g.dense ? 1 : 0;
    a.term_bits = g.term_bits;
    a.row_words = g.row_words;
    a.q_term_ids = g.q_term_ids;
    a.q_kw_w = g.q_kw_w;
    a.max_terms = ORR_BATCH_MAX_TERMS;
    int grid = g.sms < a.n_row_tiles ? g.sms : a.n_row_tiles;
    if (grid < 1) return ORR_OK;
    orr_batch_gemm_kernel<<<grid, BATCH_THREADS, BATCH_SMEM, st>>>(mqh, mqm, meh, mem, a);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}
