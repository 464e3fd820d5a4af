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

#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "orr_internal.h"

namespace {

constexpr int BM = 128;                       // queries (and corpus rows) staged per CTA per k-block
constexpr int UN = ORR_BATCH_TILE;            // corpus rows per unit = UMMA N = accumulator columns (256)
constexpr int BK = 64;                        // bf16 per k-block (128 B)
constexpr int PLANE_BYTES = BM * BK * 2;      // 16 KB
constexpr int TMEM_COLS = 2 * UN;             // two accumulator buffers = all 512 columns
constexpr int BATCH_THREADS = 320;            // TMA producer, MMA issuer, 8 epilogue warps

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, 128-byte swizzle: 8-row groups 1024 B apart (SBO), version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
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
    int32_t  n_qblocks;          // padded batch / 256
    int32_t  k_blocks;           // dim / 64
    int64_t  rows;               // rows in the shard
    const float2* rowaux;        // [rows padded to 256] {w_cos * inv|e|, w_rec * rec or -inf}
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

// ---- cluster / cta_group::2 helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {   // same offset in CTA `cta`'s smem
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// each CTA of the pair loads its own half of the operands; completion is counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once) on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// Keyword side of one thread's half unit (4 chunks x 32 rows): the query's <= 16 term bitmaps (bit j of
// word w = row 32w+j holds the term; one 16-byte load covers the 4 chunks) are added bit-sliced into 5
// count planes per chunk.  Issued BEFORE the accumulator is awaited so the loads overlap the MMAs.
struct KwPlanes { uint32_t p[UN / 64][5]; };
__device__ __forceinline__ void kw_planes(const BatchArgs& a, int b, int64_t word0, KwPlanes& kp) {
#pragma unroll
    for (int c = 0; c < UN / 64; ++c)
#pragma unroll
        for (int i = 0; i < 5; ++i) kp.p[c][i] = 0u;
    const int4* ip = reinterpret_cast<const int4*>(a.q_term_ids + (int64_t)b * ORR_BATCH_MAX_TERMS);
#pragma unroll 1
    for (int g = 0; g < ORR_BATCH_MAX_TERMS / 4; ++g) {
        const int4 id4 = __ldg(ip + g);
        if (id4.x < 0) break;                                                // ids are packed front to back
        const int id[4] = {id4.x, id4.y, id4.z, id4.w};
        uint4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = id[i] >= 0 ? __ldg(reinterpret_cast<const uint4*>(a.term_bits + (int64_t)id[i] * a.row_words + word0))
                              : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t ww[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
            for (int c = 0; c < UN / 64; ++c) {
                uint32_t x = ww[c], cy;
                cy = kp.p[c][0] & x; kp.p[c][0] ^= x; x = cy;
                cy = kp.p[c][1] & x; kp.p[c][1] ^= x; x = cy;
                cy = kp.p[c][2] & x; kp.p[c][2] ^= x; x = cy;
                cy = kp.p[c][3] & x; kp.p[c][3] ^= x; x = cy;
                kp.p[c][4] ^= x;
            }
        }
    }
}

// Epilogue of one thread (= one query) over its UN/64 chunks of 32 accumulator columns.  Written for a
// scheduler that holds only two epilogue warps: every stage is 32 independent chains, and the rare
// work (a candidate above the threshold) sits behind ONE branch per chunk.
template <int MODE>
__device__ __forceinline__ void epilogue_chunks(const BatchArgs& a, uint32_t taddr, const float2* ax, int b, int64_t row0,
                                                int t, int c_begin, float qs, float thr, float kww, const KwPlanes& kp) {
#pragma unroll
    for (int cc = 0; cc < UN / 64; ++cc) {
        const int c = c_begin + cc;
        uint32_t acc[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), acc);
        float s[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float2 ra = ax[c * 32 + j];                               // smem broadcast
            s[j] = fmaf(__uint_as_float(acc[j]) * qs, ra.x, ra.y);
        }
        if (kww != 0.f) {
            const uint32_t c0 = kp.p[cc][0], c1 = kp.p[cc][1], c2 = kp.p[cc][2], c3 = kp.p[cc][3], c4 = kp.p[cc][4];
            if ((c0 | c1 | c2 | c3 | c4) != 0u) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t cnt = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2) |
                                         (((c3 >> j) & 1u) << 3) | (((c4 >> j) & 1u) << 4);
                    s[j] = fmaf(kww, (float)cnt, s[j]);
                }
            }
        }
        if (MODE == 0) {
            bool any = false;
#pragma unroll
            for (int j = 0; j < 32; ++j) any |= (s[j] > thr);
            if (any) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (s[j] > thr) {
                        const uint32_t slot = atomicAdd(a.cand_count + b, 1u);
                        if (slot < (uint32_t)a.cand_cap)
                            a.cand[(int64_t)b * a.cand_cap + slot] =
                                make_uint2((uint32_t)(row0 + c * 32 + j), __float_as_uint(s[j]));
                    }
                }
            }
        } else {
            float4* dst = reinterpret_cast<float4*>(a.dense + (int64_t)b * a.dense_ld + ((int64_t)t * UN + c * 32));
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
        }
    }
}

// Unit = 256 queries x 256 corpus rows on a CTA PAIR (cta_group::2, UMMA 256x256x16): each CTA stages
// its own 128 queries (A half) and 128 rows (B half), so a k-block costs every SM 64 KB of L2->smem
// traffic for 2x the MMA work of a 128x128 single-CTA unit.  PASSES = 3: split precision
// (hi.hi + hi.mid + mid.hi); PASSES = 1: bf16 screen only (wider selection margin, see orr_api.cu).
template <int PASSES> struct GemmCfg {
    static constexpr int PLANES = PASSES == 3 ? 4 : 2;
    static constexpr int STAGE = PLANES * PLANE_BYTES;          // 64 KB / 32 KB per CTA
    static constexpr int NSTAGE = PASSES == 3 ? 3 : 6;
    static constexpr int SMEM = NSTAGE * STAGE + 2 * UN * 8 + 256 + 1024;
};
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

template <int PASSES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BATCH_THREADS, 1)
orr_batch_gemm_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qmid,
                      const __grid_constant__ CUtensorMap map_ehi, const __grid_constant__ CUtensorMap map_emid,
                      const BatchArgs a) {
    using C = GemmCfg<PASSES>;
    extern __shared__ uint8_t smem_raw[];
    // the dynamic smem base is only 16-B aligned by contract; both CTAs compute the same offset
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_mem = smem;                                              // NSTAGE x STAGE, 1024-B aligned
    float2* aux = reinterpret_cast<float2*>(smem + C::NSTAGE * C::STAGE);   // 2 x UN
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE + 2 * UN * 8);
    uint64_t* full_bar = bars;                          // [NSTAGE]  leader's copy is the live one
    uint64_t* empty_bar = bars + C::NSTAGE;             // [NSTAGE]  per CTA (multicast commit)
    uint64_t* tfull_bar = bars + 2 * C::NSTAGE;         // [2]       per CTA (multicast commit)
    uint64_t* tempty_bar = bars + 2 * C::NSTAGE + 2;    // [2]       leader's copy: 16 epilogue warps of the pair
    uint64_t* aux_full = bars + 2 * C::NSTAGE + 4;      // [2]       per CTA
    uint64_t* aux_empty = bars + 2 * C::NSTAGE + 6;     // [2]       per CTA: 8 local epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&tfull_bar[b]), 1);
            mbar_init(smem_u32(&tempty_bar[b]), 16);
            mbar_init(smem_u32(&aux_full[b]), 1);
            mbar_init(smem_u32(&aux_empty[b]), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // barriers of BOTH CTAs initialised, TMEM allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int n_clusters = (int)gridDim.x >> 1, cid = (int)blockIdx.x >> 1;
    const int units_per_tile = a.n_qblocks;
    const int my_tiles = (a.n_row_tiles > cid) ? (a.n_row_tiles - 1 - cid) / n_clusters + 1 : 0;
    const int my_units = my_tiles * units_per_tile;

    if (warp == 0) {
        // ===================== TMA producer (one per CTA) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = 0; u < my_units; ++u) {
                const int t = cid + (u / units_per_tile) * n_clusters;
                const int row_tile = t * a.row_tile_stride;
                const int qb = u % units_per_tile;
                const int buf = u & 1;
                mbar_wait(smem_u32(&aux_empty[buf]), ((uint32_t)(u >> 1) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&aux_full[buf]), UN * 8);
                bulk_g2s(smem_u32(aux + buf * UN), a.rowaux + (int64_t)row_tile * UN, UN * 8, smem_u32(&aux_full[buf]));
                const int qrow = qb * 256 + (int)rank * BM, erow = row_tile * UN + (int)rank * BM;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t fb_local = smem_u32(&full_bar[stage]);
                    const uint32_t fb = mapa_shared(fb_local, 0);
                    const uint32_t base = smem_u32(stage_mem + stage * C::STAGE);
                    if (leader) mbar_expect_tx(fb_local, 2 * C::STAGE);     // both halves land on this barrier
                    if (PASSES == 3) {
                        tma_load_2d_pair(base + 0 * PLANE_BYTES, &map_qhi, kb * BK, qrow, fb);
                        tma_load_2d_pair(base + 1 * PLANE_BYTES, &map_qmid, kb * BK, qrow, fb);
                        tma_load_2d_pair(base + 2 * PLANE_BYTES, &map_ehi, kb * BK, erow, fb);
                        tma_load_2d_pair(base + 3 * PLANE_BYTES, &map_emid, kb * BK, erow, fb);
                    } else {
                        tma_load_2d_pair(base + 0 * PLANE_BYTES, &map_qhi, kb * BK, qrow, fb);
                        tma_load_2d_pair(base + 1 * PLANE_BYTES, &map_ehi, kb * BK, erow, fb);
                    }
                    if (++stage == C::NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            int stage = 0; uint32_t phase = 0;
            for (int u = 0; u < my_units; ++u) {
                const int buf = u & 1;
                mbar_wait(smem_u32(&tempty_bar[buf]), ((uint32_t)(u >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * UN);
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t base = smem_u32(stage_mem + stage * C::STAGE);
                        if (PASSES == 3) {
                            const uint64_t qhi = umma_desc(base), qmid = umma_desc(base + PLANE_BYTES);
                            const uint64_t ehi = umma_desc(base + 2 * PLANE_BYTES), emid = umma_desc(base + 3 * PLANE_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);      // 16 bf16 = 32 B = 2 x 16-B units
                                umma_bf16_pair(tmem_d, qhi + adv, ehi + adv, IDESC2, (kb | k) ? 1u : 0u);
                                umma_bf16_pair(tmem_d, qhi + adv, emid + adv, IDESC2, 1u);
                                umma_bf16_pair(tmem_d, qmid + adv, ehi + adv, IDESC2, 1u);
                            }
                        } else {
                            const uint64_t qhi = umma_desc(base), ehi = umma_desc(base + PLANE_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                umma_bf16_pair(tmem_d, qhi + adv, ehi + adv, IDESC2, (kb | k) ? 1u : 0u);
                            }
                        }
                        umma_commit_pair(smem_u32(&empty_bar[stage]));       // frees the stage in both CTAs
                        if (kb == a.k_blocks - 1) umma_commit_pair(smem_u32(&tfull_bar[buf]));
                    }
                    __syncwarp();
                    if (++stage == C::NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of each CTA: its 128 queries x 256 rows) =====================
        // A warp may only read its own TMEM lane quarter (warp % 4); the two warps that share a
        // quarter split the unit's 8 column chunks.  One thread = one query.
        const int quarter = warp & 3;
        const int c_begin = ((warp - 2) >> 2) * (UN / 64);
        const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
        const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
        for (int u = 0; u < my_units; ++u) {
            const int t = cid + (u / units_per_tile) * n_clusters;
            const int row_tile = t * a.row_tile_stride;
            const int qb = u % units_per_tile;
            const int buf = u & 1;
            const int b = qb * 256 + (int)rank * BM + quarter * 32 + lane;   // this thread's query
            const float qs = a.qscale[b];
            const float thr = a.mode == 0 ? a.thr[b] : 0.f;
            const float kww = a.q_kw_w ? a.q_kw_w[b] : 0.f;
            const int64_t row0 = (int64_t)row_tile * UN;
            KwPlanes kp;
            if (kww != 0.f) kw_planes(a, b, (row0 >> 5) + c_begin, kp);
            mbar_wait(smem_u32(&aux_full[buf]), (uint32_t)(u >> 1) & 1u);
            mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)(u >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float2* ax = aux + buf * UN;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * UN);
            if (a.mode == 0)
                epilogue_chunks<0>(a, taddr, ax, b, row0, t, c_begin, qs, thr, kww, kp);
            else
                epilogue_chunks<1>(a, taddr, ax, b, row0, t, c_begin, qs, thr, kww, kp);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&aux_empty[buf]));
                mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
            }
        }
    }
    // no CTA may leave while its peer can still signal its barriers or read its smem
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
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

template <int PASSES>
static int launch_gemm_t(const CUtensorMap& mqh, const CUtensorMap& mqm, const CUtensorMap& meh, const CUtensorMap& mem,
                         const BatchArgs& a, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        ORR_CUDA_OK(cudaFuncSetAttribute(orr_batch_gemm_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmCfg<PASSES>::SMEM));
        configured = true;
    }
    orr_batch_gemm_kernel<PASSES><<<grid, BATCH_THREADS, GemmCfg<PASSES>::SMEM, st>>>(mqh, mqm, meh, mem, a);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_launch_gemm(const OrrBatchGemm& g, cudaStream_t st) {
    if (g.dim % BK != 0) { orr_set_error("batch path needs dim %% 64 == 0 (dim=%d)", g.dim); return ORR_E_UNSUPPORTED; }
    if (g.batch_padded % 256 != 0) { orr_set_error("batch path: padded batch %d not a multiple of 256", g.batch_padded); return ORR_E_INTERNAL; }
    CUtensorMap mqh, mqm, meh, mem;
    int rc;
    if ((rc = make_plane_map(&mqh, g.qhi, g.batch_padded, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&mqm, g.qmid, g.batch_padded, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&meh, g.ehi, g.rows, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&mem, g.emid, g.rows, g.dim, BM)) != ORR_OK) return rc;
    BatchArgs a{};
    const int64_t all_tiles = (g.rows + UN - 1) / UN;
    a.row_tile_stride = g.tile_stride < 1 ? 1 : g.tile_stride;
    a.n_row_tiles = (int32_t)((all_tiles + a.row_tile_stride - 1) / a.row_tile_stride);
    a.n_qblocks = g.batch_padded / 256;
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
    const int pairs = std::min<int64_t>(g.sms / 2, a.n_row_tiles);
    if (pairs < 1) return ORR_OK;
    return g.passes == 1 ? launch_gemm_t<1>(mqh, mqm, meh, mem, a, 2 * pairs, st)
                         : launch_gemm_t<3>(mqh, mqm, meh, mem, a, 2 * pairs, st);
}
