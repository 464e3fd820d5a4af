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
#include <cuda_fp16.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
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

constexpr int ORR_BATCH_MAX_TERMS = ORR_BATCH_TERMS;
constexpr int MAX_QBLOCKS = ORR_BATCH_MAX_QUERIES / ORR_BATCH_TILE;   // query blocks one launch walks per row tile (4)
constexpr uint16_t NO_TERM = 0xFFFFu;

struct BatchArgs {
    int32_t  n_row_tiles;        // row tiles this launch walks
    int32_t  row_tile_stride;    // 1 = every tile; >1 = sampling pass (tile t -> t * stride)
    int32_t  n_qblocks;          // padded batch / 256  (<= MAX_QBLOCKS)
    int32_t  k_blocks;           // dim / 64
    int64_t  rows;               // rows in the shard
    const float* rowrec;         // [rows padded to 256] w_rec * exp(-age/30d); -inf for tombstones and padding
    const float* qscale;         // [B padded] inv|q| (0 for padding / zero queries)
    const float* thr;            // [B padded] candidate threshold (main pass)
    // main pass output
    uint2*    cand;              // [B][cand_cap] (row, score bits)
    uint32_t* cand_count;        // [B]
    int32_t   cand_cap;
    // dense output (sampling pass / debug): scores[b][dense_ld]
    float*    dense;
    int64_t   dense_ld;
    int32_t   dense_half;        // dense holds __half (same buffer, ld in elements)
    // keyword side: per query up to ORR_BATCH_MAX_TERMS term bitmaps over rows
    // tile-major: term_bits[row tile][slot][8 words] — the 256 rows of one tile for every term slot are one
    // contiguous slot_cap x 32 B region, read once per tile and shared by all queries of the batch
    const uint32_t* term_bits;   // bit j of word w = row 256*tile + 32*w + j holds the term
    int64_t   slot_cap;          // term slots per row tile (the pool's capacity)
    const int32_t* q_term_ids;   // [B padded][ORR_BATCH_MAX_TERMS] term slot or -1, packed front to back
    const float*   q_kw_w;       // [B padded] w_kw / |terms_b| (0 if none)
    int32_t   planes_tiled;      // row planes in k-block-major tiles (else row-major)
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
    // default .release.cta semantics: the accumulator reads it orders are tcgen05 loads already completed by
    // tcgen05.wait::ld + tcgen05.fence::before_thread_sync; a cluster-scope release would also drain every
    // outstanding candidate store (a MEMBAR.GPU per unit and warp)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// arrives (once) on the barrier at this offset in every CTA of `mask` when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_mask(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// the same box lands at the same smem offset of every CTA in `mask`; the bytes are counted on the barrier at `bar`'s offset
// in the LEADER of each destination's pair (`bar` is this CTA's own barrier address with the pair bit, bit 24, cleared)
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu), "h"(mask) : "memory");
}

// ---- TMEM loads split into issue and wait so the load of chunk c+1 overlaps the math of chunk c ----
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// the registers are in/out operands of the wait so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

// ---- keyword side -----------------------------------------------------------------------------------
// One thread = one query; its half unit is 4 chunks x 32 rows = one 16-byte word group per term bitmap
// (bit j of word w = row 32w+j holds the term).  A query's <= 16 term words are added bit-sliced into
// 5 count planes per chunk; the epilogue then adds kww * 2^p to the rows whose plane-p bit is set.
constexpr int HC = UN / 64;                   // chunks per thread per unit (4)
struct KwPlanes { uint32_t p[HC][5]; };

__device__ __forceinline__ void kw_clear(KwPlanes& kp) {
#pragma unroll
    for (int c = 0; c < HC; ++c)
#pragma unroll
        for (int i = 0; i < 5; ++i) kp.p[c][i] = 0u;
}
template <int DEPTH>                           // planes a carry can reach (3 is enough for the first 4 terms)
__device__ __forceinline__ void kw_add(KwPlanes& kp, const uint4& w) {
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int c = 0; c < HC; ++c) {
        uint32_t x = ww[c], cy;
#pragma unroll
        for (int i = 0; i < DEPTH - 1; ++i) { cy = kp.p[c][i] & x; kp.p[c][i] ^= x; x = cy; }
        kp.p[c][DEPTH - 1] ^= x;
    }
}
// Queries with more than 4 terms: all 16 term words at once through a carry-save adder tree (Harley-Seal
// counter: 15 CSAs = 30 LOP3 per 32-row chunk for 16 one-bit inputs, against ~10 per term for the rippling
// kw_add) — absent terms are zero words.
__device__ __forceinline__ uint32_t u4c(const uint4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    l = a ^ b ^ c;
    h = (a & b) | ((a ^ b) & c);
}
__device__ __forceinline__ void kw_tree(KwPlanes& kp, const uint4 (&w0)[4], const uint4 (&w1)[12]) {
#pragma unroll
    for (int c = 0; c < HC; ++c) {
        uint32_t x[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = u4c(w0[i], c);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[4 + i] = u4c(w1[i], c);
        uint32_t ones, twos, fours, eights, sixteens, tA, tB, fA, fB, eA, eB;
        csa(tA, ones, x[0], x[1], x[2]);
        csa(tB, ones, ones, x[3], x[4]);
        csa(fA, twos, tA, tB, 0u);
        csa(tA, ones, ones, x[5], x[6]);
        csa(tB, ones, ones, x[7], x[8]);
        csa(fB, twos, twos, tA, tB);
        csa(eA, fours, fA, fB, 0u);
        csa(tA, ones, ones, x[9], x[10]);
        csa(tB, ones, ones, x[11], x[12]);
        csa(fA, twos, twos, tA, tB);
        csa(tA, ones, ones, x[13], x[14]);
        csa(tB, ones, ones, x[15], 0u);
        csa(fB, twos, twos, tA, tB);
        csa(eB, fours, fours, fA, fB);
        csa(sixteens, eights, eA, eB, 0u);
        kp.p[c][0] = ones; kp.p[c][1] = twos; kp.p[c][2] = fours; kp.p[c][3] = eights; kp.p[c][4] = sixteens;
    }
}
// `tile_half` = 2 * row tile + (0 | 1): the thread's 128-row half of the tile = 4 of the slot's 8 words
// brings the 16 bytes kw_load(id, tile_half) will read into L2 (no register, no stall): issued one unit ahead for the 12
// term words that are not register-prefetched, so the real load pays L2 latency instead of an HBM round trip
__device__ __forceinline__ void kw_prefetch_l2(const BatchArgs& a, uint32_t id, int64_t tile_half) {
    if (id != NO_TERM) {
        const uint4* p = reinterpret_cast<const uint4*>(a.term_bits) + ((tile_half >> 1) * a.slot_cap + id) * 2 + (tile_half & 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
}
__device__ __forceinline__ uint4 kw_load(const BatchArgs& a, uint32_t id, int64_t tile_half) {
    return id != NO_TERM ? __ldg(reinterpret_cast<const uint4*>(a.term_bits) + ((tile_half >> 1) * a.slot_cap + id) * 2 + (tile_half & 1))
                         : make_uint4(0u, 0u, 0u, 0u);
}

// Epilogue of one thread (= one query) over its HC chunks of 32 accumulator columns.  The operand planes
// hold rows scaled to w_cos / |e|, so the fused screen score is  acc * (1/|q|) + w_rec*rec  (+ keyword):
// one FFMA per score, the row term read as broadcast float4 from smem.  The common path ends in a max
// tree; the rare work (a candidate above the query's threshold) sits behind ONE branch per chunk.
template <int MODE, bool HAS_KW>
__device__ __forceinline__ void epilogue_chunks(const BatchArgs& a, uint32_t taddr, const float* rec, int b, int64_t row0,
                                                int64_t dense_col0, int c_begin, float qs, float thr, float kww,
                                                const KwPlanes& kp) {
    uint32_t acc[2][32];
    tmem_ld32_issue(taddr + (uint32_t)(c_begin * 32), acc[0]);
#pragma unroll
    for (int cc = 0; cc < HC; ++cc) {
        const int c = c_begin + cc;
        uint32_t (&r)[32] = acc[cc & 1];
        tmem_ld32_wait(r);
        if (cc + 1 < HC) tmem_ld32_issue(taddr + (uint32_t)((c + 1) * 32), acc[(cc + 1) & 1]);
        float s[32];
        const float4* rc4 = reinterpret_cast<const float4*>(rec + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 rr = rc4[j];                                       // smem broadcast
            s[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), qs, rr.x);
            s[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), qs, rr.y);
            s[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), qs, rr.z);
            s[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), qs, rr.w);
        }
        // adds kww * (terms matched) to every row of the chunk
        auto add_keywords = [&]() {
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const uint32_t m = kp.p[cc][p];
                if (__any_sync(0xffffffffu, m != 0u)) {                     // warp-uniform; higher planes are almost always empty
                    const float add = kww * (float)(1 << p);
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (m & (1u << j)) s[j] += add;
                }
            }
        };
        if (HAS_KW && MODE != 0) add_keywords();
        if (MODE == 0) {
            float mx = s[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) mx = fmaxf(mx, s[j]);
            // Main pass: only rows above the query's threshold matter.  The keyword side is bit-sliced (5 count planes per
            // 32-row chunk); it is never added to all 32 scores here:
            //   1. the chunk's LARGEST match count comes out of the planes in 5 steps (walk down from the top plane, keeping
            //      the rows that still have every higher bit set): kw_room = kww * max count is the most any row can gain;
            //   2. only if  best keyword-free score + kw_room > thr  (warp vote) are the rows with  s > thr - kw_room
            //      collected (32 compares), and only THOSE rows get their own count extracted from the planes.
            // 16-term queries with frequent terms have non-empty planes almost everywhere; adding the planes to every row of
            // every chunk that passed a looser gate (the sum of the non-empty planes' weights) cost 0.6 ms of a 2.3 ms pass.
            int maxcnt = 0;
            if (HAS_KW) {
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int p = 4; p >= 0; --p) {
                    const uint32_t t = m & kp.p[cc][p];
                    if (t) { m = t; maxcnt |= 1 << p; }
                }
            }
            const float kw_room = maxcnt ? fmaf(fabsf(kww), (float)maxcnt, 1.0e-6f) : 0.f;
            if (__any_sync(0xffffffffu, mx + kw_room > thr)) {              // rare: ~1000 candidates per query and pass
                const float lim = thr - kw_room;
                uint32_t pass = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) pass |= (s[j] > lim) ? (1u << j) : 0u;
                while (pass) {
                    const int j = __ffs(pass) - 1;
                    pass &= pass - 1u;
                    float v[16];                                            // s[j] by a select tree (registers cannot be indexed)
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = (j & 16) ? s[16 + i] : s[i];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (j & 8) ? v[8 + i] : v[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = (j & 4) ? v[4 + i] : v[i];
                    v[0] = (j & 2) ? v[2] : v[0]; v[1] = (j & 2) ? v[3] : v[1];
                    float sv = (j & 1) ? v[1] : v[0];
                    if (HAS_KW) {                                           // this row's own match count
                        const int cnt = (int)((kp.p[cc][0] >> j) & 1u) | ((int)((kp.p[cc][1] >> j) & 1u) << 1) | ((int)((kp.p[cc][2] >> j) & 1u) << 2) |
                                        ((int)((kp.p[cc][3] >> j) & 1u) << 3) | ((int)((kp.p[cc][4] >> j) & 1u) << 4);
                        sv = fmaf(kww, (float)cnt, sv);
                    }
                    if (sv > thr) {
                        const uint32_t slot = atomicAdd(a.cand_count + b, 1u);
                        if (slot < (uint32_t)a.cand_cap)
                            a.cand[(int64_t)b * a.cand_cap + slot] = make_uint2((uint32_t)(row0 + c * 32 + j), __float_as_uint(sv));
                    }
                }
            }
        } else {
            if (a.dense_half) {                                             // sampling pass: 64 B per thread and chunk
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.dense) + (int64_t)b * a.dense_ld + (dense_col0 + c * 32));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __half2 h0 = __floats2half2_rn(s[8 * j + 0], s[8 * j + 1]), h1 = __floats2half2_rn(s[8 * j + 2], s[8 * j + 3]);
                    const __half2 h2 = __floats2half2_rn(s[8 * j + 4], s[8 * j + 5]), h3 = __floats2half2_rn(s[8 * j + 6], s[8 * j + 7]);
                    dst[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                        *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                }
            } else {
                float4* dst = reinterpret_cast<float4*>(a.dense + (int64_t)b * a.dense_ld + (dense_col0 + c * 32));
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
            }
        }
    }
}

// Unit = 256 queries x 256 corpus rows on a CTA PAIR (cta_group::2, UMMA 256x256x16): each CTA stages
// its own 128 queries (A half) and 128 rows (B half), so a k-block costs every SM 64 KB of L2->smem
// traffic for 2x the MMA work of a 128x128 single-CTA unit.  PASSES = 3: split precision
// (hi.hi + hi.mid + mid.hi); PASSES = 1: bf16 screen only (wider selection margin, see orr_api.cu).
// MODE 0 = main pass (threshold + candidate append), 1 = dense score store (sampling pass / debug).
template <int PASSES> struct GemmCfg {
    static constexpr int PLANES = PASSES == 3 ? 4 : 2;
    static constexpr int STAGE = PLANES * PLANE_BYTES;          // 64 KB / 32 KB per CTA
    static constexpr int NSTAGE = PASSES == 3 ? 3 : 6;
    static constexpr int REC_OFF = NSTAGE * STAGE;              // float rec[2][UN]
    static constexpr int QC_OFF = REC_OFF + 2 * UN * 4;         // float3-ish: qs, thr, kww  [MAX_QBLOCKS][BM] each
    static constexpr int ID_OFF = QC_OFF + 3 * MAX_QBLOCKS * BM * 4;          // uint16 ids [MAX_QBLOCKS][BM][16]
    static constexpr int BAR_OFF = ID_OFF + MAX_QBLOCKS * BM * ORR_BATCH_MAX_TERMS * 2;
    static constexpr int SMEM = BAR_OFF + 256 + 1024;           // barriers + alignment slack
};
static_assert(GemmCfg<3>::SMEM <= 232448 && GemmCfg<1>::SMEM <= 232448, "batched GEMM kernel exceeds 227 KB of shared memory");
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// CL = CTAs per cluster: 2 = one pair; 4 = two pairs that work on two different row tiles in lockstep while the QUERY
// k-blocks — the operand that is the same for every row tile — are TMA-MULTICAST to both pairs: each of the 4 CTAs loads
// one 64-query quarter of the k-block and multicasts it to the two CTAs that need it, so a query k-block crosses the
// L2 -> SM fabric once per 4 SMs instead of once per 2.  Both batched configs sit at the L2 fabric's throughput cap with
// pairs alone (ncu: 9.1 TB/s L2->SM at the power-capped clock; half of it query tiles re-read for every row tile).
template <int PASSES, int MODE, int CL>
__global__ void __launch_bounds__(BATCH_THREADS, 1)
orr_batch_gemm_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qmid,
                      const __grid_constant__ CUtensorMap map_ehi, const __grid_constant__ CUtensorMap map_emid,
                      const BatchArgs a) {
    using C = GemmCfg<PASSES>;
    constexpr int PP = CL / 2;                                              // pairs per cluster
    extern __shared__ uint8_t smem_raw[];
    // the dynamic smem base is only 16-B aligned by contract; every CTA computes the same offset
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_mem = smem;                                              // NSTAGE x STAGE, 1024-B aligned
    float* rec = reinterpret_cast<float*>(smem + C::REC_OFF);               // [2][UN]
    float* q_qs = reinterpret_cast<float*>(smem + C::QC_OFF);               // [MAX_QBLOCKS][BM]
    float* q_thr = q_qs + MAX_QBLOCKS * BM;
    float* q_kww = q_thr + MAX_QBLOCKS * BM;
    uint16_t* q_ids = reinterpret_cast<uint16_t*>(smem + C::ID_OFF);        // [MAX_QBLOCKS][BM][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* full_bar = bars;                          // [NSTAGE]  the pair leader's copy is the live one
    uint64_t* empty_bar = bars + C::NSTAGE;             // [NSTAGE]  per CTA: one arrival per pair of the cluster (multicast commit)
    uint64_t* tfull_bar = bars + 2 * C::NSTAGE;         // [2]       per CTA (multicast commit of its own pair)
    uint64_t* tempty_bar = bars + 2 * C::NSTAGE + 2;    // [2]       pair leader's copy: 16 epilogue warps of the pair
    uint64_t* aux_full = bars + 2 * C::NSTAGE + 4;      // [2]       per CTA
    uint64_t* aux_empty = bars + 2 * C::NSTAGE + 6;     // [2]       per CTA: 8 local epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t half = rank & 1u;                    // which 128 queries / 128 rows of the pair's unit this CTA stages
    const uint32_t pair = rank >> 1;                    // pair within the cluster
    const uint32_t pair_leader = rank & ~1u;
    const bool leader = half == 0;
    const uint16_t pair_mask = (uint16_t)(3u << pair_leader);
    const uint16_t all_mask = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), PP); }
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
    // this CTA's queries (128 per query block): per-query constants and term slots, staged once
    for (int i = threadIdx.x; i < a.n_qblocks * BM; i += BATCH_THREADS) {
        const int b = (i / BM) * 256 + (int)half * BM + (i % BM);
        q_qs[i] = a.qscale[b];
        q_thr[i] = MODE == 0 ? a.thr[b] : 0.f;
        q_kww[i] = a.q_kw_w ? a.q_kw_w[b] : 0.f;
        if (a.q_term_ids) {
            const int32_t* src = a.q_term_ids + (int64_t)b * ORR_BATCH_MAX_TERMS;
#pragma unroll
            for (int t = 0; t < ORR_BATCH_MAX_TERMS; ++t) q_ids[i * ORR_BATCH_MAX_TERMS + t] = src[t] < 0 ? NO_TERM : (uint16_t)src[t];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // barriers of EVERY CTA initialised, TMEM allocated, queries staged
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    // Work split: cluster g takes tile groups g, g + n_clusters, ...; group j = row tiles j*PP .. j*PP + PP - 1, one per pair.
    // Every CTA of a cluster runs the SAME number of units (the stage ring is shared by the multicast); a pair whose tile
    // does not exist (odd tile count) runs the unit on the group's first tile and discards it.
    const int n_clusters = (int)gridDim.x / CL, cid = (int)blockIdx.x / CL;
    const int n_groups = (a.n_row_tiles + PP - 1) / PP;
    const int units_per_tile = a.n_qblocks;
    const int my_groups = (n_groups > cid) ? (n_groups - 1 - cid) / n_clusters + 1 : 0;
    const int my_units = my_groups * units_per_tile;
    auto tile_of = [&](int u, bool* valid) {
        const int first = (cid + (u / units_per_tile) * n_clusters) * PP;
        const int t = first + (int)pair;
        *valid = t < a.n_row_tiles;
        return *valid ? t : first;
    };

    if (warp == 0) {
        // ===================== TMA producer (one per CTA) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = 0; u < my_units; ++u) {
                bool valid;
                const int t = tile_of(u, &valid);
                const int row_tile = t * a.row_tile_stride;
                const int qb = u % units_per_tile;
                const int buf = u & 1;
                mbar_wait(smem_u32(&aux_empty[buf]), ((uint32_t)(u >> 1) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&aux_full[buf]), UN * 4);
                bulk_g2s(smem_u32(rec + buf * UN), a.rowrec + (int64_t)row_tile * UN, UN * 4, smem_u32(&aux_full[buf]));
                const int qrow = qb * 256 + (int)half * BM;
                // this CTA's 128 rows of the unit are ONE 128-row tile of the k-block-major row planes
                const int64_t etile = ((int64_t)row_tile * (UN / BM) + (int64_t)half) * a.k_blocks;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);     // free in every CTA this one writes to
                    const uint32_t fb_local = smem_u32(&full_bar[stage]);
                    const uint32_t fb = mapa_shared(fb_local, pair_leader);
                    const uint32_t base = smem_u32(stage_mem + stage * C::STAGE);
                    if (leader) mbar_expect_tx(fb_local, 2 * C::STAGE);     // both halves of the pair land on this barrier
                    constexpr int E0 = PASSES == 3 ? 2 : 1;                 // first row plane of a stage
                    if (CL == 2) {
                        tma_load_2d_pair(base + 0 * PLANE_BYTES, &map_qhi, kb * BK, qrow, fb);
                        if (PASSES == 3) tma_load_2d_pair(base + 1 * PLANE_BYTES, &map_qmid, kb * BK, qrow, fb);
                    } else {
                        // this CTA's quarter (64 queries) of its half, to the same-half CTA of both pairs
                        const uint16_t qmask = (uint16_t)(0x5u << half);
                        const uint32_t qoff = pair * (uint32_t)(PLANE_BYTES / 2);
                        tma_load_2d_pair_mc(base + 0 * PLANE_BYTES + qoff, &map_qhi, kb * BK, qrow + (int)pair * (BM / 2), fb_local, qmask);
                        if (PASSES == 3)
                            tma_load_2d_pair_mc(base + 1 * PLANE_BYTES + qoff, &map_qmid, kb * BK, qrow + (int)pair * (BM / 2), fb_local, qmask);
                    }
                    // tiled planes: row (etile + kb) * 128 of the [tiles * kblocks * 128][64] view; row-major: (column, row)
                    const int ec0 = a.planes_tiled ? 0 : kb * BK;
                    const int ec1 = a.planes_tiled ? (int)((etile + kb) * BM) : row_tile * UN + (int)half * BM;
                    tma_load_2d_pair(base + E0 * PLANE_BYTES, &map_ehi, ec0, ec1, fb);
                    if (PASSES == 3) tma_load_2d_pair(base + 3 * PLANE_BYTES, &map_emid, ec0, ec1, fb);
                    if (++stage == C::NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (the leader CTA of each pair) =====================
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
                        umma_commit_mask(smem_u32(&empty_bar[stage]), all_mask);       // this pair is done with the stage, in every CTA
                        if (kb == a.k_blocks - 1) umma_commit_mask(smem_u32(&tfull_bar[buf]), pair_mask);
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
        const int c_begin = ((warp - 2) >> 2) * HC;
        const int qloc = quarter * 32 + lane;                               // query within the CTA's 128
        const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), pair_leader);
        const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty_bar[1]), pair_leader);
        const bool has_kw = a.q_term_ids != nullptr;
        // the first 4 term words of the NEXT unit are loaded while this unit's scores are processed
        uint4 wn[4];
        auto prefetch_words = [&](int u) {
            bool valid;
            const int row_tile = tile_of(u, &valid) * a.row_tile_stride;
            const int qi = (u % units_per_tile) * BM + qloc;
            const int64_t word0 = (int64_t)row_tile * 2 + (c_begin / HC);
            const uint2 id4 = *reinterpret_cast<const uint2*>(q_ids + qi * ORR_BATCH_MAX_TERMS);
            wn[0] = kw_load(a, id4.x & 0xFFFFu, word0); wn[1] = kw_load(a, id4.x >> 16, word0);
            wn[2] = kw_load(a, id4.y & 0xFFFFu, word0); wn[3] = kw_load(a, id4.y >> 16, word0);
        };
        if (has_kw && my_units > 0) prefetch_words(0);
        for (int u = 0; u < my_units; ++u) {
            bool valid;
            const int t = tile_of(u, &valid);
            const int row_tile = t * a.row_tile_stride;
            const int qb = u % units_per_tile;
            const int buf = u & 1;
            const int qi = qb * BM + qloc;
            const int b = qb * 256 + (int)half * BM + qloc;                 // this thread's query
            const float qs = q_qs[qi], thr = q_thr[qi], kww = q_kww[qi];
            const int64_t row0 = (int64_t)row_tile * UN;
            KwPlanes kp;
            if (has_kw) {
                const uint2* idp = reinterpret_cast<const uint2*>(q_ids + qi * ORR_BATCH_MAX_TERMS);
                const uint2 id1 = idp[1];
                if (!__any_sync(0xffffffffu, (id1.x & 0xFFFFu) != NO_TERM)) {
                    // every query of the warp has <= 4 terms: their words were prefetched during the last unit
                    kw_clear(kp);
                    if (kww != 0.f) { kw_add<3>(kp, wn[0]); kw_add<3>(kp, wn[1]); kw_add<3>(kp, wn[2]); kw_add<3>(kp, wn[3]); }
                } else {
                    // up to 16 terms: the other 12 words in flight together, then one adder tree per chunk
                    const int64_t word0 = (int64_t)row_tile * 2 + (c_begin / HC);
                    const uint2 id2 = idp[2], id3 = idp[3];
                    uint4 w[12];
                    w[0] = kw_load(a, id1.x & 0xFFFFu, word0); w[1] = kw_load(a, id1.x >> 16, word0);
                    w[2] = kw_load(a, id1.y & 0xFFFFu, word0); w[3] = kw_load(a, id1.y >> 16, word0);
                    w[4] = kw_load(a, id2.x & 0xFFFFu, word0); w[5] = kw_load(a, id2.x >> 16, word0);
                    w[6] = kw_load(a, id2.y & 0xFFFFu, word0); w[7] = kw_load(a, id2.y >> 16, word0);
                    w[8] = kw_load(a, id3.x & 0xFFFFu, word0); w[9] = kw_load(a, id3.x >> 16, word0);
                    w[10] = kw_load(a, id3.y & 0xFFFFu, word0); w[11] = kw_load(a, id3.y >> 16, word0);
                    kw_tree(kp, wn, w);
                    if (u + 1 < my_units) {                                 // the next unit's 12 further words -> L2, now
                        bool v2;
                        const int64_t nword0 = (int64_t)(tile_of(u + 1, &v2) * a.row_tile_stride) * 2 + (c_begin / HC);
                        const uint2* nidp = reinterpret_cast<const uint2*>(q_ids + (((u + 1) % units_per_tile) * BM + qloc) * ORR_BATCH_MAX_TERMS);
#pragma unroll
                        for (int t = 1; t < 4; ++t) {
                            const uint2 idn = nidp[t];
                            kw_prefetch_l2(a, idn.x & 0xFFFFu, nword0); kw_prefetch_l2(a, idn.x >> 16, nword0);
                            kw_prefetch_l2(a, idn.y & 0xFFFFu, nword0); kw_prefetch_l2(a, idn.y >> 16, nword0);
                        }
                    }
                }
                if (u + 1 < my_units) prefetch_words(u + 1);
            }
            mbar_wait(smem_u32(&aux_full[buf]), (uint32_t)(u >> 1) & 1u);
            mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)(u >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * UN);
            if (valid) {                                                    // warp-uniform; a filler unit's scores are discarded
                if (has_kw) epilogue_chunks<MODE, true>(a, taddr, rec + buf * UN, b, row0, (int64_t)t * UN, c_begin, qs, thr, kww, kp);
                else epilogue_chunks<MODE, false>(a, taddr, rec + buf * UN, b, row0, (int64_t)t * UN, c_begin, qs, thr, 0.f, kp);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&aux_empty[buf]));
                mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
            }
        }
    }
    // no CTA may leave while a peer can still signal its barriers or write its smem
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- split planes of the store: x^ = w_cos * x / |x| = hi + mid (+ dropped lo), bf16 each ----------------
// Rows are stored pre-scaled so the accumulator already is w_cos * cos * |q| and the epilogue needs one FFMA
// per score.  Rows whose squared norm is 0, non-finite or overflows fp32 get all-zero planes (screen cosine 0,
// as CosineSimilarity returns for a zero norm, RecallSearchService.cs:84-85).
// LAYOUT of a row plane: k-block-major tiles.  The GEMM's TMA box is 128 rows x 64 columns (one 128-byte swizzle row per
// corpus row); in a row-major plane that box is 128 separate 128-byte pieces 2*dim bytes apart, and the 12 k-blocks of a
// unit revisit every DRAM page 12 times, ~0.6 us apart: the main pass of the single-query-block config reached only
// 4.5 TB/s of HBM.  Here element (row r, column c) lives at
//     ((r / 128) * (dim / 64) + c / 64) * 128 * 64  +  (r % 128) * 64  +  c % 64
// so every box is ONE contiguous 16 KB piece and a CTA's 12 k-blocks of a unit are 192 KB of sequential HBM.  The TMA
// tensor map sees a [rows_padded * dim / 64][64] matrix.  Rows of the last tile beyond `first + n` are zero-filled.
__global__ void __launch_bounds__(256) orr_build_planes_kernel(const float* emb, __nv_bfloat16* hi, __nv_bfloat16* mid,
                                                               int64_t first, int64_t n, int64_t n_fill, int dim, float w_cos, int tiled) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int kblocks = dim / BK;
    for (int64_t i = gw; i < n_fill; i += W) {
        const int64_t row = first + i;
        const bool real = i < n;                                            // else: padding row of the last tile
        const float4* x4 = reinterpret_cast<const float4*>(emb + row * dim);
        float ss = 0.f;
        if (real)
            for (int c = lane; c < dim / 4; c += 32) {
                const float4 v = x4[c];
                ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float scale = (real && ss > 0.f && ss < 3e38f) ? w_cos * rsqrtf(ss) : 0.f;
        const int64_t tile_base = tiled ? (row / BM) * (int64_t)kblocks * BM * BK + (row % BM) * BK : row * (int64_t)dim;
        for (int c = lane; c < dim / 4; c += 32) {                          // second read of the row hits L1/L2
            const float4 v = scale != 0.f ? x4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float e[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
            __nv_bfloat16 h[4], m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float ek = scale != 0.f ? e[k] : 0.f;                 // NaN/Inf elements of a rejected row -> 0
                h[k] = __float2bfloat16_rn(ek);
                m[k] = __float2bfloat16_rn(ek - __bfloat162float(h[k]));
            }
            const int col = 4 * c;
            const int64_t at = tiled ? tile_base + (int64_t)(col / BK) * BM * BK + (col % BK) : tile_base + col;
            if (hi) {
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(hi + at);
                h2[0] = __halves2bfloat162(h[0], h[1]); h2[1] = __halves2bfloat162(h[2], h[3]);
            }
            if (mid) {
                __nv_bfloat162* m2 = reinterpret_cast<__nv_bfloat162*>(mid + at);
                m2[0] = __halves2bfloat162(m[0], m[1]); m2[1] = __halves2bfloat162(m[2], m[3]);
            }
        }
    }
}

// qbad[b] = 1: the query's squared norm is not representable in fp32 (under/overflow, NaN) although the query is not
// all-zero — the screen cannot rank it (the reference accumulates the norm in fp64, RecallSearchService.cs:77-82); the
// finalize kernel flags such queries and they re-run singly (-> exact path).
__global__ void __launch_bounds__(128) orr_prep_queries_kernel(const float* q, __nv_bfloat16* qhi, __nv_bfloat16* qmid,
                                                               float* qscale, int32_t* qbad, int batch, int dim) {
    const int b = blockIdx.x;
    __shared__ float red[4], redm[4];
    float ss = 0.f, mx = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        const float v = b < batch ? q[(int64_t)b * dim + c] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        qhi[(int64_t)b * dim + c] = h;
        qmid[(int64_t)b * dim + c] = __float2bfloat16_rn(v - __bfloat162float(h));
        ss = fmaf(v, v, ss);
        mx = fmaxf(mx, fabsf(v));
        if (v != v) mx = INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = ss; redm[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float t = red[0] + red[1] + red[2] + red[3];
        const float m = fmaxf(fmaxf(redm[0], redm[1]), fmaxf(redm[2], redm[3]));
        const bool ok = t >= 1e-30f && t <= 1e30f;
        qscale[b] = ok ? rsqrtf(t) : 0.f;
        if (qbad) qbad[b] = (!ok && m != 0.f) ? 1 : 0;
    }
}

// per-batch row side of the fused score: w_rec * exp(-age/30d); tombstones and padding -> -inf
__global__ void orr_build_rowrec_kernel(const int64_t* ticks, float* rowrec, int64_t rows, int64_t rows_padded,
                                        int64_t now_ticks, float w_rec, float decay_per_2p20) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_padded) return;
    float v = -INFINITY;
    if (i < rows) {
        const int64_t tk = ticks[i];
        if (tk != ORR_DEAD_TICKS) {
            int64_t age20 = (now_ticks - tk) >> 20;
            age20 = age20 < 0 ? 0 : (age20 > 0x7fffffffLL ? 0x7fffffffLL : age20);
            v = w_rec * __expf(-(float)(int32_t)age20 * decay_per_2p20);
        }
    }
    rowrec[i] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row plane in k-block-major tiles (see orr_build_planes_kernel): a [ceil(rows/128) * dim/64 * 128][64] matrix of bf16
int make_row_plane_map(CUtensorMap* map, const void* base, int64_t rows, int dim);

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

int make_row_plane_map(CUtensorMap* map, const void* base, int64_t rows, int dim) {
    const int64_t tiles = (rows + BM - 1) / BM;
    return make_plane_map(map, base, tiles * (dim / BK) * BM, BK, BM);
}

}  // namespace

// ---- host-side launchers ---------------------------------------------------------------------------
// k-block-major tiles (default) or plain row-major planes (ORR_PLANES_TILED=0: the round-1 layout, kept for A/B runs)
bool orr_batch_planes_tiled() {
    static const int tiled = [] { const char* e = getenv("ORR_PLANES_TILED"); return (e && *e == '0') ? 0 : 1; }();
    return tiled != 0;
}

int64_t orr_batch_plane_elems(int64_t capacity_rows, int dim) {              // capacity rounded up to whole 128-row tiles
    return (capacity_rows + BM - 1) / BM * BM * (int64_t)dim;
}

int orr_batch_build_planes(const float* emb, void* hi, void* mid, int64_t first, int64_t n, int dim, float w_cos,
                           cudaStream_t st) {
    if (n <= 0) return ORR_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_fill = (first + n + BM - 1) / BM * BM - first;          // up to the end of the last tile (zero rows)
    orr_build_planes_kernel<<<sms * 8, 256, 0, st>>>(emb, (__nv_bfloat16*)hi, (__nv_bfloat16*)mid, first, n, n_fill, dim, w_cos,
                                                     orr_batch_planes_tiled() ? 1 : 0);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_prep_queries(const float* q_dev, void* qhi, void* qmid, float* qscale, int32_t* qbad, int batch, int batch_padded,
                           int dim, cudaStream_t st) {
    orr_prep_queries_kernel<<<batch_padded, 128, 0, st>>>(q_dev, (__nv_bfloat16*)qhi, (__nv_bfloat16*)qmid, qscale, qbad, batch, dim);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_build_rowrec(const int64_t* ticks, float* rowrec, int64_t rows, int64_t rows_padded, int64_t now_ticks,
                           const OrrWeights& w, cudaStream_t st) {
    const float decay = (float)(1048576.0 / ((double)ORR_TICKS_PER_DAY * w.recency_days));
    orr_build_rowrec_kernel<<<(unsigned)((rows_padded + 255) / 256), 256, 0, st>>>(ticks, rowrec, rows, rows_padded, now_ticks,
                                                                                  (float)w.w_rec, decay);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

template <int PASSES, int MODE, int CL>
static int launch_gemm_cl(const CUtensorMap& mqh, const CUtensorMap& mqm, const CUtensorMap& meh, const CUtensorMap& mem,
                          const BatchArgs& a, int grid, cudaStream_t st) {
    ORR_SMEM_OPT_IN((orr_batch_gemm_kernel<PASSES, MODE, CL>), GemmCfg<PASSES>::SMEM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(BATCH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = GemmCfg<PASSES>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ORR_CUDA_OK(cudaLaunchKernelEx(&cfg, orr_batch_gemm_kernel<PASSES, MODE, CL>, mqh, mqm, meh, mem, a));
    return ORR_OK;
}

// how many clusters of CL CTAs of this kernel the device can hold at once (GPC boundaries can strand SMs)
template <int PASSES, int MODE, int CL>
static int max_clusters(int sms) {
    cudaFuncSetAttribute(orr_batch_gemm_kernel<PASSES, MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<PASSES>::SMEM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sms / CL * CL), 1, 1);
    cfg.blockDim = dim3(BATCH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = GemmCfg<PASSES>::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, orr_batch_gemm_kernel<PASSES, MODE, CL>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Cluster shape of the batched GEMM: 4 (query k-blocks multicast to two pairs) when the device can keep (nearly) as many
// SMs busy with clusters of 4 as with pairs, else 2.  ORR_BATCH_CLUSTER=2|4 overrides (experiments).
struct ClusterPlan { int cl; int n_clusters; };
template <int PASSES, int MODE>
static ClusterPlan plan_clusters(int sms) {
    static int choice[64] = {};                                              // per device ordinal: 0 = undecided
    static int clusters[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!choice[dev]) {
        const int c2 = max_clusters<PASSES, MODE, 2>(sms), c4 = max_clusters<PASSES, MODE, 4>(sms);
        clusters[dev][0] = c2 > 0 ? std::min(c2, sms / 2) : sms / 2;
        clusters[dev][1] = std::min(c4, sms / 4);
        int pick = (4 * clusters[dev][1] >= 2 * clusters[dev][0] - 4) ? 4 : 2;   // tolerate up to 4 stranded SMs for the halved query traffic
        if (const char* e = getenv("ORR_BATCH_CLUSTER")) { const int v = atoi(e); if (v == 2 || (v == 4 && clusters[dev][1] > 0)) pick = v; }
        if (getenv("ORR_BATCH_TRACE")) fprintf(stderr, "[orr batch] device %d: %d pairs / %d clusters of 4 fit; cluster size %d\n", dev, c2, c4, pick);
        choice[dev] = pick;
    }
    return choice[dev] == 4 ? ClusterPlan{4, clusters[dev][1]} : ClusterPlan{2, clusters[dev][0]};
}
template <int PASSES, int MODE>
static int launch_gemm_t(const CUtensorMap& mqh, const CUtensorMap& mqm, const CUtensorMap& meh, const CUtensorMap& mem,
                         const BatchArgs& a, const ClusterPlan& plan, cudaStream_t st) {
    if (plan.cl == 4) {
        const int n = std::min(plan.n_clusters, (a.n_row_tiles + 1) / 2);
        return launch_gemm_cl<PASSES, MODE, 4>(mqh, mqm, meh, mem, a, 4 * std::max(1, n), st);
    }
    const int pairs = std::min(plan.n_clusters, a.n_row_tiles);
    return launch_gemm_cl<PASSES, MODE, 2>(mqh, mqm, meh, mem, a, 2 * std::max(1, pairs), st);
}

int orr_batch_launch_gemm(const OrrBatchGemm& g, cudaStream_t st) {
    if (g.dim % BK != 0) { orr_set_error("batch path needs dim %% 64 == 0 (dim=%d)", g.dim); return ORR_E_UNSUPPORTED; }
    if (g.batch_padded % 256 != 0 || g.batch_padded > ORR_BATCH_MAX_QUERIES) {
        orr_set_error("batch path: padded batch %d not a multiple of 256 in [256, %d]", g.batch_padded, ORR_BATCH_MAX_QUERIES);
        return ORR_E_INTERNAL;
    }
    const bool dense = g.dense != nullptr;
    const ClusterPlan plan = g.passes == 1 ? (dense ? plan_clusters<1, 1>(g.sms) : plan_clusters<1, 0>(g.sms))
                                           : (dense ? plan_clusters<3, 1>(g.sms) : plan_clusters<3, 0>(g.sms));
    const int q_box = plan.cl == 4 ? BM / 2 : BM;           // clusters of 4: every CTA loads (and multicasts) a 64-query quarter
    CUtensorMap mqh, mqm, meh, mem;
    int rc;
    if ((rc = make_plane_map(&mqh, g.qhi, g.batch_padded, g.dim, q_box)) != ORR_OK) return rc;
    if ((rc = make_plane_map(&mqm, g.qmid, g.batch_padded, g.dim, q_box)) != ORR_OK) return rc;
    const bool tiled = orr_batch_planes_tiled();
    const void* mid_plane = g.passes == 1 ? g.ehi : g.emid;  // the bf16 screen never touches the mid planes: their maps alias the hi planes
    if ((rc = tiled ? make_row_plane_map(&meh, g.ehi, g.rows, g.dim) : make_plane_map(&meh, g.ehi, g.rows, g.dim, BM)) != ORR_OK) return rc;
    if ((rc = tiled ? make_row_plane_map(&mem, mid_plane, g.rows, g.dim) : make_plane_map(&mem, mid_plane, g.rows, g.dim, BM)) != ORR_OK) return rc;
    BatchArgs a{};
    const int64_t all_tiles = (g.rows + UN - 1) / UN;
    a.row_tile_stride = g.tile_stride < 1 ? 1 : g.tile_stride;
    a.n_row_tiles = (int32_t)((all_tiles + a.row_tile_stride - 1) / a.row_tile_stride);
    a.n_qblocks = g.batch_padded / 256;
    a.k_blocks = g.dim / BK;
    a.rows = g.rows;
    a.rowrec = (const float*)g.rowaux;
    a.qscale = g.qscale;
    a.thr = g.thr;
    a.cand = (uint2*)g.cand;
    a.cand_count = g.cand_count;
    a.cand_cap = g.cand_cap;
    a.dense = g.dense;
    a.dense_ld = g.dense_ld;
    a.dense_half = g.dense_half;
    a.term_bits = g.term_bits;
    a.slot_cap = g.slot_cap;
    a.q_term_ids = g.q_term_ids;
    a.q_kw_w = g.q_kw_w;
    a.planes_tiled = tiled ? 1 : 0;
    if (a.n_row_tiles < 1) return ORR_OK;
    if (dense)
        return g.passes == 1 ? launch_gemm_t<1, 1>(mqh, mqm, meh, mem, a, plan, st)
                             : launch_gemm_t<3, 1>(mqh, mqm, meh, mem, a, plan, st);
    return g.passes == 1 ? launch_gemm_t<1, 0>(mqh, mqm, meh, mem, a, plan, st)
                         : launch_gemm_t<3, 0>(mqh, mqm, meh, mem, a, plan, st);
}
