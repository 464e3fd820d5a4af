// orr_scan.cu — K1: the fused single-query scan (sm_100a).
//
// Replaces the per-chunk loop of RecallSearchService.SearchAsync
// (src/OmniRecall.Api/Services/RecallSearchService.cs:28-33: ScoreChunk :59-67 =
// CosineSimilarity :69-88 + KeywordScore :90-113 + RecencyScore :115-119) for SELECTION
// only: one pass over the HBM-resident row-major fp32 store computes an fp32 score per
// row and keeps the best candidates; K3 (orr_rescore.cu) then re-scores those candidates
// with the reference's exact fp64 arithmetic and proves the selection was safe.
//
// Shape of the kernel (HBM-bound: 4*D + 8 + 4*slots bytes per row, ~6 flop per 4 bytes):
//   - persistent grid, one CTA per SM, 8 warps; every warp is its own TMA producer and
//     consumer: lane 0 issues cp.async.bulk (UBLKCP) of one contiguous tile of rows
//     (<= 12 KB) into the warp's private 2-stage smem ring, completion on an mbarrier;
//     8 warps x 2 stages x 12 KB = 192 KB in flight per SM, no inter-warp sync in the loop;
//   - the query lives in registers (D/128 float4 per lane), rows are read back from smem
//     with conflict-free LDS.128, dot and ||b||^2 accumulate in 8 fp32 chains per lane and
//     finish with a warp-shuffle butterfly;
//   - per row the warp also probes the chunk's hashed term set (lane-parallel compares +
//     one REDUX.OR), evaluates exp(-age/30d) and fuses 0.7/0.2/0.1;
//   - top-k: one (score,row) entry per lane = a 32-entry per-warp list; a row enters only
//     if it beats the warp minimum (REDUX.MIN + ballot to find the new minimum);
//   - tail: per-CTA bitonic merge of the warp lists, then the LAST CTA to finish (ticket)
//     radix-selects the global survivors and the discard bound tau for K3's check.
#include <cfloat>
#include <cstdlib>

#include "orr_internal.h"

static int env_int(const char* name, int dflt) {      // tuning knobs for experiments
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// Rows are read once per query: L2 evict_first keeps the stream from thrashing L2.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_nohint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t ldg_nc_s64(const int64_t* p) {
    int64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// monotone float -> uint32 (larger float => larger key); key 0 is reserved for "empty"
__device__ __forceinline__ uint32_t order_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
// Returns 0 after a warp vote over a predicate computed from `dep`: the result is only
// available once every lane has produced its `dep`, i.e. once all the loads feeding it
// have returned.  Opaque to the compiler on purpose.
__device__ __forceinline__ uint32_t all_lanes_done(uint32_t dep) {
    uint32_t z;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0xffffffff;\n\t"
        "vote.sync.all.pred q, p, 0xffffffff;\n\t"
        "selp.u32 %0, 0, 0, q;\n\t}"
        : "=r"(z)
        : "r"(dep));
    return z;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct ScanArgs {
    OrrShard  sh;
    const float* q;          // device fp32[dim]
    OrrProbes pr;
    float     w_cos, w_kw, w_rec;
    float     decay_per_2p20;    // 2^20 / (ticks per day * recency_days)
    float     inv_nterms;        // 1 / |terms|
    int64_t   now_ticks;
    int32_t   warps;             // consumer warps per CTA
    int32_t   stages;            // TMA stages per warp
    int32_t   l2_hint;           // 1: evict_first on the row stream
    int32_t   stage_bytes;       // bytes of one smem stage (tile_rows * dim * 4)
    int32_t   n_surv;            // survivors to hand to K3 (64, 128 or 256)
    uint2*    cta_cands;         // [grid][n_surv]
    float*    cta_floor;         // [grid]
    int32_t*  sel;               // {n_surv_out, tau_bits, ticket, ...}
    uint32_t* surv_rows;
    // text mode: up to 32 query terms whose matches come from row bitmaps (orr_textmatch.cu) instead of probes
    const uint32_t* bm_bits;     // [n_bm][bm_row_words] or NULL
    int64_t   bm_row_words;
    int32_t   n_bm;
    float*    dense;             // diagnostic (orr_debug_scan_scores): every row's fp32 scan score, or NULL
};

// ---- the per-row fp32 epilogue ---------------------------------------------------------
__device__ __forceinline__ float fuse_row(const ScanArgs& a, float dot, float nb, float inv_qn,
                                          int64_t ticks, const uint32_t* th, int spl, int bm_matches) {
    if (ticks == ORR_DEAD_TICKS) return -INFINITY;                 // tombstone
    // cosine (RecallSearchService.cs:84-87); zero row => 0
    float cosv = 0.f;
    bool force = false;
    if (nb > 0.f) {
        cosv = dot * inv_qn * rsqrtf(nb);
        // magnitudes where fp32 products over/underflow: let K3 decide exactly
        force = !(nb >= 1e-30f && nb <= 1e30f) || !(fabsf(dot) <= 3e38f);
    } else if (!(nb == 0.f)) {
        force = true;                                              // NaN
    }
    // keyword (:110-112): which query terms does the chunk's hashed term set contain
    float kw = (float)bm_matches * a.inv_nterms;                    // terms matched through row bitmaps
    if (a.pr.n_probes > 0) {
        uint32_t m0 = 0, m1 = 0;
        for (int p = 0; p < a.pr.n_probes; ++p) {
            const uint32_t h = a.pr.h32[p];
            bool hit = (th[0] == h);
            if (spl > 1) hit |= (th[1] == h);
            if (spl > 2) hit |= (th[2] == h) | (th[3] == h);
            if (hit) {
                const uint32_t t = a.pr.term[p];
                if (t < 32) m0 |= 1u << t; else m1 |= 1u << (t - 32);
            }
        }
        m0 = __reduce_or_sync(FULL, m0);
        if (a.pr.n_terms > 32) m1 = __reduce_or_sync(FULL, m1);
        kw = (float)(__popc(m0) + __popc(m1) + bm_matches) * a.inv_nterms;
    }
    // recency (:115-119): age in units of 2^20 ticks (0.1 s) fits int32 for ~7 years, beyond
    // which exp(-age/30d) < 1e-37; selection-grade only, K3 recomputes it in fp64
    int64_t age20 = (a.now_ticks - ticks) >> 20;
    age20 = age20 < 0 ? 0 : (age20 > 0x7fffffffLL ? 0x7fffffffLL : age20);
    const float rec = __expf(-(float)(int32_t)age20 * a.decay_per_2p20);
    float s = a.w_cos * cosv + a.w_kw * kw + a.w_rec * rec;
    if (force) s = FLT_MAX;
    if (s != s) s = -FLT_MAX;                                      // NaN ranks last (:34)
    return fminf(fmaxf(s, -FLT_MAX), FLT_MAX);
}

// NV = float4 per lane per row (dim/128) with the query in registers; NV == 0: generic dim,
// query read from smem.  TR = rows per tile.
template <int NV, int TR, bool PIPE>
__global__ void __launch_bounds__(ORR_SCAN_WARPS * 32, 1) orr_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x;
    const int dim = a.sh.dim;
    const int nv4 = dim >> 2;                                    // float4 per row
    const uint32_t q_bytes = (uint32_t)((dim * 4 + 127) & ~127);
    float* sq = reinterpret_cast<float*>(smem);
    uint8_t* stage_base = smem + q_bytes;
    const uint32_t ring_bytes = (uint32_t)a.warps * (uint32_t)a.stages * (uint32_t)a.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + ring_bytes);

    // ---- prologue: query -> smem (+registers), barriers ----
    for (int i = tid; i < nv4; i += nthreads)
        reinterpret_cast<float4*>(sq)[i] = __ldg(reinterpret_cast<const float4*>(a.q) + i);
    if (tid < a.warps * a.stages) mbar_init(smem_u32(&bars[tid]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    float4 qr[NV > 0 ? NV : 1];
    float qn = 0.f, qmax = 0.f;
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            qr[j] = reinterpret_cast<const float4*>(sq)[j * 32 + lane];
            qn += qr[j].x * qr[j].x + qr[j].y * qr[j].y + qr[j].z * qr[j].z + qr[j].w * qr[j].w;
            qmax = fmaxf(qmax, fmaxf(fmaxf(fabsf(qr[j].x), fabsf(qr[j].y)), fmaxf(fabsf(qr[j].z), fabsf(qr[j].w))));
        }
    } else {
        for (int i = lane; i < nv4; i += 32) {
            float4 v = reinterpret_cast<const float4*>(sq)[i];
            qn += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            qmax = fmaxf(qmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
    qn = warp_sum(qn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qmax = fmaxf(qmax, __shfl_xor_sync(FULL, qmax, o));
    const float inv_qn = (qn > 0.f) ? rsqrtf(qn) : 0.f;            // :84 normA <= 0 -> 0
    // The selection proof assumes |fp32 score - exact| <= eps, which needs ||q||^2 representable in fp32: a query whose
    // squared norm under/overflows (|q_i| ~ 1e-20 or 1e+19; the reference accumulates it in fp64 and returns a real
    // cosine) or holds NaN makes every fp32 cosine meaningless.  An all-zero query is fine (cosine 0 everywhere, :84).
    // Such a query reports tau = +inf: K3's bound check fails and the caller escalates to the exact path.
    const bool q_unscreenable = !(qmax == 0.f) && !(qn >= 1e-30f && qn <= 1e30f);

    // ---- per-warp pipeline over interleaved tiles ----
    const int64_t rows = a.sh.rows;
    const int64_t n_tiles = (rows + TR - 1) / TR;
    const int64_t gw = (int64_t)blockIdx.x * a.warps + warp;
    const int64_t W = (int64_t)gridDim.x * a.warps;
    const uint32_t my_stage0 = smem_u32(stage_base) + (uint32_t)warp * (uint32_t)a.stages * (uint32_t)a.stage_bytes;
    const uint32_t my_bar0 = smem_u32(&bars[warp * a.stages]);
    const uint32_t row_bytes = (uint32_t)dim * 4u;
    const uint64_t policy = policy_evict_first();
    const int spl = a.sh.slots >> 5;                              // term words per lane per row

    auto issue = [&](int64_t tile, int stage, uint32_t zero) {
        const int64_t r0 = tile * TR;
        const int64_t nr = (rows - r0 < TR) ? (rows - r0) : TR;
        const uint32_t bytes = (uint32_t)nr * row_bytes + zero;
        const uint32_t bar = my_bar0 + (uint32_t)stage * 8u;
        mbar_expect_tx(bar, bytes);
        if (a.l2_hint)
            bulk_g2s(my_stage0 + (uint32_t)stage * (uint32_t)a.stage_bytes,
                     a.sh.emb + r0 * (int64_t)dim, bytes, bar, policy);
        else
            bulk_g2s_nohint(my_stage0 + (uint32_t)stage * (uint32_t)a.stage_bytes,
                            a.sh.emb + r0 * (int64_t)dim, bytes, bar);
    };

    if (lane == 0) {
        for (int s = 0; s < a.stages; ++s) {
            const int64_t t = gw + (int64_t)s * W;
            if (t < n_tiles) issue(t, s, 0u);
        }
    }

    // per-warp top list: one entry per lane
    float es = -INFINITY;
    uint32_t er = 0xffffffffu;
    float wmin = -INFINITY;
    int wmin_lane = 0;

    // Software pipeline.  A tile's scalar epilogue (shuffle butterflies, keyword probe,
    // recency, fuse, list insert) is a long dependent chain; it is issued one iteration late,
    // after the NEXT tile's LDS/FMA phase, so the two chains overlap inside one warp and the
    // per-row ticks/term loads have a whole iteration to land.
    float pd[TR], pn[TR];                      // pending tile: per-lane partial sums
    uint32_t pth[TR][4];                       // pending tile: this lane's term words
    int64_t ptk = 0, pr0 = 0;
    int pnr = 0;
    uint32_t pbw = 0u;                         // pending tile: lane t's bitmap word of bitmap term t
#pragma unroll
    for (int r = 0; r < TR; ++r) { pd[r] = pn[r] = 0.f; pth[r][0] = pth[r][1] = pth[r][2] = pth[r][3] = 0u; }

    auto finish_pending = [&]() {
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            if (r < pnr) {                                               // warp-uniform
                const float dot = warp_sum(pd[r]);
                const float nb = warp_sum(pn[r]);
                const int64_t ticks = __shfl_sync(FULL, ptk, r);
                const int bm = a.n_bm ? __popc(__ballot_sync(FULL, (pbw >> ((uint32_t)(pr0 + r) & 31u)) & 1u)) : 0;
                const float s = fuse_row(a, dot, nb, inv_qn, ticks, pth[r], spl, bm);
                if (a.dense != nullptr && lane == 0) a.dense[pr0 + r] = s;
                if (s > wmin) {                                          // warp-uniform
                    if (lane == wmin_lane) { es = s; er = (uint32_t)(pr0 + r); }
                    const uint32_t k = order_key(es);
                    const uint32_t mk = __reduce_min_sync(FULL, k);
                    wmin_lane = __ffs(__ballot_sync(FULL, k == mk)) - 1;
                    wmin = key_to_float(mk);
                }
            }
        }
    };

    int stage = 0;
    uint32_t parity = 0;
    for (int64_t tile = gw; tile < n_tiles; tile += W) {
        const int64_t r0 = tile * TR;
        const int nr = (int)((rows - r0 < TR) ? (rows - r0) : TR);

        // this tile's per-row scalars straight from global; consumed one iteration later
        int64_t tk = 0;
        if (lane < nr) tk = ldg_nc_s64(a.sh.ticks + r0 + lane);
        uint32_t bw = 0u;                      // TR divides 32: a tile's rows share one bitmap word
        if (lane < a.n_bm) bw = ldg_nc_u32(a.bm_bits + (int64_t)lane * a.bm_row_words + (r0 >> 5));
        uint32_t th[TR][4];
#pragma unroll
        for (int r = 0; r < TR; ++r) { th[r][0] = th[r][1] = th[r][2] = th[r][3] = 0u; }
        if (a.pr.n_probes > 0) {
#pragma unroll
            for (int r = 0; r < TR; ++r) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    if (r < nr && w < spl)
                        th[r][w] = ldg_nc_u32(a.sh.terms32 + (r0 + r) * (int64_t)a.sh.slots + w * 32 + lane);
                }
            }
        }

        mbar_wait(my_bar0 + (uint32_t)stage * 8u, parity);
        const uint8_t* sptr = stage_base + ((size_t)warp * a.stages + stage) * (size_t)a.stage_bytes;

        float cd[TR], cn[TR];
        uint32_t dep = 0;
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
            if (r < nr) {
                const float4* xrow = reinterpret_cast<const float4*>(sptr + (size_t)r * row_bytes);
                if (NV > 0) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const float4 x = xrow[j * 32 + lane];
                        d0 = fmaf(qr[j].x, x.x, d0); d1 = fmaf(qr[j].y, x.y, d1);
                        d2 = fmaf(qr[j].z, x.z, d2); d3 = fmaf(qr[j].w, x.w, d3);
                        n0 = fmaf(x.x, x.x, n0); n1 = fmaf(x.y, x.y, n1);
                        n2 = fmaf(x.z, x.z, n2); n3 = fmaf(x.w, x.w, n3);
                    }
                } else {
                    for (int i = lane; i < nv4; i += 32) {
                        const float4 x = xrow[i];
                        const float4 qq = reinterpret_cast<const float4*>(sq)[i];
                        d0 = fmaf(qq.x, x.x, d0); d1 = fmaf(qq.y, x.y, d1);
                        d2 = fmaf(qq.z, x.z, d2); d3 = fmaf(qq.w, x.w, d3);
                        n0 = fmaf(x.x, x.x, n0); n1 = fmaf(x.y, x.y, n1);
                        n2 = fmaf(x.z, x.z, n2); n3 = fmaf(x.w, x.w, n3);
                    }
                }
            }
            cd[r] = (d0 + d1) + (d2 + d3);
            cn[r] = (n0 + n1) + (n2 + n3);
            dep |= __float_as_uint(cd[r]) | __float_as_uint(cn[r]);
        }

        // Refill this stage.  The vote consumes a value derived from every lane's loads, so the
        // bulk copy cannot be issued before all of this tile's smem reads have returned.
        {
            const uint32_t z = all_lanes_done(dep);                      // always 0
            const int64_t nt = tile + (int64_t)a.stages * W;
            if (lane == 0 && nt < n_tiles) issue(nt, stage, z);
        }

        if (PIPE) finish_pending();                                      // previous tile
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            pd[r] = cd[r]; pn[r] = cn[r];
            pth[r][0] = th[r][0]; pth[r][1] = th[r][1]; pth[r][2] = th[r][2]; pth[r][3] = th[r][3];
        }
        ptk = tk; pr0 = r0; pnr = nr; pbw = bw;
        if (!PIPE) { finish_pending(); pnr = 0; }                        // this tile, immediately
        if (++stage == a.stages) { stage = 0; parity ^= 1u; }
    }
    finish_pending();                                                    // last tile

    // ---- tail 1: per-CTA merge of the warp lists -> best n_surv of this CTA ----
    __syncthreads();                                                 // all stages idle
    uint64_t* mkeys = reinterpret_cast<uint64_t*>(stage_base);       // 256 x u64, reuses the ring
    float* wfloor = reinterpret_cast<float*>(stage_base + 256 * 8);  // per-warp discard bound
    for (int i = tid; i < 256; i += nthreads) mkeys[i] = 0ull;
    __syncthreads();
    {
        // sortable: score key high, ~row low (equal scores: lower row first)
        const uint32_t k = (es == -INFINITY) ? 0u : order_key(es);
        mkeys[warp * 32 + lane] = k ? (((uint64_t)k << 32) | (uint64_t)(~er)) : 0ull;
        if (lane == 0) wfloor[warp] = wmin;                          // -inf unless the list filled
    }
    __syncthreads();
    const int M = a.n_surv;
    float cta_floor = -INFINITY;
    if (M < 256) {
        // bitonic sort, descending
        for (int k2 = 2; k2 <= 256; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < 256; i += nthreads) {
                    const int p = i ^ j;
                    if (p > i) {
                        const uint64_t x = mkeys[i], y = mkeys[p];
                        const bool desc = ((i & k2) == 0);
                        if (desc ? (x < y) : (x > y)) { mkeys[i] = y; mkeys[p] = x; }
                    }
                }
                __syncthreads();
            }
        }
        const uint64_t next = mkeys[M];
        if (next) cta_floor = key_to_float((uint32_t)(next >> 32));
    }
    for (int w = 0; w < a.warps; ++w) cta_floor = fmaxf(cta_floor, wfloor[w]);
    for (int i = tid; i < M; i += nthreads) {
        const uint64_t e = (i < 256) ? mkeys[i] : 0ull;
        a.cta_cands[(int64_t)blockIdx.x * M + i] = make_uint2((uint32_t)(e >> 32), ~(uint32_t)e);
    }
    if (tid == 0) a.cta_floor[blockIdx.x] = cta_floor;

    // ---- tail 2: the last CTA to arrive selects the global survivors ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.sel[2], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining, s_cnt_gt, s_cnt_eq, s_total;
    __shared__ float s_wtau[ORR_SCAN_WARPS];
    const int E = (int)gridDim.x * M;
    const volatile uint2* cands = a.cta_cands;
    if (tid == 0) { s_prefix = 0; s_remaining = (uint32_t)M; s_cnt_gt = 0; s_cnt_eq = 0; s_total = 0; }
    __syncthreads();
    {
        uint32_t real = 0;
        for (int i = tid; i < E; i += nthreads) real += (cands[i].x != 0u);
        real = __reduce_add_sync(FULL, real);
        if (lane == 0 && real) atomicAdd(&s_total, real);
    }
    __syncthreads();
    const uint32_t total_real = s_total;

    uint32_t T = 0;                                                  // threshold key
    if (total_real > (uint32_t)M) {
        // MSB-first radix select of the M-th largest key
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < 256; i += nthreads) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t pmask = pass ? (0xffffffffu << (shift + 8)) : 0u;
            for (int i = tid; i < E; i += nthreads) {
                const uint32_t k = cands[i].x;
                if (k != 0u && (k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t rem = s_remaining;
                int d = 255;
                for (; d > 0; --d) {
                    if (hist[d] >= rem) break;
                    rem -= hist[d];
                }
                s_prefix = prefix | ((uint32_t)d << shift);
                s_remaining = rem;
            }
            __syncthreads();
        }
        T = s_prefix;
    }
    // gather: keys > T all survive; keys == T fill the remaining places
    const uint32_t need_eq = (total_real > (uint32_t)M) ? s_remaining : 0xffffffffu;
    for (int i = tid; i < E; i += nthreads) {
        const uint2 c = make_uint2(cands[i].x, cands[i].y);
        if (c.x == 0u) continue;
        if (c.x > T || total_real <= (uint32_t)M) {
            a.surv_rows[atomicAdd(&s_cnt_gt, 1u)] = c.y;
        }
    }
    __syncthreads();
    const uint32_t n_gt = s_cnt_gt;
    if (total_real > (uint32_t)M) {
        for (int i = tid; i < E; i += nthreads) {
            const uint2 c = make_uint2(cands[i].x, cands[i].y);
            if (c.x == T) {
                const uint32_t slot = atomicAdd(&s_cnt_eq, 1u);
                if (slot < need_eq) a.surv_rows[n_gt + slot] = c.y;
            }
        }
    }
    __syncthreads();
    // tau: no row outside the survivor list has an fp32 score above it
    float tau = -INFINITY;
    for (int i = tid; i < (int)gridDim.x; i += nthreads)
        tau = fmaxf(tau, *((const volatile float*)&a.cta_floor[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tau = fmaxf(tau, __shfl_xor_sync(FULL, tau, o));
    if (lane == 0) s_wtau[warp] = tau;
    __syncthreads();
    if (tid == 0) {
        float t = (total_real > (uint32_t)M) ? key_to_float(T) : -INFINITY;
        for (int w = 0; w < a.warps; ++w) t = fmaxf(t, s_wtau[w]);
        if (q_unscreenable) t = INFINITY;
        const uint32_t n_eq = (total_real > (uint32_t)M) ? min(s_cnt_eq, need_eq) : 0u;
        a.sel[0] = (int32_t)(n_gt + n_eq);
        a.sel[1] = __float_as_int(t);
        a.sel[2] = 0;                                                // re-arm the ticket
        __threadfence();
    }
}

template <int NV, int TR, bool PIPE>
int launch_p(const ScanArgs& args, int grid, int smem, cudaStream_t st) {
    // opt in to the largest ring any dim needs (per device, once): the kernel is launched with `smem` <= that
    ORR_SMEM_OPT_IN((orr_scan_kernel<NV, TR, PIPE>), 227 * 1024 - 2048);
    orr_scan_kernel<NV, TR, PIPE><<<grid, args.warps * 32, smem, st>>>(args);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}
template <int NV, int TR>
int launch_t(const ScanArgs& args, int grid, int smem, cudaStream_t st) {
    // measured on B200 (profiles/r01_scan_variants.md): with 8 warps the row stream is the
    // bottleneck and the immediate epilogue (PIPE=0) is 5% faster; deferral only wins when
    // fewer warps fit (wide rows).
    static const int pipe = env_int("ORR_SCAN_PIPE", 0);
    return pipe ? launch_p<NV, TR, true>(args, grid, smem, st) : launch_p<NV, TR, false>(args, grid, smem, st);
}

}  // namespace

// smem layout: [query, 128-B padded][warps x stages x stage_bytes ring][mbarriers]; the ring
// is reused by the tail as merge scratch (256 x 8 B + per-warp floors), so it is never
// smaller than that.
static void scan_layout(int dim, int* warps, int* stages, int* tile_rows, int* stage_bytes, int* total) {
    int tr = 1;
    if (dim == 1536) tr = 2;
    if (dim == 768) tr = 4;
    int stage = (tr * dim * 4 + 127) & ~127;
    if (stage < 256) stage = 256;
    const int q_bytes = (dim * 4 + 127) & ~127;
    int w = env_int("ORR_SCAN_WARPS", ORR_SCAN_WARPS);
    int st = env_int("ORR_SCAN_STAGES", ORR_SCAN_STAGES);
    if (w < 1) w = 1;
    if (w > ORR_SCAN_WARPS) w = ORR_SCAN_WARPS;
    if (st < 2) st = 2;
    if (st > 8) st = 8;
    const int budget = 227 * 1024 - 2048;                          // minus the tail's static smem
    while (w > 1 && q_bytes + w * st * (stage + 8) > budget) --w;
    *warps = w; *stages = st; *tile_rows = tr; *stage_bytes = stage;
    *total = q_bytes + w * st * (stage + 8);
}

int orr_scan_smem_bytes(int dim, int* warps_out, int* tile_rows_out) {
    int w, st, tr, stage, total;
    scan_layout(dim, &w, &st, &tr, &stage, &total);
    if (warps_out) *warps_out = w;
    if (tile_rows_out) *tile_rows_out = tr;
    return total;
}

int orr_launch_scan(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                    const OrrWeights& w, int64_t now_ticks, int n_survivors, int grid,
                    cudaStream_t st) {
    ScanArgs a;
    a.sh = sh;
    a.q = sc.q;
    a.pr = pr;
    a.w_cos = (float)w.w_cos; a.w_kw = (float)w.w_kw; a.w_rec = (float)w.w_rec;
    a.decay_per_2p20 = (float)(1048576.0 / ((double)ORR_TICKS_PER_DAY * w.recency_days));
    a.inv_nterms = pr.n_terms > 0 ? 1.0f / (float)pr.n_terms : 0.f;
    a.now_ticks = now_ticks;
    int warps, stages, tr, stage, smem;
    scan_layout(sh.dim, &warps, &stages, &tr, &stage, &smem);
    if (warps * stages * stage < 256 * 8 + 64 || smem > 227 * 1024 - 2048) {
        orr_set_error("scan: dim %d does not fit the shared-memory ring", sh.dim);
        return ORR_E_UNSUPPORTED;
    }
    a.warps = warps;
    a.stages = stages;
    a.l2_hint = env_int("ORR_SCAN_L2_HINT", 1);
    a.stage_bytes = stage;
    a.n_surv = n_survivors;
    a.cta_cands = sc.cta_cands;
    a.cta_floor = sc.cta_floor;
    a.sel = sc.sel;
    a.surv_rows = sc.surv_rows;
    a.bm_bits = nullptr; a.bm_row_words = 0; a.n_bm = 0;
    a.dense = sc.scan_dense;
    if (sc.kw_bits) {
        if (sc.kw_terms < 1 || sc.kw_terms > 32 || pr.n_probes > 0) { orr_set_error("scan: bitmap terms must be 1..32 and exclusive of probes"); return ORR_E_INTERNAL; }
        a.bm_bits = sc.kw_bits; a.bm_row_words = sc.kw_row_words; a.n_bm = sc.kw_terms;
    }
    if (sh.dim == 3072) return launch_t<24, 1>(a, grid, smem, st);
    if (sh.dim == 1536) return launch_t<12, 2>(a, grid, smem, st);
    if (sh.dim == 768) return launch_t<6, 4>(a, grid, smem, st);
    return launch_t<0, 1>(a, grid, smem, st);
}
