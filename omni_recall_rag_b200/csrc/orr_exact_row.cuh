// orr_exact_row.cuh — the ONE definition of a chunk's exact score on the device, shared by K3 (orr_rescore.cu),
// the batched finalize kernel and the exact path (orr_exact.cu), so every path produces the same bits.
//
// Restates, with every fp64 operation spelled as a round-to-nearest intrinsic (no FMA contraction):
//   CosineSimilarity  RecallSearchService.cs:69-88   KeywordScore :110-112   RecencyScore :115-119   ScoreChunk :66
#pragma once
#include <cfloat>

#include "orr_internal.h"

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// the fields exact_row_finish / exact_row_partial read when there is no per-call query pointer or probe set
struct ExactLite {
    OrrShard sh;
    int32_t q_dim;
    OrrWeights w;
    int64_t now_ticks;
};

struct ExactArgs {
    OrrShard  sh;
    const float* q;
    int32_t   q_dim;             // 0 => no query embedding (cosine 0, :71)
    OrrProbes pr;
    OrrWeights w;
    int64_t   now_ticks;
    // text mode (orr_search_text): the keyword matches come from per-term row bitmaps produced by the
    // substring kernel (orr_textmatch.cu) instead of the hashed term table
    const uint32_t* kw_bits;     // [kw_terms][kw_row_words] or NULL
    int64_t   kw_row_words;
    int32_t   kw_terms;
};
template <class A> __device__ __forceinline__ const uint32_t* kw_bits_of(const A&) { return nullptr; }
__device__ __forceinline__ const uint32_t* kw_bits_of(const ExactArgs& a) { return a.kw_bits; }
template <class A> __device__ __forceinline__ int kw_terms_of(const A&) { return 0; }
__device__ __forceinline__ int kw_terms_of(const ExactArgs& a) { return a.kw_terms; }
template <class A> __device__ __forceinline__ int kw_count_from_bits(const A&, int64_t, int) { return 0; }
__device__ __forceinline__ int kw_count_from_bits(const ExactArgs& a, int64_t row, int lane) {
    int cnt = 0;
    for (int t = lane; t < a.kw_terms; t += 32)
        cnt += (int)((__ldg(a.kw_bits + (int64_t)t * a.kw_row_words + (row >> 5)) >> (row & 31)) & 1u);
    return __reduce_add_sync(0xffffffffu, cnt);
}

// Loads are issued in batches of EX_CHUNK float4 per lane BEFORE the dependent fp64 chains
// so a row costs ~2 memory round trips instead of one per 128 columns.  ORDER of the fp64 additions (the one departure
// from the reference's sequential loop, a few 1e-16 relative): every lane keeps one chain per float4 component over its
// lane-strided columns in increasing order, the four chains are combined as (c0 + c1) + (c2 + c3), then the shuffle
// butterfly.  Every path (K3, batched finalize, exact path, ||q||^2) uses this one definition.
constexpr int EX_CHUNK = 12;

// fp64 ||q||^2 in the lane-strided order; all lanes return the same value
template <class A>
__device__ __forceinline__ double exact_qnorm_q(const A& a, const float* q, int lane) {
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;              // one chain per float4 component, as in exact_row_partial
    const int nv4 = a.sh.dim >> 2;
    const float4* q4 = reinterpret_cast<const float4*>(q);
    for (int base = 0; base < nv4; base += 32 * EX_CHUNK) {
        float4 v[EX_CHUNK];
#pragma unroll
        for (int j = 0; j < EX_CHUNK; ++j) {
            const int i = base + j * 32 + lane;
            v[j] = (i < nv4) ? __ldg(q4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < EX_CHUNK; ++j) {
            if (base + j * 32 + lane < nv4) {
                n0 = __dadd_rn(n0, (double)__fmul_rn(v[j].x, v[j].x));
                n1 = __dadd_rn(n1, (double)__fmul_rn(v[j].y, v[j].y));
                n2 = __dadd_rn(n2, (double)__fmul_rn(v[j].z, v[j].z));
                n3 = __dadd_rn(n3, (double)__fmul_rn(v[j].w, v[j].w));
            }
        }
    }
    return warp_sum_f64(__dadd_rn(__dadd_rn(n0, n1), __dadd_rn(n2, n3)));
}

// The exact fused score of one row in two steps, so that callers holding many rows per warp can run the scalar
// fp64 tail (2 sqrt, 4 divides, exp: ~200 instructions) once per LANE instead of once per warp:
//   exact_row_partial — warp-collective: fp64 dot and ||b||^2 (lane-strided order + butterfly), keyword matches,
//                       ticks; every lane returns the same values;
//   exact_row_finish  — per thread: CosineSimilarity's tail, KeywordScore's ratio, RecencyScore, ScoreChunk.
// exact_row_q = finish(partial): every path computes a row's score with the same operations on the same values.
struct ExactPartial {
    double  dot, nB;
    int32_t matches;             // distinct query terms the row's content holds
    int32_t kw_den;              // KeywordScore's denominator (-1 = no keyword side)
    int64_t ticks;
};

// CHUNK = float4 loads per lane issued before the dependent fp64 chains; Q_SHARED = q lives in shared memory.
template <class A, int CHUNK = EX_CHUNK, bool Q_SHARED = false, class P = OrrProbes>
__device__ __forceinline__ ExactPartial exact_row_partial(const A& a, const float* q, const P& pr, int64_t row, int lane) {
    ExactPartial r;
    r.ticks = a.sh.ticks[row];
    // term hashes are fetched up front so their latency overlaps the embedding loads
    uint64_t th[4] = {0, 0, 0, 0};
    const int n_probes = orr_probe_count(pr);
    if (n_probes > 0) {
        const int spl = a.sh.slots >> 5;
        const uint64_t* t64 = a.sh.terms64 + row * (int64_t)a.sh.slots;
#pragma unroll
        for (int w = 0; w < 4; ++w) if (w < spl) th[w] = __ldg(t64 + w * 32 + lane);
    }
    double dot = 0.0, nB = 0.0;
    if (a.q_dim == a.sh.dim && a.q_dim > 0) {                       // :71-72 length check
        // 4 + 4 independent fp64 chains per lane (one per float4 component): the per-lane work is a latency chain of
        // dependent DADDs, and two chains left the fp64 pipe idle most of the time
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        const int nv4 = a.sh.dim >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(a.sh.emb + row * (int64_t)a.sh.dim);
        const float4* q4 = reinterpret_cast<const float4*>(q);
        for (int base = 0; base < nv4; base += 32 * CHUNK) {
            float4 x[CHUNK];
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const int i = base + j * 32 + lane;
                x[j] = (i < nv4) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (Q_SHARED) {
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    const int i = base + j * 32 + lane;
                    if (i < nv4) {
                        const float4 v = q4[i];
                        d0 = __dadd_rn(d0, (double)__fmul_rn(v.x, x[j].x)); b0 = __dadd_rn(b0, (double)__fmul_rn(x[j].x, x[j].x));
                        d1 = __dadd_rn(d1, (double)__fmul_rn(v.y, x[j].y)); b1 = __dadd_rn(b1, (double)__fmul_rn(x[j].y, x[j].y));
                        d2 = __dadd_rn(d2, (double)__fmul_rn(v.z, x[j].z)); b2 = __dadd_rn(b2, (double)__fmul_rn(x[j].z, x[j].z));
                        d3 = __dadd_rn(d3, (double)__fmul_rn(v.w, x[j].w)); b3 = __dadd_rn(b3, (double)__fmul_rn(x[j].w, x[j].w));
                    }
                }
            } else {
                float4 v[CHUNK];
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    const int i = base + j * 32 + lane;
                    v[j] = (i < nv4) ? __ldg(q4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    if (base + j * 32 + lane < nv4) {
                        d0 = __dadd_rn(d0, (double)__fmul_rn(v[j].x, x[j].x)); b0 = __dadd_rn(b0, (double)__fmul_rn(x[j].x, x[j].x));
                        d1 = __dadd_rn(d1, (double)__fmul_rn(v[j].y, x[j].y)); b1 = __dadd_rn(b1, (double)__fmul_rn(x[j].y, x[j].y));
                        d2 = __dadd_rn(d2, (double)__fmul_rn(v[j].z, x[j].z)); b2 = __dadd_rn(b2, (double)__fmul_rn(x[j].z, x[j].z));
                        d3 = __dadd_rn(d3, (double)__fmul_rn(v[j].w, x[j].w)); b3 = __dadd_rn(b3, (double)__fmul_rn(x[j].w, x[j].w));
                    }
                }
            }
        }
        dot = warp_sum_f64(__dadd_rn(__dadd_rn(d0, d1), __dadd_rn(d2, d3)));
        nB = warp_sum_f64(__dadd_rn(__dadd_rn(b0, b1), __dadd_rn(b2, b3)));
    }
    r.dot = dot; r.nB = nB;
    r.matches = 0; r.kw_den = -1;
    if (kw_bits_of(a) != nullptr) {
        r.matches = kw_count_from_bits(a, row, lane);
        r.kw_den = kw_terms_of(a);
    } else if (n_probes > 0) {                                      // :110-112
        uint32_t m0 = 0, m1 = 0;
        for (int p = 0; p < n_probes; ++p) {
            const uint64_t h = pr.h64[p];
            const bool hit = (th[0] == h) | (th[1] == h) | (th[2] == h) | (th[3] == h);
            if (hit) {
                const uint32_t t = orr_probe_term(pr, p);
                if (t < 32) m0 |= 1u << t; else m1 |= 1u << (t - 32);
            }
        }
        m0 = __reduce_or_sync(FULL, m0);
        m1 = __reduce_or_sync(FULL, m1);
        r.matches = __popc(m0) + __popc(m1);
        r.kw_den = pr.n_terms;
    }
    return r;
}

template <class A>
__device__ __forceinline__ double exact_row_finish(const A& a, double nA, const ExactPartial& r) {
    double cosv = 0.0;
    if (a.q_dim == a.sh.dim && a.q_dim > 0) {
        if (!(nA <= 0.0) && !(r.nB <= 0.0))                           // :84-85 (NaN falls through)
            cosv = __ddiv_rn(r.dot, __dmul_rn(__dsqrt_rn(nA), __dsqrt_rn(r.nB)));   // :87
    }
    const double kw = r.kw_den != -1 ? __ddiv_rn((double)r.matches, (double)r.kw_den) : 0.0;   // :112
    // RecencyScore: TimeSpan.TotalDays = ticks / 864e9; Math.Max(0, .); exp(-age/30)
    double age = __ddiv_rn((double)(a.now_ticks - r.ticks), 864000000000.0);
    if (!(age > 0.0)) age = 0.0;
    const double rec = exp(__ddiv_rn(-age, a.w.recency_days));
    // ScoreChunk :66
    return __dadd_rn(__dadd_rn(__dmul_rn(cosv, a.w.w_cos), __dmul_rn(kw, a.w.w_kw)),
                     __dmul_rn(rec, a.w.w_rec));
}

// exact fused score of one row, computed by a full warp; all lanes return the same value.
template <class A, int CHUNK = EX_CHUNK, bool Q_SHARED = false, class P = OrrProbes>
__device__ __forceinline__ double exact_row_q(const A& a, const float* q, const P& pr, int64_t row,
                                              int lane, double nA, int64_t* ticks_out) {
    const ExactPartial r = exact_row_partial<A, CHUNK, Q_SHARED, P>(a, q, pr, row, lane);
    *ticks_out = r.ticks;
    return exact_row_finish(a, nA, r);
}
__device__ __forceinline__ double exact_qnorm(const ExactArgs& a, int lane) { return exact_qnorm_q(a, a.q, lane); }
__device__ __forceinline__ double exact_row(const ExactArgs& a, int64_t row, int lane, double nA, int64_t* ticks_out) {
    return exact_row_q(a, a.q, a.pr, row, lane, nA, ticks_out);
}

// reference ordering: true if x ranks strictly before y
__device__ __forceinline__ bool ranks_before(const OrrExact& x, const OrrExact& y) {
    const bool xn = (x.score != x.score), yn = (y.score != y.score);
    if (xn != yn) return yn;                                          // NaN last (:34)
    if (!xn && x.score != y.score) return x.score > y.score;
    if (x.ticks != y.ticks) return x.ticks > y.ticks;                 // :35
    return x.row < y.row;                                             // stable fallback (A-6)
}

// Ordering of n <= ORR_RANK_MAX records held in shared memory WITHOUT sorting them: a record's position in the reference
// order is the number of records that rank before it (the keys are unique: the row index breaks every tie), so every
// thread counts that for its records (broadcast reads of e[j], no synchronisation) and writes the ones with
// position < n_out straight to their place.  O(n^2 / threads) comparisons against the O(log^2 n) barrier-separated
// stages of a bitonic network: 6x faster for the 64..512 records of K3, the exchange merge and the batched finalize.
// *kth_score receives the score at position k-1 (the bound checks need it); the caller synchronises before reading it.
constexpr int ORR_RANK_MAX = 1024;
__device__ __forceinline__ void rank_emit(const OrrExact* e, int n, int n_out, int k, uint64_t row_base, orr_hit* hits, double* kth_score) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const OrrExact x = e[i];
        int pos = 0;
        for (int j = 0; j < n; ++j) pos += ranks_before(e[j], x) ? 1 : 0;
        if (pos < n_out) { orr_hit h; h.row = row_base + x.row; h.score = x.score; h.created_ticks = x.ticks; hits[pos] = h; }
        if (pos == k - 1 && kth_score) *kth_score = x.score;
    }
}

inline void fill_exact_args(ExactArgs& e, const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                            const OrrWeights& w, int64_t now_ticks, int q_dim) {
    e.sh = sh; e.q = sc.q; e.q_dim = q_dim; e.pr = pr; e.w = w; e.now_ticks = now_ticks;
    e.kw_bits = sc.kw_bits; e.kw_row_words = sc.kw_row_words; e.kw_terms = sc.kw_terms;
}

}  // namespace
