// orr_exact.cu — the exact path: every live row scored with the reference's fp64 arithmetic, then the top-k selected
// under the reference's full ordering, all on own kernels (no library sort).
//
// Used when the fused fp32 scan does not apply or could not prove its selection: no query embedding — the
// reference's DEFAULT configuration (NoOpEmbeddingClient.cs:5-8 => cosine 0 everywhere, RecallSearchService.cs:71-72;
// ranking is keyword + recency only) —, top_k beyond the fused path's lists, text mode with > 32 terms, escalation.
//
//   E1  scores: every row's exact score -> one order-preserving 64-bit key per row (8 B/row) + the histogram of the
//       keys' top 12 bits.  Two kernels:
//         orr_noemb_scores_kernel<SPL>   no query embedding: reads only terms64 + ticks (8*slots + 8 B/row, HBM-bound)
//         orr_exact_scores_kernel        general: fp64 dot products over the embedding rows / text-mode bitmaps
//   E2  selection of the k best under (score desc with NaN last, CreatedAtUtc desc, row asc) — RecallSearchService.cs:34-37
//       with the stable-sort fallback — as an MSB-first radix select over the composite key
//       (score key 64 b | ticks key 64 b | ~row 32 b), 15 digits of <= 12 bits.  One launch per digit
//       (orr_sel_pass_kernel: every CTA re-derives the previous digit's choice from its histogram, then counts the next
//       digit of the rows still matching the prefix); the walk STOPS as soon as the rows at or above the chosen bin's
//       lower edge number <= 4096 (or exactly k): orr_sel_gather_kernel collects them, and its last CTA orders them
//       with a bitonic sort in shared memory and emits the hits.  On non-degenerate scores that is 2 digit passes
//       (23 bits of the score) over an 8 B/row array that is still in L2; massive exact ties (no terms, equal
//       timestamps) walk on into the ticks and row digits.  top_k > 4096 gathers exactly k rows and orders them with
//       a global-memory bitonic sort.
#include <algorithm>
#include <cstring>

#include "orr_exact_row.cuh"

namespace {

constexpr int SEL_BINS = 4096;                 // widest digit: 12 bits
constexpr int SEL_NPASS = 15;
constexpr int SEL_CAP = ORR_SORT_MAX;          // rows the in-CTA sorter orders
constexpr uint64_t SIGN64 = 0x8000000000000000ull;

struct SelDigit { int32_t word, shift, bits; };   // word 0 = score key, 1 = ticks key, 2 = ~row
__constant__ SelDigit c_digits[SEL_NPASS] = {
    {0, 52, 12}, {0, 41, 11}, {0, 30, 11}, {0, 20, 10}, {0, 10, 10}, {0, 0, 10},
    {1, 53, 11}, {1, 42, 11}, {1, 31, 11}, {1, 20, 11}, {1, 10, 10}, {1, 0, 10},
    {2, 21, 11}, {2, 10, 11}, {2, 0, 10}};

struct SelPoint {              // the walk after the digits of passes [0, i) have been chosen
    uint64_t p0, p1;           // prefix of the score key / ticks key (bits below the last chosen digit are 0)
    uint32_t p2;               // prefix of ~row
    uint32_t rem;              // how many of the rows matching the prefix the top-k still needs
    uint32_t n_above;          // rows strictly above the prefix's range: all of them are in the top-k
    uint32_t n_bin;            // rows matching the prefix
    int32_t  done;             // 1: {key >= prefix} is the candidate set (<= SEL_CAP rows, or exactly k); 2: inconsistent histogram
    int32_t  pad;
};
struct SelState {
    uint32_t hist[3][SEL_BINS];
    SelPoint pt[SEL_NPASS + 1];
    uint32_t n_gathered;
    int32_t  ticket;
    uint32_t n_verified;       // screened keys only: gathered rows whose EXACT key is still at or above the walk's lower edge
    int32_t  n_final;          // candidates the order kernel has to rank (0 until the walk has ended)
};

// order-preserving key of an exact score: larger key = ranks earlier; dead rows 0, NaN 1 (NaN sorts last, :34;
// Comparer<double> treats all NaNs as equal and -0.0 == 0.0, so those collapse to one key each)
__device__ __forceinline__ uint64_t score_key(double s, bool dead) {
    if (dead) return 0ull;
    if (s != s) return 1ull;
    s = __dadd_rn(s, 0.0);                                            // -0.0 -> +0.0
    const uint64_t b = (uint64_t)__double_as_longlong(s);
    return (b & SIGN64) ? ~b : (b | SIGN64);
}
__device__ __forceinline__ double key_score(uint64_t k) {
    if (k <= 1ull) return __longlong_as_double(0x7ff8000000000000LL);
    const uint64_t b = (k & SIGN64) ? (k ^ SIGN64) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ uint64_t ticks_key(int64_t t) { return (uint64_t)t ^ SIGN64; }

// one count into the CTA's shared histogram per distinct bin of the warp's active lanes
__device__ __forceinline__ void hist_add_warp(uint32_t* s_hist, uint32_t bin, uint32_t active) {
    const uint32_t peers = __match_any_sync(active, bin);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
}
__device__ __forceinline__ void hist_flush(uint32_t* g_hist, const uint32_t* s_hist, int bins) {
    for (int b = threadIdx.x; b < bins; b += blockDim.x) {
        const uint32_t c = s_hist[b];
        if (c) atomicAdd(&g_hist[b], c);
    }
}

// ---- E1, general form ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) orr_exact_scores_kernel(const ExactArgs a, uint64_t* skey, uint32_t* hist0) {
    __shared__ uint32_t s_hist[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool has_q = (a.q_dim == a.sh.dim && a.q_dim > 0);
    const double nA = has_q ? exact_qnorm(a, lane) : 0.0;
    // a warp takes 32 consecutive rows: the warp-collective part row by row (lane j keeps row j's partial result),
    // then the scalar fp64 tail once per lane and coalesced stores
    const int64_t n_blocks = (a.sh.rows + 31) >> 5;
    for (int64_t blk = gw; blk < n_blocks; blk += W) {
        const int64_t row0 = blk << 5;
        const int n_here = (int)min((int64_t)32, a.sh.rows - row0);
        ExactPartial mine;
        mine.dot = 0.0; mine.nB = 0.0; mine.matches = 0; mine.kw_den = -1; mine.ticks = 0;
        for (int j = 0; j < n_here; ++j) {
            const ExactPartial r = exact_row_partial(a, a.q, a.pr, row0 + j, lane);
            if (lane == j) mine = r;
        }
        const uint32_t active = __ballot_sync(FULL, lane < n_here);
        if (lane < n_here) {
            const uint64_t key = score_key(exact_row_finish(a, nA, mine), mine.ticks == ORR_DEAD_TICKS);
            skey[row0 + lane] = key;
            hist_add_warp(s_hist, (uint32_t)(key >> 52), active);
        }
    }
    __syncthreads();
    hist_flush(hist0, s_hist, SEL_BINS);
}

// ---- E1 without a query embedding: keyword + recency only --------------------------------------------------------
// A warp takes 32 consecutive rows = one contiguous 32 * slots * 4 B block of the 32-BIT term table (the scan's table:
// half the bytes of the 64-bit one), streamed with 8 coalesced 16-byte loads in flight per lane.  Every stored low word
// is compared with the probes' low words (kernel-parameter constant memory).  The per-row term masks are OR-reduced
// across the lanes that hold the row (one REDUX per 16-byte load, member masks split the warp when a load spans two
// rows), lane j keeps row j's count and runs the scalar fp64 tail (exact_row_finish: the same operations as every
// other path).
// SCREEN, NOT RESULT: a 32-bit word can collide with a probe, which can only RAISE a row's match count, so this key is an
// upper bound of the row's exact key.  The selection walk runs on these keys; the gather kernel recounts every candidate's
// matches on the 64-bit table (exact score, <= the screen's) and proves that k candidates still rank at or above the
// walk's lower edge — no row outside the candidate set can then belong to the top-k.  If a collision breaks the proof the
// caller re-runs with the 64-bit scoring kernel.  (Confirming hits inline was tried first: with Zipf terms most rows hold
// some query term, and a latency-serialised 64-bit load per hit made the kernel 3x slower than the table scan.)
// Algorithmic bytes per row: 4 * slots (terms32) + 8 (ticks); embeddings are never read.
// SPL = slots / 32: 16-byte vectors per row VR = 8 * SPL (8, 16, 32).  WIDE = more than 32 query terms (two mask words).
struct NoembArgs {
    OrrShard sh; OrrProbes pr; OrrWeights w; int64_t now_ticks;
};
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int SPL, bool WIDE>
__global__ void __launch_bounds__(512, 2) orr_noemb_scores_kernel(const NoembArgs a, uint64_t* skey, uint32_t* hist0) {
    __shared__ uint32_t s_hist[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_blocks = (a.sh.rows + 31) >> 5;
    constexpr int VR = 8 * SPL;                        // 16-byte vectors per row
    constexpr int RPL = 32 / VR;                       // rows one warp-wide load spans (4, 2, 1)
    constexpr int STEPS = (32 * VR) / 256;             // 8-load steps per 32-row block (1, 2, 4)
    const int n_probes = a.pr.n_probes;
    ExactLite ex;
    ex.sh = a.sh; ex.q_dim = 0; ex.w = a.w; ex.now_ticks = a.now_ticks;
    for (int64_t blk = gw; blk < n_blocks; blk += W) {
        const int64_t row0 = blk << 5;
        const int n_here = (int)min((int64_t)32, a.sh.rows - row0);
        const int n_vec = n_here * VR;
        const uint4* base = reinterpret_cast<const uint4*>(a.sh.terms32 + row0 * (int64_t)a.sh.slots);
        const int64_t my_ticks = lane < n_here ? __ldg(a.sh.ticks + row0 + lane) : ORR_DEAD_TICKS;
        int my_matches = 0;
        if (n_probes > 0) {
#pragma unroll 1
            for (int st = 0; st < STEPS; ++st) {
                const int v0 = st * 256;
                if (v0 >= n_vec) break;                                      // warp-uniform
                uint4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int v = v0 + i * 32 + lane;
                    x[i] = v < n_vec ? ldg_stream_u4(base + v) : make_uint4(0u, 0u, 0u, 0u);
                }
                uint32_t m0[8], m1[WIDE ? 8 : 1];
#pragma unroll
                for (int i = 0; i < 8; ++i) { m0[i] = 0u; if (WIDE) m1[i] = 0u; }
                // probe-major: one OR-chain of compares per probe over the lane's 32 words, and the per-load bookkeeping
                // only if SOME lane of the warp saw the probe in this step (a Zipf-tail term: almost never)
                for (int p = 0; p < n_probes; ++p) {
                    const uint32_t hl = a.pr.h32[p];
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < 8; ++i) any |= (x[i].x == hl) | (x[i].y == hl) | (x[i].z == hl) | (x[i].w == hl);
                    if (__any_sync(FULL, any)) {                             // warp-uniform
                        const uint32_t t = a.pr.term[p];
                        const uint32_t b0 = t < 32 ? (1u << t) : 0u, b1 = t < 32 ? 0u : (1u << (t - 32));
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool hit = (x[i].x == hl) | (x[i].y == hl) | (x[i].z == hl) | (x[i].w == hl);
                            m0[i] |= hit ? b0 : 0u;
                            if (WIDE) m1[i] |= hit ? b1 : 0u;
                        }
                    }
                }
                // OR across the lanes that hold one row (full-warp REDUX per segment: a partial member mask compiles to a
                // slow collective protocol), then the row's lane keeps the count
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int first_row = (st * 8 + i) * RPL;
#pragma unroll
                    for (int g = 0; g < RPL; ++g) {
                        const bool mine_seg = RPL == 1 || (lane / VR) == g;
                        int c = __popc(__reduce_or_sync(FULL, mine_seg ? m0[i] : 0u));
                        if (WIDE) c += __popc(__reduce_or_sync(FULL, mine_seg ? m1[i] : 0u));
                        if (lane == first_row + g) my_matches = c;
                    }
                }
            }
        }
        const uint32_t active = __ballot_sync(FULL, lane < n_here);
        if (lane < n_here) {
            ExactPartial mine;
            mine.dot = 0.0; mine.nB = 0.0; mine.ticks = my_ticks;
            mine.matches = my_matches; mine.kw_den = n_probes > 0 ? a.pr.n_terms : -1;
            const uint64_t key = score_key(exact_row_finish(ex, 0.0, mine), my_ticks == ORR_DEAD_TICKS);
            skey[row0 + lane] = key;
            hist_add_warp(s_hist, (uint32_t)(key >> 52), active);
        }
    }
    __syncthreads();
    hist_flush(hist0, s_hist, SEL_BINS);
}

// ---- E2: the digit walk ---------------------------------------------------------------------------------------------
__device__ __forceinline__ SelPoint sel_initial(uint32_t k) {
    SelPoint p;
    p.p0 = 0ull; p.p1 = 0ull; p.p2 = 0u; p.rem = k; p.n_above = 0u; p.n_bin = 0u; p.done = 0; p.pad = 0;
    return p;
}

// Chooses digit `pass` from its histogram: the bin d with  count(bins > d) < rem <= count(bins >= d).  Every thread of
// the CTA returns the same point.  blockDim.x must be a multiple of 32 (<= 1024).
__device__ SelPoint sel_advance(const SelPoint& prev, const uint32_t* hist, int pass) {
    __shared__ uint32_t s_wsum[32];
    __shared__ SelPoint s_out;
    if (prev.done) return prev;
    const SelDigit dg = c_digits[pass];
    const int bins = 1 << dg.bits;
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (bins + T - 1) / T;
    const int top = bins - 1 - tid * per;                                   // thread 0 owns the highest bins
    uint32_t sum = 0u;
    for (int j = 0; j < per; ++j) { const int b = top - j; if (b >= 0) sum += hist[b]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_wsum[warp] = incl;
    if (tid == 0) { s_out = prev; s_out.done = 2; }                          // stays if the histogram holds < rem rows
    __syncthreads();
    uint32_t before = 0u;
    for (int w = 0; w < warp; ++w) before += s_wsum[w];
    const uint32_t above = before + incl - sum;                              // rows in the bins of lower-numbered threads
    if (above < prev.rem && prev.rem <= above + sum) {
        uint32_t r = prev.rem - above, gt = above;
        int d = top;
        for (int j = 0; j < per; ++j) {
            const int b = top - j;
            const uint32_t h = b >= 0 ? hist[b] : 0u;
            if (h >= r) { d = b; break; }
            r -= h; gt += h;
        }
        SelPoint cur = prev;
        if (dg.word == 0) cur.p0 |= (uint64_t)d << dg.shift;
        else if (dg.word == 1) cur.p1 |= (uint64_t)d << dg.shift;
        else cur.p2 |= (uint32_t)d << dg.shift;
        cur.n_above = prev.n_above + gt;
        cur.rem = r;
        cur.n_bin = hist[d];
        const uint32_t m = cur.n_above + cur.n_bin;
        cur.done = (m <= (uint32_t)SEL_CAP || cur.n_bin == cur.rem || pass == SEL_NPASS - 1) ? 1 : 0;
        s_out = cur;
    }
    __syncthreads();
    const SelPoint out = s_out;
    __syncthreads();                                                         // s_out / s_wsum may be reused by the next call
    return out;
}

// counts digit `pass` of the rows whose higher digits equal the prefix
__device__ __forceinline__ void sel_accumulate(const SelPoint& cur, int pass, const uint64_t* skey, const int64_t* ticks,
                                               int64_t rows, uint32_t* g_hist, uint32_t* s_hist) {
    const SelDigit dg = c_digits[pass];
    const int bins = 1 << dg.bits;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const int upper = dg.shift + dg.bits;
    const uint32_t mask = (uint32_t)bins - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += stride) {
        const uint64_t w0 = skey[row];
        if (dg.word == 0) {
            if (upper >= 64 || (w0 >> upper) == (cur.p0 >> upper)) atomicAdd(&s_hist[(uint32_t)(w0 >> dg.shift) & mask], 1u);
        } else if (w0 == cur.p0) {
            const uint64_t w1 = ticks_key(ticks[row]);
            if (dg.word == 1) {
                if (upper >= 64 || (w1 >> upper) == (cur.p1 >> upper)) atomicAdd(&s_hist[(uint32_t)(w1 >> dg.shift) & mask], 1u);
            } else if (w1 == cur.p1) {
                const uint32_t w2 = ~(uint32_t)row;
                if (upper >= 32 || (w2 >> upper) == (cur.p2 >> upper)) atomicAdd(&s_hist[(w2 >> dg.shift) & mask], 1u);
            }
        }
    }
    __syncthreads();
    hist_flush(g_hist, s_hist, bins);
}

// pass >= 1: chooses digit pass-1, then counts digit `pass`
__global__ void __launch_bounds__(256) orr_sel_pass_kernel(SelState* st, int pass, uint32_t k, const uint64_t* skey,
                                                           const int64_t* ticks, int64_t rows) {
    __shared__ uint32_t s_hist[SEL_BINS];
    ORR_GRID_DEP_WAIT();                                              // the histogram / walk state of the kernel before this one
    ORR_GRID_DEP_LAUNCH();
    const SelPoint prev = (pass == 1) ? sel_initial(k) : st->pt[pass - 1];
    const SelPoint cur = sel_advance(prev, st->hist[(pass - 1) % 3], pass - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) st->pt[pass] = cur;
    if (cur.done) return;
    if (blockIdx.x == 0)                                                     // the buffer the NEXT pass counts into
        for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[(pass + 1) % 3][i] = 0u;
    sel_accumulate(cur, pass, skey, ticks, rows, st->hist[pass % 3], s_hist);
}

constexpr int32_t SEL_FLAG_INCOMPLETE = 8;      // the walk has not ended: the host launches more digit passes
constexpr int32_t SEL_FLAG_INTERNAL = 32;       // histogram inconsistent with k (bug)
constexpr int32_t SEL_FLAG_UNPROVEN = 64;       // a 32-bit hash collision inside the candidate set: re-run with the 64-bit scoring kernel

struct GatherArgs {
    SelState* st; int32_t last_pass; uint32_t k;
    const uint64_t* skey; const int64_t* ticks; int64_t rows; uint64_t row_base;
    // keys that are a 32-bit-hash SCREEN (no-embedding path): every gathered row's matches are recounted on the 64-bit
    // table and its exact score replaces the screen's
    int32_t verify; const uint64_t* terms64; int32_t slots; OrrProbes pr; OrrWeights w; int64_t now_ticks;
    OrrExact* big; uint32_t big_cap;            // top_k > SEL_CAP: the k selected rows go here (padded to a power of two)
    OrrExact* small;                            // otherwise: <= SEL_CAP candidates, ordered by the last CTA
    orr_hit* hits; int32_t* status;
};

__global__ void __launch_bounds__(512) orr_sel_gather_kernel(const GatherArgs a) {
    ORR_GRID_DEP_WAIT();
    ORR_GRID_DEP_LAUNCH();
    const int tid = threadIdx.x;
    SelState* st = a.st;
    const SelPoint prev = (a.last_pass == 0) ? sel_initial(a.k) : st->pt[a.last_pass];
    const SelPoint cur = sel_advance(prev, st->hist[a.last_pass % 3], a.last_pass);
    if (blockIdx.x == 0 && tid == 0) st->pt[a.last_pass + 1] = cur;
    if (cur.done != 1) {
        if (blockIdx.x == 0 && tid == 0) { a.status[0] = 0; a.status[1] = cur.done == 0 ? SEL_FLAG_INCOMPLETE : SEL_FLAG_INTERNAL; }
        return;
    }
    const bool big = a.big != nullptr;
    OrrExact* out = big ? a.big : a.small;
    const uint32_t cap = big ? a.big_cap : (uint32_t)SEL_CAP;
    const bool deep = (cur.p1 | (uint64_t)cur.p2) != 0ull;                   // the prefix reaches into the ticks / row digits
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + tid; row < a.rows; row += stride) {
        const uint64_t w0 = a.skey[row];
        if (w0 < cur.p0) continue;
        int64_t tk = 0; bool have_tk = false;
        bool take = w0 > cur.p0;
        if (!take) {
            if (!deep) take = true;
            else {
                tk = a.ticks[row]; have_tk = true;
                const uint64_t w1 = ticks_key(tk);
                take = w1 > cur.p1 || (w1 == cur.p1 && ~(uint32_t)row >= cur.p2);
            }
        }
        if (!take) continue;
        if (!have_tk) tk = a.ticks[row];
        double score = key_score(w0);
        if (a.verify && tk != ORR_DEAD_TICKS) {
            // exact recount on the full hashes: one thread per candidate, 16-byte loads over the row's slots
            uint32_t m0 = 0u, m1 = 0u;
            const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(a.terms64 + row * (int64_t)a.slots);
            // 8 loads in flight per step: one by one, the 32 loads of a 64-slot row were a chain of 32 L2 round trips
            // (the kernel spent 20 us on a few hundred candidates)
            for (int v0 = 0; v0 < a.slots / 2; v0 += 8) {
                ulonglong2 h2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h2[i] = __ldg(t2 + v0 + i);        // slots is a multiple of 32: no tail
                for (int p = 0; p < a.pr.n_probes; ++p) {
                    const uint64_t h = a.pr.h64[p];
                    bool hit = false;
#pragma unroll
                    for (int i = 0; i < 8; ++i) hit |= (h2[i].x == h) | (h2[i].y == h);
                    if (hit) { const uint32_t t = a.pr.term[p]; if (t < 32) m0 |= 1u << t; else m1 |= 1u << (t - 32); }
                }
            }
            ExactLite ex = {};
            ex.q_dim = 0; ex.w = a.w; ex.now_ticks = a.now_ticks;
            ExactPartial pt;
            pt.dot = 0.0; pt.nB = 0.0; pt.ticks = tk; pt.matches = __popc(m0) + __popc(m1);
            pt.kw_den = a.pr.n_probes > 0 ? a.pr.n_terms : -1;
            score = exact_row_finish(ex, 0.0, pt);
            // does the row still rank at or above the walk's lower edge with its exact score?
            const uint64_t e0 = score_key(score, false);
            bool ge = e0 > cur.p0;
            if (!ge && e0 == cur.p0) { const uint64_t w1 = ticks_key(tk); ge = w1 > cur.p1 || (w1 == cur.p1 && ~(uint32_t)row >= cur.p2); }
            if (ge) atomicAdd(&st->n_verified, 1u);
        }
        const uint32_t slot = atomicAdd(&st->n_gathered, 1u);
        if (slot < cap) { OrrExact r; r.score = score; r.ticks = tk; r.row = (uint64_t)row; out[slot] = r; }
    }
    // ---- the last CTA to finish orders the candidates (or pads the big list for the global sort) ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t m = min(*((volatile uint32_t*)&st->n_gathered), cap);
    // screened keys: the candidate set is complete iff k of its rows still rank at or above the lower edge after the recount
    const int32_t unproven = (a.verify && *((volatile uint32_t*)&st->n_verified) < min(a.k, (uint32_t)m)) ? SEL_FLAG_UNPROVEN : 0;
    if (big) {
        uint32_t np2 = 1u;
        while (np2 < m) np2 <<= 1;
        for (uint32_t i = m + tid; i < np2 && i < cap; i += blockDim.x) {
            OrrExact v; v.score = __longlong_as_double(0x7ff8000000000000LL); v.ticks = INT64_MIN; v.row = ~0ull;
            out[i] = v;
        }
        if (tid == 0) { a.status[0] = (int32_t)m; a.status[1] = unproven; }
        return;
    }
    // <= SEL_CAP candidates: orr_order_kernel (the caller's next launch) puts them in reference order; here only the counts
    if (tid == 0) { st->n_final = (int32_t)m; a.status[0] = (int32_t)min(a.k, m); a.status[1] = unproven; }
}

// ---- top_k > SEL_CAP: global-memory bitonic sort of the k selected rows, then the hits ---------------------------------
__global__ void orr_big_bitonic_step(OrrExact* e, uint32_t n, uint32_t k2, uint32_t j) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = i ^ j;
    if (p > i) {
        const OrrExact x = e[i], y = e[p];
        const bool up = ((i & k2) == 0);
        if (up ? ranks_before(y, x) : ranks_before(x, y)) { e[i] = y; e[p] = x; }
    }
}
__global__ void orr_big_emit(const OrrExact* e, const int32_t* status, uint32_t k, uint64_t row_base, orr_hit* hits) {
    const uint32_t n = min(k, (uint32_t)status[0]);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        orr_hit h; h.row = row_base + e[i].row; h.score = e[i].score; h.created_ticks = e[i].ticks;
        hits[i] = h;
    }
}

}  // namespace

size_t orr_exact_state_bytes() { return sizeof(SelState); }

// true: the keys orr_launch_exact_scores writes for this query are the 32-bit screen (the gather must verify)
bool orr_exact_keys_are_screened(const OrrScratch& sc, const OrrProbes& pr, int q_dim, bool force_general) {
    return !force_general && q_dim == 0 && sc.kw_bits == nullptr && pr.n_probes > 0;
}

int orr_launch_exact_scores(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                            const OrrWeights& w, int64_t now_ticks, int q_dim, cudaStream_t st, bool force_general) {
    if (sh.rows == 0) return ORR_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SelState* state = reinterpret_cast<SelState*>(sc.sel_state);
    ORR_CUDA_OK(cudaMemsetAsync(state, 0, sizeof(SelState), st));
    if (q_dim == 0 && sc.kw_bits == nullptr && !force_general) {
        NoembArgs a;
        a.sh = sh; a.pr = pr; a.w = w; a.now_ticks = now_ticks;
        const int grid = sms * 2;
        const bool wide = pr.n_terms > 32;
#define ORR_NOEMB_LAUNCH(SPL)                                                                                         \
        do {                                                                                                              \
            if (wide) orr_noemb_scores_kernel<SPL, true><<<grid, 512, 0, st>>>(a, sc.skey, state->hist[0]);             \
            else orr_noemb_scores_kernel<SPL, false><<<grid, 512, 0, st>>>(a, sc.skey, state->hist[0]);                  \
        } while (0)
        if (sh.slots == 32) ORR_NOEMB_LAUNCH(1);
        else if (sh.slots == 64) ORR_NOEMB_LAUNCH(2);
        else ORR_NOEMB_LAUNCH(4);
#undef ORR_NOEMB_LAUNCH
    } else {
        ExactArgs e;
        fill_exact_args(e, sh, sc, pr, w, now_ticks, q_dim);
        orr_exact_scores_kernel<<<sms * 8, 256, 0, st>>>(e, sc.skey, state->hist[0]);
    }
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

// Digit passes [first_pass, first_pass + n_passes) (pass 0 was counted by E1), then the gather.  status[1] carries
// ORR_EXACT_INCOMPLETE when the walk needs more passes: the caller reads it back and calls again with the next passes.
int orr_launch_exact_select(const OrrShard& sh, const OrrScratch& sc, int top_k, int first_pass, int n_passes, cudaStream_t st,
                            const OrrProbes* verify_pr, const OrrWeights* verify_w, int64_t now_ticks) {
    const int64_t n = sh.rows;
    if (n == 0) { ORR_CUDA_OK(cudaMemsetAsync(sc.status, 0, 2 * sizeof(int32_t), st)); return ORR_OK; }
    if (n > 0x7fffffff) { orr_set_error("exact path: shard too large"); return ORR_E_UNSUPPORTED; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SelState* state = reinterpret_cast<SelState*>(sc.sel_state);
    const uint32_t k = (uint32_t)std::max(1, top_k);
    int last = std::max(0, first_pass - 1);
    for (int p = std::max(1, first_pass); p < first_pass + n_passes && p < SEL_NPASS; ++p) {
        const int grid = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)sms * 4);
        // the first pass of a round follows the scoring kernel (or the host's read of the previous round): a plain launch
        if (p == std::max(1, first_pass)) orr_sel_pass_kernel<<<std::max(1, grid), 256, 0, st>>>(state, p, k, sc.skey, sh.ticks, n);
        else ORR_CUDA_OK(orr_launch_dependent(orr_sel_pass_kernel, std::max(1, grid), 256, st, state, p, k, sc.skey, sh.ticks, n));
        ORR_CUDA_OK(cudaGetLastError());
        last = p;
    }
    GatherArgs g;
    g.st = state; g.last_pass = last; g.k = k; g.skey = sc.skey; g.ticks = sh.ticks; g.rows = n; g.row_base = sh.row_base;
    g.big = nullptr; g.big_cap = 0; g.small = sc.exact; g.hits = sc.hits; g.status = sc.status;
    g.verify = verify_pr != nullptr ? 1 : 0;
    g.terms64 = sh.terms64; g.slots = sh.slots; g.now_ticks = now_ticks;
    if (verify_pr) { g.pr = *verify_pr; g.w = *verify_w; } else { memset(&g.pr, 0, sizeof g.pr); memset(&g.w, 0, sizeof g.w); }
    const bool big = k > (uint32_t)SEL_CAP;
    if (big) {
        if (!sc.big || sc.big_cap < k) { orr_set_error("exact path: big-k buffer missing"); return ORR_E_INTERNAL; }
        g.big = sc.big; g.big_cap = (uint32_t)sc.big_cap;
    }
    // the gather resets nothing: n_gathered / ticket are zero from E1's memset unless an earlier round already gathered
    // (it did not: a round that gathers ends the search)
    const int grid = (int)std::min<int64_t>((n + 2047) / 2048, (int64_t)sms * 2);
    if (last >= std::max(1, first_pass)) ORR_CUDA_OK(orr_launch_dependent(orr_sel_gather_kernel, std::max(1, grid), 512, st, g));
    else orr_sel_gather_kernel<<<std::max(1, grid), 512, 0, st>>>(g);
    ORR_CUDA_OK(cudaGetLastError());
    if (!big) {
        // n_final stays 0 while the walk is incomplete: the order kernel's warps then all exit at once
        const int rc = orr_launch_order(sc.exact, &state->n_final, SEL_CAP, (int)k, sh.row_base, sc.hits, st, true);
        if (rc != ORR_OK) return rc;
    }
    if (big) {
        // valid only once the walk is done (status[1] == 0); sorting an unfinished list is harmless and is redone
        uint32_t np2 = 1u;
        while (np2 < k) np2 <<= 1;
        for (uint32_t k2 = 2; k2 <= np2; k2 <<= 1)
            for (uint32_t j = k2 >> 1; j > 0; j >>= 1) orr_big_bitonic_step<<<(np2 + 255) / 256, 256, 0, st>>>(sc.big, np2, k2, j);
        orr_big_emit<<<std::min<uint32_t>((k + 255) / 256, (uint32_t)sms * 4), 256, 0, st>>>(sc.big, sc.status, k, sh.row_base, sc.hits);
        ORR_CUDA_OK(cudaGetLastError());
    }
    return ORR_OK;
}
