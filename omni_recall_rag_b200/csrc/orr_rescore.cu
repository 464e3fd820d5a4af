// orr_rescore.cu — K3 and the exact path: the reference's fp64 arithmetic on the device.
//
// exact_row() is the one definition of a chunk's score on the GPU.  It restates
//   CosineSimilarity  RecallSearchService.cs:69-88  (fp32 products widened to fp64, fp64 sums,
//                                                    sqrt(nA)*sqrt(nB), one divide)
//   KeywordScore      :110-112  (matches / |terms| over the chunk's 64-bit term hashes)
//   RecencyScore      :115-119  (ticks -> days -> exp(-age/30))
//   ScoreChunk        :66       ((cos*0.7 + kw*0.2) + rec*0.1, no FMA contraction)
// with every fp64 operation spelled as a round-to-nearest intrinsic so nvcc cannot fuse
// multiply-adds.  The only departure from the reference's instruction stream is the ORDER of
// the fp64 additions inside the dot products (lane-strided partial sums + a fixed shuffle
// butterfly instead of one sequential chain), a difference of a few 1e-16 relative, and the
// CUDA libm exp (<= 1 ulp) instead of the host's: scores agree with the CPU oracle to ~1e-15
// relative, five orders of magnitude inside the 1e-5 contract.
//
// K3 = orr_rescore_kernel (warp per listed row: the exact score) + orr_order_kernel (warp per record: its position
// under the reference tie chain — score desc with NaN last, CreatedAtUtc desc, row asc, :34-35 plus the stable-sort
// fallback, SURVEY.md A-6 —, the hits and the selection bound check).
#include <cfloat>
#include <cuda_fp16.h>

#include "orr_internal.h"

#include "orr_exact_row.cuh"

namespace {

struct RescoreArgs {
    ExactArgs ex;
    const uint32_t* rows;        // listed local rows
    const int32_t*  n_listed;    // device count (sel[0]), or NULL: n_listed_value
    const int32_t*  tau_bits;    // device float bits (sel[1])
    int32_t   n_listed_max;
    int32_t   n_listed_value;    // the count when the host knows it (subset path): saves a 4-byte H2D copy per query
    int32_t   top_k;
    int32_t   check_bound;
    double    eps;
    OrrExact* exact;
    int32_t*  ticket;            // sel[3]
    orr_hit*  hits;
    int32_t*  status;
};

__global__ void __launch_bounds__(128) orr_rescore_kernel(const RescoreArgs a) {
    // programmatic dependent launch: the ordering kernel behind this one may be scheduled now; its CTAs wait in
    // griddepcontrol.wait until this grid has completed and its stores are visible (saves the launch gap between the two
    // kernels of K3: 2-3 us of a 24 us small-store query)
    ORR_GRID_DEP_LAUNCH();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = min(a.n_listed ? *a.n_listed : a.n_listed_value, a.n_listed_max);
    const int idx = blockIdx.x * 4 + warp;
    if (idx < n) {
        const double nA = (a.ex.q_dim == a.ex.sh.dim && a.ex.q_dim > 0) ? exact_qnorm(a.ex, lane) : 0.0;
        const int64_t row = a.rows[idx];
        int64_t ticks;
        const double s = exact_row(a.ex, row, lane, nA, &ticks);
        if (lane == 0) { a.exact[idx].score = s; a.exact[idx].ticks = ticks; a.exact[idx].row = (uint64_t)row; }
    }
}

// The records' reference order, the hits and (K3) the selection bound check.  One WARP per record, on as many SMs as
// there are records: the record's position is the number of records that rank before it (keys are unique: the row breaks
// every tie); the lanes split that count 32 records at a time and the warp STOPS as soon as k records rank before its
// own — it cannot be a hit — so the cost is ~32 comparisons for most records and n only for the k hits.
// Ordering used to be the tail of orr_rescore_kernel, run by the last CTA to finish: 4 warps on one SM working through a
// latency chain (49 us for the reference's 300 candidates, 3x the scoring itself; profiles/r02_kernels.md); as its own
// launch it costs ~3 us.  The exact path orders its gathered candidates (<= 4096) with the same kernel.
struct OrderArgs {
    const OrrExact* exact;
    const int32_t*  n_ptr;       // device count, or NULL: n_value
    int32_t   n_value, n_max, top_k;
    int32_t   check_bound;       // K3: prove the fp32 selection (needs tau_bits, eps)
    const int32_t* tau_bits;
    double    eps;
    uint64_t  row_base;
    orr_hit*  hits;
    int32_t*  status;            // {n_out, flags}; NULL: the caller has written it
};

constexpr int ORDER_STAGE = 1024;               // records staged in shared memory per round (24 KB)
__global__ void __launch_bounds__(256) orr_order_kernel(const OrderArgs a) {
    // every warp compares its record with ALL records: the list is staged in shared memory ORDER_STAGE records at a
    // time (one coalesced round trip per CTA) instead of each warp walking it through L2 32 records per round trip
    // (10 dependent round trips for the reference's 300 candidates: the kernel took 9-10 us, now one round trip)
    __shared__ OrrExact stage[ORDER_STAGE];
    ORR_GRID_DEP_WAIT();                                              // no-op unless launched as a programmatic dependent
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int n = min(a.n_ptr ? *a.n_ptr : a.n_value, a.n_max);
    const int k = max(1, a.top_k);                                    // Math.Max(1, topK) :36
    const int n_out = min(k, n);
    const float tau = a.check_bound ? __int_as_float(*a.tau_bits) : -INFINITY;
    if (i == 0 && lane == 0 && a.status) {
        a.status[0] = n_out;
        // fewer rows listed than k: no k-th score exists, the bound cannot be proven (unless nothing was discarded)
        if (n < k) a.status[1] = (a.check_bound && tau != -INFINITY) ? 1 : 0;
    }
    if ((int)((blockIdx.x * blockDim.x) >> 5) >= n) return;           // the whole CTA is beyond the list (CTA-uniform)
    const bool mine = i < n;
    OrrExact x;
    if (mine) { x.score = a.exact[i].score; x.ticks = a.exact[i].ticks; x.row = a.exact[i].row; }
    else { x.score = 0.0; x.ticks = 0; x.row = 0; }
    int pos = 0;
    bool live = mine;                                                 // false once k records rank before this one
    for (int base = 0; base < n; base += ORDER_STAGE) {
        const int m = min(ORDER_STAGE, n - base);
        if (!__syncthreads_or(live ? 1 : 0)) break;                   // previous round's readers are done; nobody left: stop staging
        for (int j = threadIdx.x; j < m; j += blockDim.x) stage[j] = a.exact[base + j];
        __syncthreads();
        if (live) {
            for (int j0 = 0; j0 < m; j0 += 32) {
                const int j = j0 + lane;
                const int c = (j < m && ranks_before(stage[j], x)) ? 1 : 0;
                pos += __reduce_add_sync(FULL, c);
                if (pos >= k) { live = false; break; }                // warp-uniform: not among the first k
            }
        }
    }
    if (!live || lane != 0) return;
    if (pos < n_out) { orr_hit h; h.row = a.row_base + x.row; h.score = x.score; h.created_ticks = x.ticks; a.hits[pos] = h; }
    if (pos == k - 1 && a.status) {
        // every row outside the list has fp32 score <= tau and |fp32 - exact| <= eps: safe iff the k-th exact score clears
        // tau by more than eps
        int flags = 0;
        if (a.check_bound && tau != -INFINITY && !(x.score - a.eps > (double)tau)) flags |= 1;
        a.status[1] = flags;
    }
}

// ---- merge of all-gathered per-GPU hit lists (multi-GPU) -----------------------------------
// lists[G][stride] hits + status[G][2] ({n, flags}); one CTA orders the union by the reference
// tie chain and writes the global top-k.  flags are OR-ed so a failed bound check on any
// shard is visible to the caller.
__device__ __forceinline__ void merge_lists_cta(const orr_hit* lists, const int32_t* status, int n_lists, int stride,
                                                int top_k, orr_hit* out, int32_t* out_status, int extra_flags,
                                                OrrExact* e) {
    const int tid = threadIdx.x;
    const int total = n_lists * stride;
    __shared__ int s_n, s_flags;
    if (tid == 0) { s_n = 0; s_flags = extra_flags; }
    __syncthreads();
    for (int i = tid; i < total; i += blockDim.x) {                  // the valid hits, compacted (their order does not matter)
        const int l = i / stride, j = i - l * stride;
        if (j < status[2 * l]) {
            const orr_hit h = lists[i];
            OrrExact v; v.score = h.score; v.ticks = h.created_ticks; v.row = h.row;
            e[atomicAdd(&s_n, 1)] = v;
        }
    }
    if (tid < n_lists) atomicOr(&s_flags, status[2 * tid + 1]);
    __syncthreads();
    const int n = s_n;
    const int n_out = min(max(1, top_k), n);
    if (n <= ORR_RANK_MAX) {
        rank_emit(e, n, n_out, max(1, top_k), 0ull, out, nullptr);
    } else {
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        for (int i = n + tid; i < np2; i += blockDim.x) { OrrExact v; v.score = __longlong_as_double(0x7ff8000000000000LL); v.ticks = INT64_MIN; v.row = ~0ull; e[i] = v; }
        __syncthreads();
        for (int k2 = 2; k2 <= np2; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < np2; i += blockDim.x) {
                    const int p = i ^ j;
                    if (p > i) {
                        const OrrExact x = e[i], y = e[p];
                        const bool up = ((i & k2) == 0);
                        if (up ? ranks_before(y, x) : ranks_before(x, y)) { e[i] = y; e[p] = x; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < n_out; i += blockDim.x) {
            orr_hit h; h.row = e[i].row; h.score = e[i].score; h.created_ticks = e[i].ticks;
            out[i] = h;
        }
    }
    if (tid == 0) { out_status[0] = n_out; out_status[1] = s_flags; }
}

__global__ void __launch_bounds__(256) orr_merge_kernel(const orr_hit* lists, const int32_t* status, int n_lists,
                                                        int stride, int top_k, orr_hit* out, int32_t* out_status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    merge_lists_cta(lists, status, n_lists, stride, top_k, out, out_status, 0, reinterpret_cast<OrrExact*>(smem_raw));
}

// batched form (row-sharded orr_search_batch): one CTA per query merges the query's n_lists gathered lists.
// lists[l][b][k], n[l][b] -> out[b][k], n_out[b]
__global__ void __launch_bounds__(128) orr_merge_batch_kernel(const orr_hit* lists, const int32_t* n, int n_lists, int batch,
                                                              int k, orr_hit* out, int32_t* n_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    OrrExact* e = reinterpret_cast<OrrExact*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int total = n_lists * k;
    int np2 = 1;
    while (np2 < total) np2 <<= 1;
    __shared__ int s_n;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < np2; i += blockDim.x) {
        OrrExact v; v.score = __longlong_as_double(0x7ff8000000000000LL); v.ticks = INT64_MIN; v.row = ~0ull;
        if (i < total) {
            const int l = i / k, j = i - l * k;
            if (j < n[(int64_t)l * batch + b]) {
                const orr_hit h = lists[((int64_t)l * batch + b) * k + j];
                v.score = h.score; v.ticks = h.created_ticks; v.row = h.row;
                atomicAdd(&s_n, 1);
            }
        }
        e[i] = v;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= np2; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const OrrExact x = e[i], y = e[p];
                    const bool up = ((i & k2) == 0);
                    if (up ? ranks_before(y, x) : ranks_before(x, y)) { e[i] = y; e[p] = x; }
                }
            }
            __syncthreads();
        }
    }
    const int m = min(k, s_n);
    for (int i = tid; i < m; i += blockDim.x) {
        orr_hit h; h.row = e[i].row; h.score = e[i].score; h.created_ticks = e[i].ticks;
        out[(int64_t)b * k + i] = h;
    }
    if (tid == 0) n_out[b] = m;
}

// ---- fused all-gather + merge over NVLink peer memory (multi-GPU) ----------------
// Every rank owns an exchange buffer of ORR_XCHG_SLOTS slots; a slot holds, per source rank, the rank's hit
// list and status as LL words: each 4 bytes of payload travel in one 8-byte {data, seq} store, so the flag
// arrives WITH the data and no fence or separate flag round trip is needed (8-byte stores are single NVLink
// transactions).  One CTA per rank: (1) PUSH the k x 24 B + 8 B of this rank into slot[seq % SLOTS][rank] of
// EVERY peer, all stores in flight at once; (2) PULL: spin on each expected word of the local slot until its
// tag is seq (bounded by a timeout so a lost peer surfaces as a status flag, not a hung GPU) and unpack into
// shared memory; (3) order the union with the reference tie chain, exactly like orr_merge_kernel.
// No NCCL launch, no host round trip, one NVLink store latency.
__device__ __forceinline__ void st_ll(uint2* p, uint32_t data, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) orr_xchg_merge_kernel(const OrrXchgArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int k = min(max(1, a.top_k), a.kmax);
    const int words = k * 6 + 2;                                   // k hits (6 words each) + {n, flags}
    const int words_max = a.kmax * 6 + 2;                          // LL words reserved per source rank
    const size_t slot_off = (size_t)a.slot * a.slot_bytes;
    // smem: gathered hits [world][k] + status [world][2], then the sort array
    uint32_t* g_hits = reinterpret_cast<uint32_t*>(smem_raw);
    int32_t* g_status = reinterpret_cast<int32_t*>(g_hits + (size_t)a.world * k * 6);
    OrrExact* e = reinterpret_cast<OrrExact*>(smem_raw + (((size_t)a.world * (k * 24 + 8)) + 15) / 16 * 16);
    __shared__ int s_timeout;
    if (tid == 0) s_timeout = 0;
    // (1) push: word i of this rank -> every peer
    const uint32_t* src_h = reinterpret_cast<const uint32_t*>(a.src_hits);
    for (int idx = tid; idx < words * a.world; idx += blockDim.x) {
        const int p = idx / words, i = idx - p * words;
        const uint32_t v = i < k * 6 ? src_h[i] : (uint32_t)a.src_status[i - k * 6];
        st_ll(reinterpret_cast<uint2*>(a.peer_base[p] + slot_off) + (size_t)a.rank * words_max + i, v, a.seq);
    }
    __syncthreads();
    // (2) pull: every expected word of the local slot
    const uint2* mine = reinterpret_cast<const uint2*>(a.peer_base[a.rank] + slot_off);
    const unsigned long long t0 = global_timer_ns();
    for (int idx = tid; idx < words * a.world; idx += blockDim.x) {
        const int r = idx / words, i = idx - r * words;
        const uint2* w = mine + (size_t)r * words_max + i;
        uint2 v = ld_ll(w);
        while (v.y != a.seq) {
            if (global_timer_ns() - t0 > a.timeout_ns) { atomicOr(&s_timeout, 1); v.x = 0u; break; }
            __nanosleep(32);
            v = ld_ll(w);
        }
        if (i < k * 6) g_hits[(size_t)r * k * 6 + i] = v.x;
        else g_status[2 * r + (i - k * 6)] = (int32_t)v.x;
    }
    __syncthreads();
    if (s_timeout && tid < a.world) g_status[2 * tid] = 0;        // incomplete exchange: return nothing, flag it
    __syncthreads();
    // (3) merge
    merge_lists_cta(reinterpret_cast<const orr_hit*>(g_hits), g_status, a.world, k, a.top_k, a.out, a.out_status,
                    s_timeout ? ORR_XCHG_FLAG_TIMEOUT : 0, e);
}

}  // namespace

int orr_launch_rescore(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                       const OrrWeights& w, int64_t now_ticks, int q_dim, int top_k,
                       int n_listed_max, bool check_bound, cudaStream_t st, int n_listed_host) {
    if (n_listed_max < 1) n_listed_max = 1;
    if (n_listed_max > ORR_SORT_MAX) { orr_set_error("rescore: %d rows exceed the sorter", n_listed_max); return ORR_E_INTERNAL; }
    RescoreArgs a;
    fill_exact_args(a.ex, sh, sc, pr, w, now_ticks, q_dim);
    a.rows = sc.surv_rows;
    a.n_listed = n_listed_host >= 0 ? nullptr : sc.sel + 0;
    a.n_listed_value = n_listed_host;
    a.tau_bits = sc.sel + 1;
    a.n_listed_max = n_listed_max;
    a.top_k = top_k;
    a.check_bound = check_bound ? 1 : 0;
    a.eps = (double)ORR_SELECT_EPS * (fabs(w.w_cos) + fabs(w.w_kw) + fabs(w.w_rec));
    a.exact = sc.exact;
    a.ticket = sc.sel + 3;
    a.hits = sc.hits;
    a.status = sc.status;
    const int grid = (n_listed_max + 3) / 4;
    orr_rescore_kernel<<<grid, 128, 0, st>>>(a);
    ORR_CUDA_OK(cudaGetLastError());
    OrderArgs o;
    o.exact = a.exact; o.n_ptr = a.n_listed; o.n_value = a.n_listed_value; o.n_max = a.n_listed_max; o.top_k = a.top_k;
    o.check_bound = a.check_bound; o.tau_bits = a.tau_bits; o.eps = a.eps; o.row_base = sh.row_base; o.hits = a.hits; o.status = a.status;
    ORR_CUDA_OK(orr_launch_dependent(orr_order_kernel, (n_listed_max + 7) / 8, 256, st, o));
    return ORR_OK;
}

// orders n (device count, <= n_max) records and writes the first top_k as hits; status is left to the caller
int orr_launch_order(const OrrExact* recs, const int32_t* n_dev, int n_max, int top_k, uint64_t row_base, orr_hit* hits, cudaStream_t st,
                     bool dependent) {
    OrderArgs o;
    o.exact = recs; o.n_ptr = n_dev; o.n_value = 0; o.n_max = n_max; o.top_k = top_k; o.check_bound = 0; o.tau_bits = nullptr; o.eps = 0.0;
    o.row_base = row_base; o.hits = hits; o.status = nullptr;
    if (dependent) ORR_CUDA_OK(orr_launch_dependent(orr_order_kernel, (n_max + 7) / 8, 256, st, o));
    else orr_order_kernel<<<(n_max + 7) / 8, 256, 0, st>>>(o);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_launch_merge(const orr_hit* lists_dev, const int32_t* status_dev, int n_lists, int stride, int top_k,
                     orr_hit* out_dev, int32_t* out_status_dev, cudaStream_t st) {
    const int total = n_lists * stride;
    if (n_lists < 1 || n_lists > 256 || stride < 1 || total > ORR_SORT_MAX) {
        orr_set_error("merge: %d lists x %d exceed the sorter", n_lists, stride);
        return ORR_E_UNSUPPORTED;
    }
    ORR_SMEM_OPT_IN((orr_merge_kernel), ORR_SORT_MAX * (int)sizeof(OrrExact));
    int np2 = 1;
    while (np2 < total) np2 <<= 1;
    orr_merge_kernel<<<1, 256, np2 * sizeof(OrrExact), st>>>(lists_dev, status_dev, n_lists, stride, top_k, out_dev,
                                                            out_status_dev);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_launch_merge_batch(const orr_hit* lists_dev, const int32_t* n_dev, int n_lists, int batch, int k, orr_hit* out_dev,
                           int32_t* n_out_dev, cudaStream_t st) {
    const int total = n_lists * k;
    if (n_lists < 1 || k < 1 || batch < 0 || total > ORR_SORT_MAX) { orr_set_error("batch merge: %d lists x %d exceed the sorter", n_lists, k); return ORR_E_UNSUPPORTED; }
    if (batch == 0) return ORR_OK;
    ORR_SMEM_OPT_IN((orr_merge_batch_kernel), ORR_SORT_MAX * (int)sizeof(OrrExact));
    int np2 = 1;
    while (np2 < total) np2 <<= 1;
    orr_merge_batch_kernel<<<batch, 128, np2 * sizeof(OrrExact), st>>>(lists_dev, n_dev, n_lists, batch, k, out_dev, n_out_dev);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_launch_xchg_merge(const OrrXchgArgs& a, cudaStream_t st) {
    const int total = a.world * a.kmax;
    if (a.world < 1 || a.world > ORR_XCHG_MAX_WORLD || a.kmax < 1 || total > ORR_SORT_MAX) {
        orr_set_error("xchg merge: %d ranks x %d exceed the sorter", a.world, a.kmax);
        return ORR_E_UNSUPPORTED;
    }
    ORR_SMEM_OPT_IN((orr_xchg_merge_kernel), 2 * ORR_SORT_MAX * (int)sizeof(OrrExact) + 1024);
    const int k = std::min(std::max(1, a.top_k), a.kmax);
    int np2 = 1;
    while (np2 < a.world * k) np2 <<= 1;
    const size_t smem = ((size_t)a.world * (k * 24 + 8) + 15) / 16 * 16 + (size_t)np2 * sizeof(OrrExact);
    orr_xchg_merge_kernel<<<1, 256, smem, st>>>(a);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

// =====================================================================================================
// Batched queries: threshold selection from the sampling pass, and the per-query finalize
// (select survivors among the GEMM candidates, exact fp64 re-score, order, bound check).
// =====================================================================================================
namespace {

__device__ OrrBatchProbes g_no_probes;     // zero-initialised: no terms

__device__ __forceinline__ uint32_t fkey(float f) {              // monotone; NaN lowest, 0 reserved
    if (f != f) return 1u;
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// thr[b] = the rstar-th largest of dense[b][0..n) (fp32, or fp16 from the sampling pass): one CTA per query.
//   Small rstar (the sampling pass asks for the ~12th best of ~60 k scores): ONE pass over the scores in which every
//   thread keeps its 4 largest keys in registers, then the rstar-th largest of those 1024.  That is the exact answer
//   unless some thread's 4th-largest key exceeds it (it may then have dropped larger ones); the CTA checks and falls
//   back to the general form.
//   General form: MSB-first radix select, 4 passes over the scores with a shared 256-bin histogram.
template <class T> __device__ __forceinline__ float thr_load(const T* x, int i);
template <> __device__ __forceinline__ float thr_load<float>(const float* x, int i) { return x[i]; }
template <> __device__ __forceinline__ float thr_load<__half>(const __half* x, int i) { return __half2float(x[i]); }

constexpr int THR_KEEP = 4;                    // keys a thread keeps in the one-pass form
constexpr int THR_FAST_RSTAR = 16;

template <class T>
__global__ void __launch_bounds__(256) orr_batch_threshold_kernel(const T* dense, int64_t ld, int n, int rstar,
                                                                  float* thr, int batch) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b >= batch) { if (tid == 0) thr[b] = INFINITY; return; }   // padding queries never produce candidates
    if (n < rstar || rstar < 1) { if (tid == 0) thr[b] = -INFINITY; return; }
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rem;
    __shared__ uint32_t s_keys[256 * THR_KEEP];
    const T* x = dense + (int64_t)b * ld;
    if (rstar <= THR_FAST_RSTAR) {
        uint32_t t[THR_KEEP];                                        // descending; 0 = empty (fkey never returns 0)
#pragma unroll
        for (int j = 0; j < THR_KEEP; ++j) t[j] = 0u;
        for (int i = tid; i < n; i += 256) {
            uint32_t k = fkey(thr_load<T>(x, i));
            if (k > t[THR_KEEP - 1]) {
#pragma unroll
                for (int j = 0; j < THR_KEEP; ++j) {                 // insertion: k ends up holding the evicted key
                    const uint32_t hi = max(t[j], k), lo = min(t[j], k);
                    t[j] = hi; k = lo;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < THR_KEEP; ++j) s_keys[j * 256 + tid] = t[j];
        if (tid == 0) { s_prefix = 0; s_rem = (uint32_t)rstar; }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t pmask = pass ? (0xffffffffu << (shift + 8)) : 0u;
#pragma unroll
            for (int j = 0; j < THR_KEEP; ++j) {
                const uint32_t k = s_keys[j * 256 + tid];
                if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t rem = s_rem;
                int d = 255;
                for (; d > 0; --d) { if (hist[d] >= rem) break; rem -= hist[d]; }
                s_prefix = prefix | ((uint32_t)d << shift);
                s_rem = rem;
            }
            __syncthreads();
        }
        const uint32_t kth = s_prefix;
        // a thread whose 4th key is above the answer may have dropped keys above it as well
        if (!__syncthreads_or(t[THR_KEEP - 1] > kth)) {
            if (tid == 0) thr[b] = fkey_inv(kth);
            return;
        }
    }
    if (tid == 0) { s_prefix = 0; s_rem = (uint32_t)rstar; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = pass ? (0xffffffffu << (shift + 8)) : 0u;
        for (int i = tid; i < n; i += 256) {
            const uint32_t k = fkey(thr_load<T>(x, i));
            if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rem = s_rem;
            int d = 255;
            for (; d > 0; --d) { if (hist[d] >= rem) break; rem -= hist[d]; }
            s_prefix = prefix | ((uint32_t)d << shift);
            s_rem = rem;
        }
        __syncthreads();
    }
    if (tid == 0) thr[b] = fkey_inv(s_prefix);
}

struct BatchExact {                   // the fields exact_row_q reads
    OrrShard sh;
    int32_t q_dim;
    OrrWeights w;
    int64_t now_ticks;
};

struct BatchFinArgs {
    BatchExact ex;
    const float* q;                   // [B][dim]
    const OrrBatchProbes* probes;     // [B] or NULL
    const uint2* cand;                // [B][cap]
    const uint32_t* cand_count;       // [B]
    const float* thr;                 // [B]
    int32_t cap, n_surv, top_k, k_stride;
    double eps;
    const int32_t* qbad;              // [B] 1 = the screen could not rank this query (norm outside fp32): always unproven
    orr_hit* hits;                    // [B][k_stride]
    int32_t* status;                  // [B][2]
};

constexpr int BATCH_FIN_THREADS = 256;

__global__ void __launch_bounds__(BATCH_FIN_THREADS, 3) orr_batch_finalize_kernel(const BatchFinArgs a) {
    extern __shared__ __align__(16) uint8_t fsm[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(fsm);                       // [cap pow2]
    OrrExact* e = reinterpret_cast<OrrExact*>(fsm + (size_t)a.cap * 8);       // [ORR_BATCH_MAX_SURV]
    float* qs = reinterpret_cast<float*>(fsm + (size_t)a.cap * 8 + (size_t)ORR_BATCH_MAX_SURV * sizeof(OrrExact));   // [dim]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t total = a.cand_count[b];
    const int c = (int)min(total, (uint32_t)a.cap);
    int flags = total > (uint32_t)a.cap ? 2 : 0;                             // candidate list overflowed
    if (a.qbad != nullptr && a.qbad[b]) flags |= 1;
    const uint2* src = a.cand + (int64_t)b * a.cap;
    // key = (monotone score key, ~row): unique per candidate, larger = better (ties: lower row first)
    for (int i = tid; i < c; i += BATCH_FIN_THREADS) {
        const uint2 v = src[i];
        keys[i] = ((uint64_t)fkey(__uint_as_float(v.y)) << 32) | (uint64_t)(~v.x);
    }
    // The best M candidates by an MSB-first radix select over the 64-bit keys (8 passes of 256 bins, ~1k
    // instructions per thread; the full bitonic sort it replaces cost ~7k and its order was never used: the
    // survivors are re-ordered by their exact scores below).  kth = the M-th largest key; keys are unique, so
    // exactly M candidates have key >= kth.
    __shared__ uint32_t hist[BATCH_FIN_THREADS];
    __shared__ unsigned long long s_prefix, s_below[BATCH_FIN_THREADS / 32];
    __shared__ uint32_t s_rem, s_cnt;
    const int M = a.n_surv;
    const int ns = min(M, c);
    unsigned long long kth = 0ull;
    if (c > M) {
        unsigned long long prefix = 0ull;
        uint32_t rem = (uint32_t)M;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            hist[tid] = 0u;
            __syncthreads();
            for (int i = tid; i < c; i += BATCH_FIN_THREADS) {
                const unsigned long long k = keys[i];
                if (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(uint32_t)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {                                                 // lane l owns bins 8l .. 8l+7 (higher = better)
                uint32_t h[8], sum = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; sum += h[j]; }
                uint32_t suffix = sum;                                       // bins of this lane and every higher lane
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_down_sync(FULL, suffix, o);
                    if (lane + o < 32) suffix += v;
                }
                const uint32_t above = suffix - sum;
                if (above < rem && rem <= suffix) {                          // the rem-th best key falls into this lane's bins
                    uint32_t r = rem - above;
                    int d = 0;
#pragma unroll
                    for (int j = 7; j >= 0; --j) {
                        if (h[j] >= r) { d = j; break; }
                        r -= h[j];
                    }
                    s_prefix = prefix | ((unsigned long long)(8 * lane + d) << shift);
                    s_rem = r;
                }
            }
            __syncthreads();
            prefix = s_prefix;
            rem = s_rem;
        }
        kth = prefix;
    }
    if (tid == 0) s_cnt = 0u;
    __syncthreads();
    unsigned long long below = 0ull;                                         // best key that does not survive
    for (int i = tid; i < c; i += BATCH_FIN_THREADS) {
        const unsigned long long k = keys[i];
        if (k >= kth) e[atomicAdd(&s_cnt, 1u)].row = (uint64_t)(~(uint32_t)k);
        else below = max(below, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below = max(below, __shfl_xor_sync(FULL, below, o));
    if (lane == 0) s_below[warp] = below;
    __syncthreads();
    float tau = a.thr[b];                                                    // no row outside the list beats it
    if (c > M) {
        unsigned long long bk = 0ull;
#pragma unroll
        for (int w = 0; w < BATCH_FIN_THREADS / 32; ++w) bk = max(bk, s_below[w]);
        tau = fmaxf(tau, fkey_inv((uint32_t)(bk >> 32)));
    }
    // exact re-score of the survivors
    const float* q = a.q + (int64_t)b * a.ex.sh.dim;
    const OrrBatchProbes& pr = a.probes ? a.probes[b] : g_no_probes;
    const bool has_q = (a.ex.q_dim == a.ex.sh.dim && a.ex.q_dim > 0);
    const double nA = has_q ? exact_qnorm_q(a.ex, q, lane) : 0.0;
    if (has_q) for (int i = tid; i < a.ex.sh.dim; i += BATCH_FIN_THREADS) qs[i] = q[i];   // the query is re-read for every row
    __syncthreads();
    // a warp takes survivors warp, warp + 8, ...; the warp-collective part (dot products, keyword matches) runs row
    // by row, lane j keeps the partial result of the group's j-th row, and the scalar fp64 tail runs once per lane
    for (int g = warp; g < ns; g += BATCH_FIN_THREADS) {
        ExactPartial mine;
        mine.dot = 0.0; mine.nB = 0.0; mine.matches = 0; mine.kw_den = -1; mine.ticks = 0;
        for (int j = 0; j < 32; ++j) {
            const int i = g + j * (BATCH_FIN_THREADS / 32);
            if (i >= ns) break;                                              // warp-uniform
            const ExactPartial r = exact_row_partial<BatchExact, 6, true, OrrBatchProbes>(a.ex, qs, pr, (int64_t)e[i].row, lane);
            if (lane == j) mine = r;
        }
        const int i = g + lane * (BATCH_FIN_THREADS / 32);
        if (i < ns) { e[i].score = exact_row_finish(a.ex, nA, mine); e[i].ticks = mine.ticks; }
    }
    __syncthreads();
    __shared__ double s_kth;
    if (tid == 0) s_kth = __longlong_as_double(0x7ff8000000000000LL);
    __syncthreads();
    const int k = max(1, a.top_k);
    const int n_out = min(k, ns);
    if (ns <= 128) {
        // few survivors: positions by counting (no barrier stages)
        rank_emit(e, ns, n_out, k, a.ex.sh.row_base, a.hits + (int64_t)b * a.k_stride, &s_kth);
    } else {
        // the bitonic network: n log^2 n / 4 compare-exchanges against the n^2 comparisons of counting (11.5k vs 262k at
        // 512 survivors); with 3 CTAs per SM its barrier latency is hidden
        int ep2 = 1;
        while (ep2 < ns) ep2 <<= 1;
        for (int i = ns + tid; i < ep2; i += BATCH_FIN_THREADS) {
            e[i].score = __longlong_as_double(0x7ff8000000000000LL); e[i].ticks = INT64_MIN; e[i].row = ~0ull;
        }
        __syncthreads();
        for (int k2 = 2; k2 <= ep2; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < ep2; i += BATCH_FIN_THREADS) {
                    const int p = i ^ j;
                    if (p > i) {
                        const OrrExact x = e[i], y = e[p];
                        const bool up = ((i & k2) == 0);
                        if (up ? ranks_before(y, x) : ranks_before(x, y)) { e[i] = y; e[p] = x; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < n_out; i += BATCH_FIN_THREADS) {
            orr_hit h; h.row = a.ex.sh.row_base + e[i].row; h.score = e[i].score; h.created_ticks = e[i].ticks;
            a.hits[(int64_t)b * a.k_stride + i] = h;
        }
        if (tid == 0 && ns >= k) s_kth = e[k - 1].score;
    }
    __syncthreads();
    if (tid == 0) {
        if (tau != -INFINITY) {
            const double sk = s_kth;                                         // NaN when fewer than k survivors
            if (!(sk - a.eps > (double)tau)) flags |= 1;                     // selection not provably safe
        }
        a.status[2 * b] = n_out;
        a.status[2 * b + 1] = flags;
    }
}

// term bitmaps: bit r of bits[t] says row r holds batch term t.  A warp takes 32 consecutive rows (one
// output word per term): the rows' 32-bit term tables are one contiguous block, streamed with 8
// 16-byte loads in flight per lane.
//   1. Every stored hash tests one bit of a 2^18-bit filter of the batch's new terms (smem, branch-free:
//      multiply, shift, LDS, shift) — a few percent pass.
//   2. The survivors of a 256-vector step are COMPACTED into a per-warp smem queue (per-lane pop-count, one warp
//      scan, predicated stores: no divergent code) — with ~2 % passing, nearly every warp had one or two lanes in
//      the probe loop and the other lanes idle (ncu: 141 instructions per warp and vector).
//   3. The queue is drained with all lanes busy: probe the open-addressing table (smem), OR the row's bit into the
//      warp's accumulator.  A step with more survivors than the queue holds (a very frequent new term) takes the
//      per-lane path instead.
constexpr int TERM_BITS_THREADS = 1024;
constexpr int TERM_ACC = 64;                  // per-warp accumulator entries (direct-mapped by slot)
constexpr int TERM_FILTER_LOG2 = 18;          // filter bits (32 KB of smem behind the probe table)
constexpr int TERM_FILTER_WORDS = (1 << TERM_FILTER_LOG2) / 32;
constexpr int TERM_QUEUE = 128;               // per-warp queue of {hash, row bit} that passed the filter
__global__ void __launch_bounds__(TERM_BITS_THREADS) orr_batch_term_bits_kernel(const uint32_t* terms32, int slots, int64_t rows,
                                                                               const uint2* table, int table_mask,
                                                                               uint32_t* bits, int64_t slot_cap) {
    extern __shared__ uint2 tab[];
    uint32_t* filt = reinterpret_cast<uint32_t*>(tab + table_mask + 1);
    // Zipf vocabularies: a handful of terms hit in most rows, i.e. up to 32 times per output word.  Each warp owns
    // its 32-row block's words, so it first ORs hits into a small shared accumulator ({slot + 1, bits}, claimed by
    // CAS) and flushes one atomic per (term, block); only accumulator conflicts go to global memory directly.
    __shared__ uint2 acc_all[TERM_BITS_THREADS / 32][TERM_ACC];
    __shared__ uint2 queue_all[TERM_BITS_THREADS / 32][TERM_QUEUE];
    for (int i = threadIdx.x; i < TERM_FILTER_WORDS; i += blockDim.x) filt[i] = 0u;
    for (int i = threadIdx.x; i < (TERM_BITS_THREADS / 32) * TERM_ACC; i += blockDim.x) (&acc_all[0][0])[i] = make_uint2(0u, 0u);
    __syncthreads();
    for (int i = threadIdx.x; i <= table_mask; i += blockDim.x) {
        const uint2 ent = table[i];
        tab[i] = ent;
        if (ent.x) {
            const uint32_t f = (ent.x * 0x9E3779B1u) >> (32 - TERM_FILTER_LOG2);
            atomicOr(filt + (f >> 5), 1u << (f & 31u));
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t* acc = reinterpret_cast<uint32_t*>(acc_all[threadIdx.x >> 5]);
    uint2* queue = queue_all[threadIdx.x >> 5];
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_blocks = (rows + 31) >> 5;
    const int vec_shift = 31 - __clz(slots >> 2);                            // uint4 per row = 8, 16 or 32

    // probe the table for hash h; a hit ORs `bit` into the block's word of that term
    auto probe = [&](uint32_t h, uint32_t bit, int64_t blk) {
        uint32_t pos = (h * 0x9E3779B1u) & (uint32_t)table_mask;
        for (;;) {
            const uint2 ent = tab[pos];
            if (ent.x == 0u) break;
            if (ent.x == h) {
                uint32_t* e = acc + 2 * (ent.y & (TERM_ACC - 1));
                const uint32_t cur = atomicCAS(e, 0u, ent.y + 1u);
                if (cur == 0u || cur == ent.y + 1u) atomicOr(e + 1, bit);
                else atomicOr(bits + (((blk >> 3) * slot_cap + ent.y) << 3) + (blk & 7), bit);
                break;
            }
            pos = (pos + 1) & (uint32_t)table_mask;
        }
    };

    for (int64_t blk = gw; blk < n_blocks; blk += W) {
        const int rows_here = (int)min((int64_t)32, rows - (blk << 5));
        const int n_vec = rows_here << vec_shift;
        const uint4* base = reinterpret_cast<const uint4*>(terms32 + (blk << 5) * slots);
        for (int v0 = 0; v0 < n_vec; v0 += 256) {
            uint4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int v = v0 + i * 32 + lane;
                x[i] = v < n_vec ? __ldg(base + v) : make_uint4(0u, 0u, 0u, 0u);
            }
            uint32_t pass = 0u;                                              // bit 4i + c: component c of x[i] passed the filter
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t hs[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t f = (hs[c] * 0x9E3779B1u) >> (32 - TERM_FILTER_LOG2);
                    pass |= ((hs[c] != 0u ? filt[f >> 5] >> (f & 31u) : 0u) & 1u) << (4 * i + c);
                }
            }
            // warp scan of the per-lane survivor counts
            const uint32_t mine = (uint32_t)__popc(pass);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            if (total == 0u) continue;
            if (total <= (uint32_t)TERM_QUEUE) {
                uint32_t at = incl - mine;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t hs[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
                    const uint32_t bit = 1u << ((v0 + i * 32 + lane) >> vec_shift);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if ((pass >> (4 * i + c)) & 1u) { queue[at] = make_uint2(hs[c], bit); ++at; }
                    }
                }
                __syncwarp();
                for (uint32_t idx = lane; idx < total; idx += 32) {
                    const uint2 qe = queue[idx];
                    probe(qe.x, qe.y, blk);
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (((pass >> (4 * i)) & 15u) == 0u) continue;
                    const uint32_t hs[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
                    const uint32_t bit = 1u << ((v0 + i * 32 + lane) >> vec_shift);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if ((pass >> (4 * i + c)) & 1u) probe(hs[c], bit, blk);
                }
            }
        }
        __syncwarp();
        for (int j = lane; j < TERM_ACC; j += 32) {
            const uint32_t sl = acc[2 * j];
            if (sl) {
                atomicOr(bits + (((blk >> 3) * slot_cap + (sl - 1u)) << 3) + (blk & 7), acc[2 * j + 1]);
                acc[2 * j] = 0u; acc[2 * j + 1] = 0u;
            }
        }
        __syncwarp();
    }
}

// clears slots [first, first + n) of every row tile (n x 32 contiguous bytes per tile)
__global__ void orr_batch_clear_slots_kernel(uint4* bits, int64_t slot_cap, int first, int n, int64_t n_tiles) {
    const int64_t total = n_tiles * n * 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = i / (2 * n), r = i - tile * 2 * n;
        bits[(tile * slot_cap + first) * 2 + r] = make_uint4(0u, 0u, 0u, 0u);
    }
}

}  // namespace

int orr_batch_launch_threshold(const void* dense, int dense_half, int64_t ld, int n, int rstar, float* thr, int batch,
                               int batch_padded, cudaStream_t st) {
    if (dense_half) orr_batch_threshold_kernel<__half><<<batch_padded, 256, 0, st>>>((const __half*)dense, ld, n, rstar, thr, batch);
    else orr_batch_threshold_kernel<float><<<batch_padded, 256, 0, st>>>((const float*)dense, ld, n, rstar, thr, batch);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_launch_finalize(const OrrShard& sh, const float* q, int q_dim, const OrrBatchProbes* probes, const OrrWeights& w,
                              int64_t now_ticks, const void* cand, const uint32_t* cand_count, const float* thr, int cap,
                              int n_surv, int top_k, int k_stride, double eps, const int32_t* qbad, orr_hit* hits, int32_t* status,
                              int batch, cudaStream_t st) {
    if (n_surv > ORR_BATCH_MAX_SURV || (cap & (cap - 1)) != 0) { orr_set_error("batch finalize: bad sizes"); return ORR_E_INTERNAL; }
    const int smem = cap * 8 + ORR_BATCH_MAX_SURV * (int)sizeof(OrrExact) + sh.dim * (int)sizeof(float);
    ORR_SMEM_OPT_IN((orr_batch_finalize_kernel), 8192 * 8 + ORR_BATCH_MAX_SURV * 24 + 8192 * 4);
    BatchFinArgs a;
    a.ex.sh = sh; a.ex.q_dim = q_dim; a.ex.w = w; a.ex.now_ticks = now_ticks;
    a.q = q; a.probes = probes; a.cand = (const uint2*)cand; a.cand_count = cand_count; a.thr = thr;
    a.cap = cap; a.n_surv = n_surv; a.top_k = top_k; a.k_stride = k_stride; a.eps = eps; a.qbad = qbad; a.hits = hits; a.status = status;
    orr_batch_finalize_kernel<<<batch, BATCH_FIN_THREADS, smem, st>>>(a);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

int orr_batch_launch_term_bits(const uint32_t* terms32, int slots, int64_t rows, const void* table, int table_slots,
                               uint32_t* bits, int64_t slot_cap, int first_new, int n_new, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (rows + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE;
    if (n_new > 0 && n_tiles > 0) {
        orr_batch_clear_slots_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<uint4*>(bits), slot_cap, first_new, n_new, n_tiles);
        ORR_CUDA_OK(cudaGetLastError());
    }
    const int smem = table_slots * 8 + TERM_FILTER_WORDS * 4;             // probe table + filter (+ 48 KB static: accumulators, queues)
    ORR_SMEM_OPT_IN((orr_batch_term_bits_kernel), 160 * 1024);
    const int per_sm = smem <= 64 * 1024 ? 2 : 1;
    orr_batch_term_bits_kernel<<<sms * per_sm, TERM_BITS_THREADS, smem, st>>>(terms32, slots, rows, (const uint2*)table,
                                                                              table_slots - 1, bits, slot_cap);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}
