#include <utility>
// orr_internal.h — shared declarations of liborr.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/orr.h"

// ---- error plumbing -----------------------------------------------------------------------
void orr_set_error(const char* fmt, ...);
#define ORR_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            orr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                        \
            return _e == cudaErrorMemoryAllocation ? ORR_E_OOM : ORR_E_CUDA;                \
        }                                                                                   \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs
// (orr_cluster) must opt in on each of them.  One bit per device ordinal and call site.
#define ORR_SMEM_OPT_IN(func, bytes)                                                                        \
    do {                                                                                                    \
        static std::atomic<uint64_t> _done{0};                                                              \
        int _dev = 0;                                                                                       \
        cudaGetDevice(&_dev);                                                                               \
        if (!((_done.load(std::memory_order_relaxed) >> (_dev & 63)) & 1ull)) {                             \
            ORR_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _done.fetch_or(1ull << (_dev & 63));                                                            \
        }                                                                                                   \
    } while (0)

// ---- constants ----------------------------------------------------------------------------
constexpr int64_t ORR_DEAD_TICKS = INT64_MIN;   // tombstone marker in the ticks column
constexpr int ORR_SCAN_WARPS = 8;               // consumer warps per scan CTA
constexpr int ORR_SCAN_STAGES = 2;              // TMA stages per warp
constexpr int ORR_WARP_LIST = 32;               // per-warp register top list (one entry per lane)
constexpr int ORR_MAX_SURVIVORS = 256;          // rows re-scored exactly on the fused path
constexpr int ORR_FUSED_MAX_K = 224;            // largest top_k the fused path serves
constexpr int ORR_SORT_MAX = 4096;              // entries the single-CTA exact sorter handles
constexpr float ORR_SELECT_EPS = 2.0e-5f;       // bound on |fp32 scan score - exact score| (unit weights)

// A query's keyword side as the kernels see it (passed by value in kernel params).
struct OrrProbes {
    int32_t  n_terms;                            // denominator of KeywordScore (:112)
    int32_t  n_probes;
    uint32_t h32[ORR_MAX_QUERY_PROBES];          // low word of the 64-bit hash (scan)
    uint64_t h64[ORR_MAX_QUERY_PROBES];          // full hash (exact re-score)
    uint8_t  term[ORR_MAX_QUERY_PROBES];         // query term index each probe satisfies
};

// the batched path's probes: identity (one 64-bit hash per term), <= ORR_BATCH_TERMS terms
struct OrrBatchProbes {
    int32_t  n_terms;
    int32_t  reserved;
    uint64_t h64[16];
};
#if defined(__CUDACC__)
__device__ __forceinline__ int orr_probe_count(const OrrProbes& p) { return p.n_probes; }
__device__ __forceinline__ uint32_t orr_probe_term(const OrrProbes& p, int i) { return p.term[i]; }
__device__ __forceinline__ int orr_probe_count(const OrrBatchProbes& p) { return p.n_terms; }
__device__ __forceinline__ uint32_t orr_probe_term(const OrrBatchProbes&, int i) { return (uint32_t)i; }
#endif

struct OrrWeights {
    double w_cos, w_kw, w_rec, recency_days;
};

// exact per-row record produced by the re-score kernels and ordered by the sorter
struct OrrExact {
    double   score;
    int64_t  ticks;
    uint64_t row;      // LOCAL row index
};

// per-search device scratch (one per concurrent search; see SearchCtx in orr_api.cu)
struct OrrScratch {
    float*    q;              // [dim] query embedding on device
    uint2*    cta_cands;      // [grid][ORR_MAX_SURVIVORS] (ordered-key, row) per CTA
    float*    cta_floor;      // [grid] best score a CTA discarded (-inf if none)
    uint32_t* surv_rows;      // [ORR_SORT_MAX] rows to re-score exactly
    OrrExact* exact;          // [ORR_SORT_MAX] exact records
    int32_t*  sel;            // [8]: {n_surv, tau_bits, ticket_scan, ticket_rescore, ...}
    orr_hit*  hits;           // [max k] results
    int32_t*  status;         // [2] {n_out, flags}
    // exact path (orr_exact.cu), lazily allocated
    uint64_t* skey;           // [capacity] order-preserving key of every row's exact score
    void*     sel_state;      // digit-walk state (histograms, prefix points, counters)
    OrrExact* big;            // top_k > ORR_SORT_MAX: the k selected rows (power-of-two capacity)
    int64_t   big_cap;
    // text mode: per-term row bitmaps of the current query (NULL outside orr_search_text)
    const uint32_t* kw_bits; int64_t kw_row_words; int32_t kw_terms;
    float*    scan_dense;     // diagnostic: every row's fp32 scan score (orr_debug_scan_scores), else NULL
};

struct OrrShard {              // device view of one shard
    const float*    emb;       // [rows][dim]
    const int64_t*  ticks;     // [rows]  ORR_DEAD_TICKS = tombstone
    const uint32_t* terms32;   // [rows][slots] low words, 0 = empty slot
    const uint64_t* terms64;   // [rows][slots] full hashes, 0 = empty slot
    int64_t rows;              // rows in use (live + dead)
    int32_t dim;
    int32_t slots;
    uint64_t row_base;
};

// ---- launchers (each returns ORR_OK or sets the error) -----------------------------------
int orr_scan_smem_bytes(int dim, int* warps_out, int* tile_rows_out);

// K1: fused fp32 scan + per-warp register top list + per-CTA merge; the last CTA to finish
// selects the global survivors.
int orr_launch_scan(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                    const OrrWeights& w, int64_t now_ticks, int n_survivors, int grid,
                    cudaStream_t st);

// K3: exact fp64 re-score of the listed rows (warp per row); the last CTA orders them by the
// reference tie chain, runs the selection bound check and emits the top-k hits.
int orr_launch_rescore(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                       const OrrWeights& w, int64_t now_ticks, int q_dim, int top_k,
                       int n_listed_max, bool check_bound, cudaStream_t st, int n_listed_host = -1);

// warp-per-record ordering of <= n_max records (orr_rescore.cu); n read from the device
int orr_launch_order(const OrrExact* recs, const int32_t* n_dev, int n_max, int top_k, uint64_t row_base, orr_hit* hits, cudaStream_t st,
                     bool dependent = false);

// Launches `kernel` as a PROGRAMMATIC DEPENDENT of the kernel before it on `st`: its CTAs may become resident while the
// predecessor still runs and must execute griddepcontrol.wait before touching anything the predecessor writes; the
// predecessor calls griddepcontrol.launch_dependents to allow it.  Hides the 2-3 us launch gap between the short kernels
// of a latency-bound chain (K3's re-score -> order, the exact path's digit passes -> gather -> order).
template <class... KArgs, class... Args>
inline cudaError_t orr_launch_dependent(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define ORR_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define ORR_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")

// exact path (orr_exact.cu): every row's fp64 score as an order-preserving key, then an MSB-first radix select of the
// top-k under (score desc / NaN last, ticks desc, row asc).  orr_launch_exact_select runs digit passes
// [first_pass, first_pass + n_passes) and the gather; status[1] & ORR_EXACT_INCOMPLETE asks for the next passes.
constexpr int32_t ORR_EXACT_INCOMPLETE = 8;
constexpr int32_t ORR_EXACT_INTERNAL = 32;
constexpr int32_t ORR_EXACT_UNPROVEN = 64;      // the 32-bit screen's candidate set could not be proven complete: use the 64-bit kernel
constexpr int ORR_EXACT_PASSES = 15;
size_t orr_exact_state_bytes();
bool orr_exact_keys_are_screened(const OrrScratch& sc, const OrrProbes& pr, int q_dim, bool force_general);
int orr_launch_exact_scores(const OrrShard& sh, const OrrScratch& sc, const OrrProbes& pr,
                            const OrrWeights& w, int64_t now_ticks, int q_dim, cudaStream_t st, bool force_general = false);
int orr_launch_exact_select(const OrrShard& sh, const OrrScratch& sc, int top_k, int first_pass, int n_passes, cudaStream_t st,
                            const OrrProbes* verify_pr = nullptr, const OrrWeights* verify_w = nullptr, int64_t now_ticks = 0);

int orr_launch_merge(const orr_hit* lists_dev, const int32_t* status_dev, int n_lists, int stride, int top_k,
                     orr_hit* out_dev, int32_t* out_status_dev, cudaStream_t st);

// ---- text mode (orr_textmatch.cu): exact substring keyword matching over the chunk text kept in HBM ----
constexpr int ORR_TEXT_MAX_TERM_BYTES = 256;     // longest query term the substring kernel takes
constexpr int ORR_TEXT_TERMS_BYTES = 4096;       // all terms of one query
struct OrrTextTerms {                            // device-side description of a query's terms
    int32_t  n_terms;
    int32_t  max_len;
    uint16_t off[ORR_MAX_QUERY_TERMS + 1];
    uint8_t  bytes[ORR_TEXT_TERMS_BYTES];
};
struct OrrTextView {
    const uint8_t*  text;      // byte arena
    const uint64_t* off;       // [rows] start of each row's lower-cased UTF-8 content
    const uint32_t* len;       // [rows] byte length
};
// bits[t][row >> 5] bit (row & 31) = row's text contains term t.  rows_list == NULL: every row in [0, rows);
// otherwise only the n_list listed rows (the words must have been cleared by the caller).
int orr_launch_text_bits(const OrrTextView& tv, int64_t rows, const OrrTextTerms* terms_dev, int n_terms, int max_len,
                         const uint32_t* rows_list, int n_list, uint32_t* bits, int64_t row_words, cudaStream_t st);

// fused all-gather + merge over peer memory (orr_xchg.cu owns the buffers)
constexpr int ORR_XCHG_MAX_WORLD = 16;
constexpr int ORR_XCHG_SLOTS = 4;
constexpr int ORR_XCHG_FLAG_TIMEOUT = 4;         // OR-ed into the merged status flags when a peer never arrived
struct OrrXchgArgs {
    const orr_hit* src_hits; const int32_t* src_status;     // this rank's local result (device)
    uint8_t* peer_base[ORR_XCHG_MAX_WORLD];                  // every rank's exchange buffer; [rank] is the local one
    size_t   slot_bytes;
    int32_t  world, rank, kmax, top_k, slot;
    uint32_t seq;
    unsigned long long timeout_ns;
    orr_hit* out; int32_t* out_status;
};
int orr_launch_xchg_merge(const OrrXchgArgs& a, cudaStream_t st);

int orr_launch_merge_batch(const orr_hit* lists_dev, const int32_t* n_dev, int n_lists, int batch, int k, orr_hit* out_dev,
                           int32_t* n_out_dev, cudaStream_t st);

int orr_launch_synth_fill(float* emb, int64_t* ticks, uint32_t* terms32, uint64_t* terms64,
                          int dim, int slots, const orr_synth_spec& spec, uint64_t first_row,
                          int64_t local_first, int64_t n, uint8_t* text, uint64_t* text_off, uint32_t* text_len,
                          uint64_t text_base, cudaStream_t st);

// ---- batched path (orr_batch.cu) ------------------------------------------------------------
constexpr int ORR_BATCH_TERMS = 16;          // query terms the batched epilogue handles per query
constexpr int ORR_BATCH_TILE = 256;          // queries per unit == corpus rows per unit (one CTA pair)
constexpr int ORR_BATCH_MAX_QUERIES = 1024;  // queries one GEMM launch takes (their constants and term slots sit in smem)
constexpr int ORR_BATCH_MAX_TERM_SLOTS = 65535;  // term bitmap slots are addressed with 16 bits in the kernel
struct OrrBatchGemm {
    const void* qhi; const void* qmid;       // bf16 [batch_padded][dim]
    const void* ehi; const void* emid;       // bf16 [rows][dim]
    const void* rowaux;                      // float [rows padded to ORR_BATCH_TILE]: w_rec * recency, -inf = dead/padding
    const float* qscale;                     // [batch_padded]
    const float* thr;                        // [batch_padded] (main pass)
    void* cand; uint32_t* cand_count; int32_t cand_cap;
    float* dense; int64_t dense_ld;          // dense score output (sampling / debug) or NULL
    int32_t dense_half;                      // 1: the dense scores are written as fp16 (the sampling pass: they only feed the threshold estimate)
    const uint32_t* term_bits; int64_t slot_cap; const int32_t* q_term_ids; const float* q_kw_w;   // tile-major bitmaps
    int64_t rows; int32_t dim; int32_t batch_padded; int32_t tile_stride; int32_t sms;
    int32_t passes;                          // 3 = split precision (default), 1 = bf16 screen
};
bool orr_batch_planes_tiled();
int64_t orr_batch_plane_elems(int64_t capacity_rows, int dim);               // bf16 elements of one row plane (tile-padded)
int orr_batch_build_planes(const float* emb, void* hi, void* mid, int64_t first, int64_t n, int dim, float w_cos,
                           cudaStream_t st);
int orr_batch_prep_queries(const float* q_dev, void* qhi, void* qmid, float* qscale, int32_t* qbad, int batch, int batch_padded,
                           int dim, cudaStream_t st);
int orr_batch_build_rowrec(const int64_t* ticks, float* rowrec, int64_t rows, int64_t rows_padded, int64_t now_ticks,
                           const OrrWeights& w, cudaStream_t st);
int orr_batch_launch_gemm(const OrrBatchGemm& g, cudaStream_t st);
int orr_batch_launch_threshold(const void* dense, int dense_half, int64_t ld, int n, int rstar, float* thr, int batch,
                               int batch_padded, cudaStream_t st);
int orr_batch_launch_finalize(const OrrShard& sh, const float* q, int q_dim, const OrrBatchProbes* probes, const OrrWeights& w,
                              int64_t now_ticks, const void* cand, const uint32_t* cand_count, const float* thr, int cap,
                              int n_surv, int top_k, int k_stride, double eps, const int32_t* qbad, orr_hit* hits, int32_t* status,
                              int batch, cudaStream_t st);
// bits is tile-major [row tile][slot_cap][8 words]; slots [first_new, first_new + n_new) are cleared first
int orr_batch_launch_term_bits(const uint32_t* terms32, int slots, int64_t rows, const void* table, int table_slots,
                               uint32_t* bits, int64_t slot_cap, int first_new, int n_new, cudaStream_t st);
constexpr int ORR_BATCH_MAX_SURV = 1024;      // deepest per-query survivor list the finalize kernel re-scores
constexpr int ORR_BATCH_SAMPLE_HITS = 12;     // sampled rows expected above a query's threshold
constexpr float ORR_BATCH_EPS = 2.0e-4f;     // bound on |bf16x3 GEMM score - exact score| (unit weights)

// text
uint64_t orr_hash_bytes(const char* s, int64_t n);
std::vector<std::string> orr_distinct_lower_tokens(const char* s, int64_t n);   // white-space split, lower-case, distinct (A-2)
bool orr_is_stop_word(const std::string& t);
std::string orr_lower_invariant(const char* s, int64_t n);

// live vocabulary + substring expansion of query terms (orr_vocab.cu); the caller serialises access
struct OrrVocab;
OrrVocab* orr_vocab_new(int device);
void orr_vocab_free(OrrVocab* v);
uint32_t orr_vocab_add(OrrVocab* v, const char* word, size_t len, uint32_t count);   // returns the word's id; count = chunks holding it
uint32_t orr_vocab_find(const OrrVocab* v, const char* word, size_t len);             // 0xffffffff if absent; read-only
void orr_vocab_addref(OrrVocab* v, uint32_t id, uint32_t count);
void orr_vocab_release(OrrVocab* v, uint32_t id);
void orr_vocab_clear(OrrVocab* v);
uint64_t orr_vocab_live_words(const OrrVocab* v);
uint64_t orr_vocab_words(const OrrVocab* v);
void orr_vocab_serialize(const OrrVocab* v, std::vector<uint8_t>* out);
int orr_vocab_deserialize(OrrVocab* v, const uint8_t* p, size_t len);
int orr_vocab_expand(OrrVocab* v, const std::vector<std::string>& terms, uint64_t* probe_hash, int32_t* probe_term, int32_t cap,
                     int32_t* n_probes);
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline uint32_t orr_hash_low(uint64_t h) {
    uint32_t l = (uint32_t)h;
    return l ? l : 0x9E3779B9u;
}
