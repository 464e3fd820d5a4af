// orr_api.cu — the C ABI (include/orr.h): HBM-resident store + search orchestration.
//
// Store = the chunk side of InMemoryIngestionStore
// (src/OmniRecall.Api/Services/InMemoryIngestionStore.cs:8-9,17-25,50-55) laid out for the
// scan: row-major fp32 embeddings, an int64 CreatedAtUtc column (INT64_MIN = tombstone) and
// two fixed-width hashed term tables (32-bit for the scan, 64-bit for the exact re-score).
// Rows are append-only; replace-by-document tombstones the old rows and appends.
//
// Search = RecallSearchService.SearchAsync's scoring loop and ordering (:26-37):
//   fused   K1 scan (fp32 select, orr_scan.cu) -> K3 exact re-score/order/bound check
//   exact   full fp64 pass + stable two-key radix sort (no-embedding mode, large k, or when
//           K3's bound check could not prove the fp32 selection safe)
//   subset  candidate_cap > 0: the reference's "300 most recent chunks" pre-selection
//           (:26, InMemoryIngestionStore.cs:57-65) on the host tick mirror, exact scoring of
//           just those rows.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "orr_internal.h"

namespace {

struct SearchCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    OrrScratch sc{};
    float* h_q = nullptr;          // pinned
    orr_hit* h_hits = nullptr;     // pinned
    int32_t* h_status = nullptr;   // pinned [2]
    uint32_t* h_rows = nullptr;    // pinned [ORR_SORT_MAX]
    int32_t hits_cap = 0;
    int64_t exact_rows_cap = 0;
    // text mode scratch
    OrrTextTerms* d_terms = nullptr; OrrTextTerms* h_terms = nullptr;   // device / pinned
    uint32_t* d_kw_bits = nullptr; size_t kw_bits_words = 0;
    // orr_search_device leases: the scratch is busy until `done` (recorded on the caller's stream) has completed
    cudaEvent_t done = nullptr;
    bool busy = false;
    // host-API contexts: the kernels write hits and status STRAIGHT into the pinned host buffers (sc.hits == h_hits,
    // sc.status == h_status; pinned memory is device-accessible under UVA), so a call has no D2H copies at all
    bool zero_copy = false;
    // subset path: the capped row list last uploaded to sc.surv_rows
    uint32_t* d_cap_rows = nullptr;            // [ORR_SORT_MAX], its own buffer: the fused path overwrites sc.surv_rows
    uint64_t cap_version = ~0ull; int32_t cap_value = -1; int32_t cap_n = 0;
};

thread_local orr_timing g_timing{};
thread_local int32_t g_batch_terms_built = 0;

// 32-bit term hash (never 0) -> bitmap slot: flat open-addressing table, ~10 ns per query term on the host
// (a node-based map cost 0.6 ms per 4096-term batch, all of it inside the timed step)
struct FlatTermMap {
    std::vector<uint32_t> keys; std::vector<int32_t> vals; uint32_t mask = 0;
    void reset(size_t entries) {
        size_t n = 1024;
        while (n < 4 * entries) n <<= 1;
        keys.assign(n, 0u); vals.assign(n, -1); mask = (uint32_t)(n - 1);
    }
    void clear() { std::fill(keys.begin(), keys.end(), 0u); }
    int32_t get(uint32_t h) const {
        for (uint32_t pos = (h * 0x9E3779B1u) & mask;; pos = (pos + 1) & mask) {
            if (keys[pos] == h) return vals[pos];
            if (keys[pos] == 0u) return -1;
        }
    }
    void put(uint32_t h, int32_t v) {
        uint32_t pos = (h * 0x9E3779B1u) & mask;
        while (keys[pos] != 0u && keys[pos] != h) pos = (pos + 1) & mask;
        keys[pos] = h; vals[pos] = v;
    }
};

// state of the batched (tcgen05) path: split planes of the store + per-call scratch
struct BatchState {
    std::mutex mu;                       // batched searches are serialised per store
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* ehi = nullptr; void* emid = nullptr;     // bf16 [capacity][dim]
    float* rowaux = nullptr;                       // [capacity padded] w_rec * recency of the current batch's clock
    int64_t planes_rows = 0;                       // rows whose hi plane is built
    int64_t mid_rows = 0;                          // rows whose mid plane is built (allocated on the first bf16x3 pass)
    int bcap = 0, kcap = 0;
    float* q = nullptr; void* qhi = nullptr; void* qmid = nullptr;
    float* qscale = nullptr; float* thr = nullptr; float* kww = nullptr; int32_t* qterm = nullptr; int32_t* qbad = nullptr;
    void* cand = nullptr; uint32_t* cand_count = nullptr; orr_hit* hits = nullptr; int32_t* status = nullptr;
    OrrBatchProbes* probes = nullptr;
    // pinned staging of the per-call host arrays (one batched search at a time per store)
    int32_t* h_qterm = nullptr; float* h_kww = nullptr; OrrBatchProbes* h_probes = nullptr; uint2* h_table = nullptr;
    int32_t* h_status = nullptr;
    float* dense = nullptr; size_t dense_elems = 0;
    uint32_t* term_bits = nullptr; void* table = nullptr;
    // persistent per-term row bitmaps: 32-bit term hash -> slot in the tile-major term_bits[tile][slot][8 words]
    FlatTermMap term_slot;
    std::vector<uint32_t> missing;                 // terms of the current batch that have no bitmap yet
    uint64_t term_version = ~0ull; int64_t term_row_words = 0;
    int32_t term_slots_used = 0, term_slots_cap = 0;
};
constexpr int BATCH_CAND_CAP = 4096;
constexpr int BATCH_TABLE_SLOTS = 16384;       // largest smem probe table (128 KB)
constexpr int BATCH_MAX_TERM_IDS = 12288;       // distinct terms one GEMM launch handles (table load <= 0.75)

}  // namespace

struct orr_store {
    orr_config cfg{};
    int sms = 148;
    float* d_emb = nullptr;
    int64_t* d_ticks = nullptr;
    uint32_t* d_terms32 = nullptr;
    uint64_t* d_terms64 = nullptr;
    int64_t rows_used = 0;
    int64_t live_rows = 0;
    uint64_t version = 0;
    std::unordered_map<uint64_t, std::vector<std::pair<int64_t, int64_t>>> docs;   // doc -> [first,count) runs
    std::vector<int64_t> h_ticks;                                                   // host mirror (lazy)
    std::shared_mutex mu;                                                           // searches shared, mutators exclusive
    std::mutex mutator_mu;                                                          // one mutator at a time (taken BEFORE mu); the bulk loader
                                                                                    // copies under it alone and takes mu only to publish
    std::mutex pool_mu;
    std::vector<std::unique_ptr<SearchCtx>> pool;
    // orr_search_device: one scratch context per IN-FLIGHT call (several host threads may enqueue on different streams);
    // a context is reused once the event recorded behind its kernels has completed.  Mutators that move rows
    // (compact, load) wait for every in-flight device search first.
    std::mutex dev_mu;
    std::vector<std::unique_ptr<SearchCtx>> dev_pool;
    SearchCtx* dev_last = nullptr;                                                  // the call orr_search_device_timing reports
    int64_t dev_rows = 0;
    std::once_flag batch_once;
    std::mutex cap_mu;
    uint64_t cap_version = ~0ull;
    int32_t cap_value = -1;
    std::vector<uint32_t> cap_rows;
    cudaStream_t mut_stream = nullptr;
    // text mode: lower-cased UTF-8 chunk content (orr_store_upsert_document_chunks_text)
    uint8_t* d_text = nullptr; uint64_t* d_text_off = nullptr; uint32_t* d_text_len = nullptr;
    uint64_t text_used = 0, text_cap = 0;
    int64_t text_rows = 0;            // rows appended WITH text; text mode needs text_rows == rows_used
    double text_bytes_per_row = 1024.0;
    bool synth_text = false;          // option "synth_text": orr_store_fill_synthetic also writes the chunk text
    // text-level ingest (orr_store_upsert_document_texts): live vocabulary + which words each document holds
    OrrVocab* vocab = nullptr;
    std::mutex vocab_mu;
    std::unordered_map<uint64_t, std::vector<uint32_t>> doc_words;    // doc -> vocabulary ids, one per (chunk, distinct word)
    std::unordered_map<uint64_t, int32_t> doc_overflow;               // doc -> chunks with more distinct tokens than term_slots
    int64_t overflow_rows = 0;        // > 0: the hashed term table is incomplete, keyword matching goes through text mode
    bool keep_text = false;           // option "keep_text": upsert_document_texts also keeps the lower-cased Content in HBM
    bool synth_vocab = false;         // option "synth_vocab": fill_synthetic registers the 2^20 synthetic tokens
    std::unique_ptr<BatchState> batch;
    int batch_passes = 0;            // 0 = auto (bf16 screen, bf16x3 for what it cannot prove), 1 = screen only, 3 = bf16x3
    std::atomic<int> batch_hold{0};  // auto mode: batches still to run bf16x3 first after a screen that mostly failed
};

namespace {

OrrShard shard_view(const orr_store* s) {
    OrrShard v;
    v.emb = s->d_emb; v.ticks = s->d_ticks; v.terms32 = s->d_terms32; v.terms64 = s->d_terms64;
    v.rows = s->rows_used; v.dim = s->cfg.dim; v.slots = s->cfg.term_slots; v.row_base = s->cfg.row_base;
    return v;
}
OrrWeights weights_of(const orr_store* s) {
    return OrrWeights{s->cfg.w_cos, s->cfg.w_kw, s->cfg.w_rec, s->cfg.recency_days};
}

void free_ctx(SearchCtx* c) {
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    else cudaDeviceSynchronize();
    cudaFree(c->sc.q); cudaFree(c->sc.cta_cands); cudaFree(c->sc.cta_floor); cudaFree(c->sc.surv_rows);
    cudaFree(c->sc.exact); cudaFree(c->sc.sel);
    if (!c->zero_copy) { cudaFree(c->sc.hits); cudaFree(c->sc.status); }
    cudaFree(c->sc.skey); cudaFree(c->sc.sel_state); cudaFree(c->sc.big);
    cudaFree(c->d_terms); cudaFree(c->d_kw_bits); cudaFreeHost(c->h_terms); cudaFree(c->d_cap_rows);
    cudaFreeHost(c->h_q); cudaFreeHost(c->h_hits); cudaFreeHost(c->h_status); cudaFreeHost(c->h_rows);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->done) cudaEventDestroy(c->done);
    if (c->stream) cudaStreamDestroy(c->stream);
}

int ensure_hits(SearchCtx* c, int k) {
    if (k <= c->hits_cap) return ORR_OK;
    int cap = std::max(k, 256);
    if (!c->zero_copy) cudaFree(c->sc.hits);
    c->sc.hits = nullptr;
    cudaFreeHost(c->h_hits); c->h_hits = nullptr;
    c->hits_cap = 0;
    ORR_CUDA_OK(cudaMallocHost(&c->h_hits, sizeof(orr_hit) * (size_t)cap));
    if (c->zero_copy) c->sc.hits = c->h_hits;
    else ORR_CUDA_OK(cudaMalloc(&c->sc.hits, sizeof(orr_hit) * (size_t)cap));
    c->hits_cap = cap;
    return ORR_OK;
}

// the reference's candidate pre-selection as a device-resident row list: re-uploaded only when the store (or the cap) changed
int capped_rows(orr_store* s, int32_t cap, std::vector<uint32_t>* out);
int ensure_cap_rows(orr_store* s, SearchCtx* c, int32_t cap, uint64_t version) {
    if (c->cap_version == version && c->cap_value == cap && c->d_cap_rows) return ORR_OK;
    if (!c->d_cap_rows) ORR_CUDA_OK(cudaMalloc(&c->d_cap_rows, sizeof(uint32_t) * ORR_SORT_MAX));
    std::vector<uint32_t> rows;
    int rc = capped_rows(s, cap, &rows);
    if (rc != ORR_OK) return rc;
    ORR_CUDA_OK(cudaStreamSynchronize(c->stream));               // h_rows may still feed an earlier upload
    memcpy(c->h_rows, rows.data(), sizeof(uint32_t) * rows.size());
    ORR_CUDA_OK(cudaMemcpyAsync(c->d_cap_rows, c->h_rows, sizeof(uint32_t) * rows.size(), cudaMemcpyHostToDevice, c->stream));
    c->cap_version = version; c->cap_value = cap; c->cap_n = (int32_t)rows.size();
    return ORR_OK;
}

// results -> host: nothing to do for a zero-copy context (the kernels wrote them there), else two D2H copies
int fetch_results(SearchCtx* c, int k) {
    if (c->zero_copy) return ORR_OK;
    ORR_CUDA_OK(cudaMemcpyAsync(c->h_status, c->sc.status, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, c->stream));
    ORR_CUDA_OK(cudaMemcpyAsync(c->h_hits, c->sc.hits, sizeof(orr_hit) * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    return ORR_OK;
}

int make_ctx(orr_store* s, std::unique_ptr<SearchCtx>& out, bool own_stream) {
    std::unique_ptr<SearchCtx> c(new SearchCtx());
    const int dim = s->cfg.dim;
    c->zero_copy = own_stream;                     // host-API contexts; orr_search_device writes to the caller's device buffers
    int rc = [&]() -> int {
        if (own_stream) ORR_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto& e : c->ev) ORR_CUDA_OK(cudaEventCreate(&e));
        ORR_CUDA_OK(cudaMalloc(&c->sc.q, sizeof(float) * (size_t)dim));
        ORR_CUDA_OK(cudaMalloc(&c->sc.cta_cands, sizeof(uint2) * (size_t)s->sms * ORR_MAX_SURVIVORS));
        ORR_CUDA_OK(cudaMalloc(&c->sc.cta_floor, sizeof(float) * (size_t)s->sms));
        ORR_CUDA_OK(cudaMalloc(&c->sc.surv_rows, sizeof(uint32_t) * ORR_SORT_MAX));
        ORR_CUDA_OK(cudaMalloc(&c->sc.exact, sizeof(OrrExact) * ORR_SORT_MAX));
        ORR_CUDA_OK(cudaMalloc(&c->sc.sel, sizeof(int32_t) * 8));
        ORR_CUDA_OK(cudaMemset(c->sc.sel, 0, sizeof(int32_t) * 8));
        ORR_CUDA_OK(cudaMallocHost(&c->h_q, sizeof(float) * (size_t)dim));
        ORR_CUDA_OK(cudaMallocHost(&c->h_status, sizeof(int32_t) * 2));
        if (c->zero_copy) c->sc.status = c->h_status;
        else ORR_CUDA_OK(cudaMalloc(&c->sc.status, sizeof(int32_t) * 2));
        ORR_CUDA_OK(cudaMallocHost(&c->h_rows, sizeof(uint32_t) * ORR_SORT_MAX));
        return ensure_hits(c.get(), 256);
    }();
    if (rc != ORR_OK) { free_ctx(c.get()); return rc; }
    out = std::move(c);
    return ORR_OK;
}

int ensure_exact_buffers(orr_store* s, SearchCtx* c, int k) {
    const int64_t need = s->cfg.capacity_rows;
    if (c->exact_rows_cap < need) {
        ORR_CUDA_OK(cudaMalloc(&c->sc.skey, sizeof(uint64_t) * (size_t)need));
        ORR_CUDA_OK(cudaMalloc(&c->sc.sel_state, orr_exact_state_bytes()));
        c->exact_rows_cap = need;
    }
    if (k > ORR_SORT_MAX && c->sc.big_cap < k) {              // the in-CTA sorter's capacity: larger k sorts in global memory
        int64_t np2 = 1;
        while (np2 < k) np2 <<= 1;
        cudaFree(c->sc.big); c->sc.big = nullptr; c->sc.big_cap = 0;
        ORR_CUDA_OK(cudaMalloc(&c->sc.big, sizeof(OrrExact) * (size_t)np2));
        c->sc.big_cap = np2;
    }
    return ORR_OK;
}

struct CtxLease {
    orr_store* s;
    std::unique_ptr<SearchCtx> c;
    CtxLease(orr_store* st) : s(st) {}
    int acquire() {
        {
            std::lock_guard<std::mutex> g(s->pool_mu);
            if (!s->pool.empty()) { c = std::move(s->pool.back()); s->pool.pop_back(); return ORR_OK; }
        }
        return make_ctx(s, c, true);
    }
    ~CtxLease() {
        if (c) { std::lock_guard<std::mutex> g(s->pool_mu); s->pool.push_back(std::move(c)); }
    }
};

int build_probes(int32_t n_terms, const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
                 OrrProbes* out) {
    memset(out, 0, sizeof *out);
    if (n_terms < 0 || n_probes < 0 || (n_probes > 0 && !probe_hash)) {
        orr_set_error("search: bad term arguments");
        return ORR_E_INVALID;
    }
    if (n_terms == 0) return ORR_OK;                                   // keyword 0 (:100-101)
    if (n_terms > ORR_MAX_QUERY_TERMS || n_probes > ORR_MAX_QUERY_PROBES) {
        orr_set_error("search: %d terms / %d probes exceed the limits (%d / %d)", n_terms, n_probes,
                      ORR_MAX_QUERY_TERMS, ORR_MAX_QUERY_PROBES);
        return ORR_E_UNSUPPORTED;
    }
    if (!probe_term && n_probes != n_terms) {
        orr_set_error("search: probe_term is NULL but n_probes != n_terms");
        return ORR_E_INVALID;
    }
    out->n_terms = n_terms;
    out->n_probes = n_probes;
    for (int i = 0; i < n_probes; ++i) {
        const int32_t t = probe_term ? probe_term[i] : i;
        if (t < 0 || t >= n_terms) { orr_set_error("search: probe_term[%d]=%d out of range", i, t); return ORR_E_INVALID; }
        const uint64_t h = probe_hash[i] ? probe_hash[i] : 1ULL;
        out->h64[i] = h;
        out->h32[i] = orr_hash_low(h);
        out->term[i] = (uint8_t)t;
    }
    return ORR_OK;
}

int survivors_for(int k) { return k <= 32 ? 64 : (k <= 96 ? 128 : 256); }

// the reference's candidate pre-selection: newest `cap` live rows, (ticks desc, row asc)
int capped_rows(orr_store* s, int32_t cap, std::vector<uint32_t>* out) {
    std::lock_guard<std::mutex> g(s->cap_mu);
    if (s->cap_version == s->version && s->cap_value == cap) { *out = s->cap_rows; return ORR_OK; }
    if ((int64_t)s->h_ticks.size() < s->rows_used) {                   // refresh the host mirror
        const size_t have = s->h_ticks.size();
        s->h_ticks.resize((size_t)s->rows_used);
        ORR_CUDA_OK(cudaMemcpy(s->h_ticks.data() + have, s->d_ticks + have,
                               sizeof(int64_t) * (s->h_ticks.size() - have), cudaMemcpyDeviceToHost));
    }
    std::vector<uint32_t> rows;
    rows.reserve((size_t)s->live_rows);
    for (int64_t i = 0; i < s->rows_used; ++i)
        if (s->h_ticks[(size_t)i] != ORR_DEAD_TICKS) rows.push_back((uint32_t)i);
    const auto& tk = s->h_ticks;
    auto newer = [&](uint32_t a, uint32_t b) { return tk[a] != tk[b] ? tk[a] > tk[b] : a < b; };
    const size_t keep = std::min<size_t>(rows.size(), (size_t)std::max(1, cap));
    std::partial_sort(rows.begin(), rows.begin() + (ptrdiff_t)keep, rows.end(), newer);
    rows.resize(keep);
    s->cap_rows = rows; s->cap_version = s->version; s->cap_value = cap;
    *out = std::move(rows);
    return ORR_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void orr_config_default(orr_config* cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->abi_version = ORR_ABI_VERSION;
    cfg->device = 0;
    cfg->dim = 3072;
    cfg->term_slots = 128;                                 // the reference's default chunk is 120 words (appsettings.json:22)
    cfg->capacity_rows = 1 << 20;
    cfg->row_base = 0;
    cfg->w_cos = 0.7; cfg->w_kw = 0.2; cfg->w_rec = 0.1;   // RecallSearchService.cs:66
    cfg->recency_days = 30.0;                              // :118
}

int orr_store_create(const orr_config* cfg, orr_store** out) {
    if (!cfg || !out) { orr_set_error("orr_store_create: NULL argument"); return ORR_E_INVALID; }
    *out = nullptr;
    if (cfg->abi_version != ORR_ABI_VERSION) { orr_set_error("ABI version %d != %d", cfg->abi_version, ORR_ABI_VERSION); return ORR_E_INVALID; }
    if (cfg->dim < 4 || cfg->dim > 8192 || (cfg->dim & 3)) { orr_set_error("dim %d must be a multiple of 4 in [4, 8192]", cfg->dim); return ORR_E_UNSUPPORTED; }
    if (cfg->term_slots != 32 && cfg->term_slots != 64 && cfg->term_slots != 128) { orr_set_error("term_slots must be 32, 64 or 128"); return ORR_E_UNSUPPORTED; }
    if (cfg->capacity_rows < 1 || cfg->capacity_rows > 0xfffffff0LL) { orr_set_error("capacity_rows out of range"); return ORR_E_INVALID; }
    if (!(cfg->recency_days > 0.0)) { orr_set_error("recency_days must be positive"); return ORR_E_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device || cfg->device < 0) {
        orr_set_error("no usable CUDA device %d (found %d); liborr has no CPU fallback", cfg->device, ndev);
        cudaGetLastError();
        return ORR_E_CUDA;
    }
    ORR_CUDA_OK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop{};
    ORR_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        orr_set_error("device %d is sm_%d%d; liborr is built for sm_100a only", cfg->device, prop.major, prop.minor);
        return ORR_E_CUDA;
    }
    std::unique_ptr<orr_store> s(new orr_store());
    s->cfg = *cfg;
    s->sms = prop.multiProcessorCount;
    if (const char* e = getenv("ORR_BATCH_PASSES")) { const int v = atoi(e); s->batch_passes = (v == 1 || v == 3) ? v : 0; }
    const size_t cap = (size_t)cfg->capacity_rows;
    int rc = [&]() -> int {
        ORR_CUDA_OK(cudaMalloc(&s->d_emb, cap * (size_t)cfg->dim * sizeof(float)));
        ORR_CUDA_OK(cudaMalloc(&s->d_ticks, cap * sizeof(int64_t)));
        ORR_CUDA_OK(cudaMalloc(&s->d_terms32, cap * (size_t)cfg->term_slots * sizeof(uint32_t)));
        ORR_CUDA_OK(cudaMalloc(&s->d_terms64, cap * (size_t)cfg->term_slots * sizeof(uint64_t)));
        ORR_CUDA_OK(cudaStreamCreateWithFlags(&s->mut_stream, cudaStreamNonBlocking));
        return ORR_OK;
    }();
    if (rc != ORR_OK) { orr_store_destroy(s.release()); return rc; }
    *out = s.release();
    return ORR_OK;
}

void orr_store_destroy(orr_store* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    for (auto& c : s->pool) free_ctx(c.get());
    for (auto& c : s->dev_pool) free_ctx(c.get());
    if (s->mut_stream) cudaStreamDestroy(s->mut_stream);
    if (s->batch) {
        BatchState* b = s->batch.get();
        if (b->stream) cudaStreamSynchronize(b->stream);
        void* ptrs[] = {b->ehi, b->emid, b->rowaux, b->q, b->qhi, b->qmid, b->qscale, b->thr, b->kww, b->qterm,
                        b->cand, b->cand_count, b->hits, b->status, b->probes, b->dense, b->term_bits, b->table, b->qbad};
        for (void* p : ptrs) cudaFree(p);
        cudaFreeHost(b->h_qterm); cudaFreeHost(b->h_kww); cudaFreeHost(b->h_probes); cudaFreeHost(b->h_table); cudaFreeHost(b->h_status);
        for (auto& e : b->ev) if (e) cudaEventDestroy(e);
        if (b->stream) cudaStreamDestroy(b->stream);
    }
    cudaFree(s->d_emb); cudaFree(s->d_ticks); cudaFree(s->d_terms32); cudaFree(s->d_terms64);
    cudaFree(s->d_text); cudaFree(s->d_text_off); cudaFree(s->d_text_len);
    orr_vocab_free(s->vocab);
    delete s;
}

int orr_store_set_option(orr_store* s, const char* name, double value) {
    if (!s || !name) { orr_set_error("orr_store_set_option: NULL argument"); return ORR_E_INVALID; }
    std::unique_lock<std::shared_mutex> lock(s->mu);
    if (!strcmp(name, "batch_passes")) {
        if (value != 0.0 && value != 1.0 && value != 3.0) { orr_set_error("batch_passes must be 0 (auto), 1 or 3"); return ORR_E_INVALID; }
        s->batch_passes = (int)value;
        s->batch_hold = 0;
        return ORR_OK;
    }
    if (!strcmp(name, "synth_text")) { s->synth_text = value != 0.0; return ORR_OK; }
    if (!strcmp(name, "synth_vocab")) { s->synth_vocab = value != 0.0; return ORR_OK; }
    if (!strcmp(name, "keep_text")) {
        if (s->rows_used != 0 && (value != 0.0) != s->keep_text) { orr_set_error("keep_text must be set before the first row arrives"); return ORR_E_INVALID; }
        s->keep_text = value != 0.0;
        return ORR_OK;
    }
    if (!strcmp(name, "text_bytes_per_row")) {
        if (s->d_text) { orr_set_error("text_bytes_per_row must be set before the first chunk text arrives"); return ORR_E_INVALID; }
        if (!(value >= 16.0 && value <= 1048576.0)) { orr_set_error("text_bytes_per_row out of range"); return ORR_E_INVALID; }
        s->text_bytes_per_row = value;
        return ORR_OK;
    }
    orr_set_error("orr_store_set_option: unknown option '%s'", name);
    return ORR_E_INVALID;
}

int64_t orr_store_count(const orr_store* s) { return s ? s->live_rows : 0; }
int64_t orr_store_rows_used(const orr_store* s) { return s ? s->rows_used : 0; }

static int ensure_text_arena(orr_store* s) {
    if (s->d_text) return ORR_OK;
    const uint64_t cap_bytes = (uint64_t)((double)s->cfg.capacity_rows * s->text_bytes_per_row) + 4096;
    ORR_CUDA_OK(cudaMalloc(&s->d_text, cap_bytes + 64));
    ORR_CUDA_OK(cudaMalloc(&s->d_text_off, sizeof(uint64_t) * (size_t)s->cfg.capacity_rows));
    ORR_CUDA_OK(cudaMalloc(&s->d_text_len, sizeof(uint32_t) * (size_t)s->cfg.capacity_rows));
    s->text_cap = cap_bytes;
    return ORR_OK;
}

static void release_doc_words(orr_store* s, uint64_t doc_key);
static int wait_device_searches(orr_store* s);

static int tombstone_locked(orr_store* s, uint64_t doc_key) {
    auto it = s->docs.find(doc_key);
    if (it == s->docs.end()) return ORR_OK;
    for (auto& run : it->second) {
        std::vector<int64_t> dead((size_t)run.second, ORR_DEAD_TICKS);
        ORR_CUDA_OK(cudaMemcpy(s->d_ticks + run.first, dead.data(), sizeof(int64_t) * dead.size(), cudaMemcpyHostToDevice));
        for (int64_t r = run.first; r < run.first + run.second && r < (int64_t)s->h_ticks.size(); ++r)
            s->h_ticks[(size_t)r] = ORR_DEAD_TICKS;
        s->live_rows -= run.second;
    }
    s->docs.erase(it);
    s->version++;
    return ORR_OK;
}

static int upsert_impl(orr_store* s, uint64_t doc_key, int32_t n, const float* emb,
                       const uint8_t* has_emb, const int64_t* created_ticks,
                       const uint64_t* term_hashes, const uint32_t* term_offsets,
                       const char* text, const uint64_t* text_offsets, uint64_t* out_rows) {
    if (!s || n < 0 || (n > 0 && !created_ticks)) { orr_set_error("upsert: bad argument"); return ORR_E_INVALID; }
    if ((text != nullptr) != (text_offsets != nullptr)) { orr_set_error("upsert: text and text_offsets go together"); return ORR_E_INVALID; }
    if (n == 0) return ORR_OK;                         // UpsertChunksAsync ignores empty batches (:19-20)
    const int dim = s->cfg.dim, slots = s->cfg.term_slots;
    for (int32_t i = 0; i < n; ++i) {
        if (created_ticks[i] == ORR_DEAD_TICKS) { orr_set_error("upsert: ticks value reserved"); return ORR_E_INVALID; }
        if (term_offsets) {
            if (term_offsets[i + 1] < term_offsets[i] || (int64_t)(term_offsets[i + 1] - term_offsets[i]) > slots) {
                orr_set_error("upsert: chunk %d has %u distinct terms, store has %d slots", i,
                              term_offsets[i + 1] - term_offsets[i], slots);
                return ORR_E_UNSUPPORTED;
            }
        }
    }
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    std::unique_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    if (s->rows_used + n > s->cfg.capacity_rows) {
        orr_set_error("upsert: store full (%lld + %d > %lld rows)", (long long)s->rows_used, n, (long long)s->cfg.capacity_rows);
        return ORR_E_OOM;
    }
    const int64_t first = s->rows_used;
    if (text_offsets) {
        for (int32_t i = 0; i < n; ++i)
            if (text_offsets[i + 1] < text_offsets[i] || text_offsets[i + 1] - text_offsets[i] > 0xffffffffull) {
                orr_set_error("upsert: bad text offsets at chunk %d", i);
                return ORR_E_INVALID;
            }
        if (s->text_rows != first) {
            orr_set_error("upsert: chunk text must be given for every row of the store or for none (%lld of %lld rows have it)",
                          (long long)s->text_rows, (long long)first);
            return ORR_E_INVALID;
        }
        int rc0 = ensure_text_arena(s);
        if (rc0 != ORR_OK) return rc0;
        const uint64_t total = text_offsets[n] - text_offsets[0];
        if (s->text_used + total > s->text_cap) {
            orr_set_error("upsert: text arena full (%llu + %llu > %llu bytes; raise option text_bytes_per_row)",
                          (unsigned long long)s->text_used, (unsigned long long)total, (unsigned long long)s->text_cap);
            return ORR_E_OOM;
        }
    }
    int rc = tombstone_locked(s, doc_key);
    if (rc != ORR_OK) return rc;
    release_doc_words(s, doc_key);
    if (text_offsets) {
        const uint64_t total = text_offsets[n] - text_offsets[0];
        std::vector<uint64_t> off((size_t)n);
        std::vector<uint32_t> len((size_t)n);
        for (int32_t i = 0; i < n; ++i) {
            off[(size_t)i] = s->text_used + (text_offsets[i] - text_offsets[0]);
            len[(size_t)i] = (uint32_t)(text_offsets[i + 1] - text_offsets[i]);
        }
        if (total) ORR_CUDA_OK(cudaMemcpy(s->d_text + s->text_used, text + text_offsets[0], total, cudaMemcpyHostToDevice));
        ORR_CUDA_OK(cudaMemcpy(s->d_text_off + first, off.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice));
        ORR_CUDA_OK(cudaMemcpy(s->d_text_len + first, len.data(), sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        s->text_used += total;
        s->text_rows += n;
    }
    if (emb) {
        ORR_CUDA_OK(cudaMemcpy(s->d_emb + first * (int64_t)dim, emb, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice));
        if (has_emb)
            for (int32_t i = 0; i < n; ++i)
                if (!has_emb[i]) ORR_CUDA_OK(cudaMemset(s->d_emb + (first + i) * (int64_t)dim, 0, sizeof(float) * (size_t)dim));
    } else {
        ORR_CUDA_OK(cudaMemset(s->d_emb + first * (int64_t)dim, 0, sizeof(float) * (size_t)n * dim));
    }
    std::vector<uint32_t> t32((size_t)n * slots, 0u);
    std::vector<uint64_t> t64((size_t)n * slots, 0ull);
    if (term_hashes && term_offsets) {
        for (int32_t i = 0; i < n; ++i) {
            int w = 0;
            for (uint32_t j = term_offsets[i]; j < term_offsets[i + 1]; ++j, ++w) {
                const uint64_t h = term_hashes[j] ? term_hashes[j] : 1ULL;
                t64[(size_t)i * slots + w] = h;
                t32[(size_t)i * slots + w] = orr_hash_low(h);
            }
        }
    }
    ORR_CUDA_OK(cudaMemcpy(s->d_terms32 + first * (int64_t)slots, t32.data(), sizeof(uint32_t) * t32.size(), cudaMemcpyHostToDevice));
    ORR_CUDA_OK(cudaMemcpy(s->d_terms64 + first * (int64_t)slots, t64.data(), sizeof(uint64_t) * t64.size(), cudaMemcpyHostToDevice));
    ORR_CUDA_OK(cudaMemcpy(s->d_ticks + first, created_ticks, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
    if ((int64_t)s->h_ticks.size() == first) s->h_ticks.insert(s->h_ticks.end(), created_ticks, created_ticks + n);
    s->docs[doc_key].push_back({first, (int64_t)n});
    s->rows_used += n;
    s->live_rows += n;
    s->version++;
    if (out_rows) for (int32_t i = 0; i < n; ++i) out_rows[i] = s->cfg.row_base + (uint64_t)(first + i);
    return ORR_OK;
}

int orr_store_upsert_document_chunks(orr_store* s, uint64_t doc_key, int32_t n, const float* emb,
                                     const uint8_t* has_emb, const int64_t* created_ticks,
                                     const uint64_t* term_hashes, const uint32_t* term_offsets,
                                     uint64_t* out_rows) {
    return upsert_impl(s, doc_key, n, emb, has_emb, created_ticks, term_hashes, term_offsets, nullptr, nullptr, out_rows);
}

int orr_store_upsert_document_chunks_text(orr_store* s, uint64_t doc_key, int32_t n, const float* emb,
                                          const uint8_t* has_emb, const int64_t* created_ticks,
                                          const uint64_t* term_hashes, const uint32_t* term_offsets,
                                          const char* text_lower_utf8, const uint64_t* text_offsets, uint64_t* out_rows) {
    if (n > 0 && (!text_lower_utf8 || !text_offsets)) { orr_set_error("upsert_text: text and text_offsets are required"); return ORR_E_INVALID; }
    return upsert_impl(s, doc_key, n, emb, has_emb, created_ticks, term_hashes, term_offsets, text_lower_utf8, text_offsets, out_rows);
}

// the document's words leave the vocabulary (their reference counts drop; a word nobody holds stops matching)
static void release_doc_words(orr_store* s, uint64_t doc_key) {
    std::lock_guard<std::mutex> g(s->vocab_mu);
    auto it = s->doc_words.find(doc_key);
    if (it != s->doc_words.end()) {
        if (s->vocab) for (uint32_t id : it->second) orr_vocab_release(s->vocab, id);
        s->doc_words.erase(it);
    }
    auto ov = s->doc_overflow.find(doc_key);
    if (ov != s->doc_overflow.end()) { s->overflow_rows -= ov->second; s->doc_overflow.erase(ov); }
}

int orr_store_delete_document(orr_store* s, uint64_t doc_key) {
    if (!s) { orr_set_error("delete: NULL store"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    std::unique_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    int rc = tombstone_locked(s, doc_key);
    if (rc == ORR_OK) release_doc_words(s, doc_key);
    return rc;
}

int64_t orr_store_vocab_size(const orr_store* s) { return (s && s->vocab) ? (int64_t)orr_vocab_live_words(s->vocab) : 0; }

// ---- text-level ingest: Content strings in, the library tokenises / hashes / tracks the vocabulary --------------
int orr_store_upsert_document_texts(orr_store* s, uint64_t doc_key, int32_t n, const float* emb, const uint8_t* has_emb,
                                    const int64_t* created_ticks, const char* contents_utf8, const uint64_t* content_offsets,
                                    uint64_t* out_rows) {
    if (!s || n < 0 || (n > 0 && (!created_ticks || !contents_utf8 || !content_offsets))) { orr_set_error("upsert_texts: bad argument"); return ORR_E_INVALID; }
    if (n == 0) return ORR_OK;
    const int slots = s->cfg.term_slots;
    std::vector<uint64_t> hashes;
    std::vector<uint32_t> hoff((size_t)n + 1, 0u);
    std::vector<std::vector<std::string>> toks((size_t)n);
    std::string text;                              // lower-cased contents (text mode)
    std::vector<uint64_t> toff((size_t)n + 1, 0ull);
    int32_t overflow = 0;
    try {
        for (int32_t i = 0; i < n; ++i) {
            if (content_offsets[i + 1] < content_offsets[i]) { orr_set_error("upsert_texts: bad content offsets at chunk %d", i); return ORR_E_INVALID; }
            const char* c = contents_utf8 + content_offsets[i];
            const int64_t cl = (int64_t)(content_offsets[i + 1] - content_offsets[i]);
            toks[(size_t)i] = orr_distinct_lower_tokens(c, cl);
            size_t keep = toks[(size_t)i].size();
            if (keep > (size_t)slots) {
                // the slot table cannot hold the chunk's term set: with the text kept in HBM the keyword side is evaluated
                // there (text mode) for as long as such a row is live; without it the chunk cannot be represented
                if (!s->keep_text) {
                    orr_set_error("upsert_texts: chunk %d has %zu distinct tokens, the store has %d term slots (create the store with "
                                  "more slots, or set option keep_text so that text mode carries it)", i, keep, slots);
                    return ORR_E_UNSUPPORTED;
                }
                keep = (size_t)slots;
                ++overflow;
            }
            for (size_t t = 0; t < keep; ++t) hashes.push_back(orr_hash_bytes(toks[(size_t)i][t].data(), (int64_t)toks[(size_t)i][t].size()));
            hoff[(size_t)i + 1] = (uint32_t)hashes.size();
            if (s->keep_text) { text += orr_lower_invariant(c, cl); toff[(size_t)i + 1] = text.size(); }
        }
    } catch (const std::exception& e) { orr_set_error("upsert_texts: %s", e.what()); return ORR_E_OOM; }
    if (hashes.empty()) hashes.push_back(0ull);
    if (text.empty()) text.push_back('\0');
    int rc = upsert_impl(s, doc_key, n, emb, has_emb, created_ticks, hashes.data(), hoff.data(),
                         s->keep_text ? text.data() : nullptr, s->keep_text ? toff.data() : nullptr, out_rows);
    if (rc != ORR_OK) return rc;
    // vocabulary: the document's previous words were released by upsert_impl's tombstoning; register the new ones
    std::lock_guard<std::mutex> g(s->vocab_mu);
    if (!s->vocab) s->vocab = orr_vocab_new(s->cfg.device);
    std::vector<uint32_t>& ids = s->doc_words[doc_key];
    for (int32_t i = 0; i < n; ++i)
        for (const std::string& w : toks[(size_t)i]) ids.push_back(orr_vocab_add(s->vocab, w.data(), w.size(), 1u));
    if (overflow) { s->doc_overflow[doc_key] += overflow; s->overflow_rows += overflow; }
    return ORR_OK;
}

// ---- bulk ingest (warm load / hydration of a scan-everything store) ---------------------------
// Many documents per call.  What the per-document entry point does under the exclusive lock — five synchronous copies per
// document — is split here:
//   1. tokenising / lower-casing / hashing of every chunk on all host cores (no lock);
//   2. the new rows are written BEHIND rows_used through two pinned staging buffers on the mutator stream, under the
//      mutator mutex only: searches keep running, they cannot see rows beyond rows_used;
//   3. one short exclusive section publishes: replaced documents' old rows are tombstoned by ONE kernel over their run
//      list, rows_used / live_rows / version move, the document table and the vocabulary are updated.
}  // extern "C"

namespace {
__global__ void orr_tombstone_runs_kernel(int64_t* ticks, const int64_t* runs /* [n][2] first,count */, int n_runs) {
    for (int r = blockIdx.x; r < n_runs; r += gridDim.x) {
        const int64_t first = runs[2 * r], count = runs[2 * r + 1];
        for (int64_t i = threadIdx.x; i < count; i += blockDim.x) ticks[first + i] = ORR_DEAD_TICKS;
    }
}

struct Staging {                      // two pinned buffers used alternately: fill one while the other's copy is in flight
    static constexpr size_t BYTES = (size_t)32 << 20;
    uint8_t* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int turn = 0;
    cudaStream_t st = nullptr;
    int init(cudaStream_t stream) {
        st = stream;
        for (int i = 0; i < 2; ++i) {
            ORR_CUDA_OK(cudaMallocHost(&buf[i], BYTES));
            ORR_CUDA_OK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return ORR_OK;
    }
    ~Staging() { for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); } }
    // copies `bytes` produced by fill(dst_host, offset, n) slice by slice to `dev`
    template <class Fill>
    int send(void* dev, size_t bytes, size_t granule, Fill fill) {
        const size_t slice = std::max(granule, BYTES / granule * granule);
        for (size_t off = 0; off < bytes; off += slice) {
            const size_t n = std::min(slice, bytes - off);
            ORR_CUDA_OK(cudaEventSynchronize(ev[turn]));          // the buffer's previous copy has left it
            fill(buf[turn], off, n);
            ORR_CUDA_OK(cudaMemcpyAsync((uint8_t*)dev + off, buf[turn], n, cudaMemcpyHostToDevice, st));
            ORR_CUDA_OK(cudaEventRecord(ev[turn], st));
            turn ^= 1;
        }
        return ORR_OK;
    }
};

template <class F>
void parallel_chunks(int64_t n, F fn) {
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const int64_t T = std::max<int64_t>(1, std::min<int64_t>(hw, n / 64 + 1));
    std::vector<std::thread> th;
    for (int64_t t = 1; t < T; ++t) th.emplace_back([=] { fn(n * t / T, n * (t + 1) / T, (int)t); });
    fn(0, n / T, 0);
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

int orr_store_upsert_documents_texts(orr_store* s, int32_t n_docs, const uint64_t* doc_keys, const uint32_t* doc_chunk_offsets,
                                     const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
                                     const char* contents_utf8, const uint64_t* content_offsets, uint64_t* out_rows) {
    if (!s || n_docs < 0 || (n_docs > 0 && (!doc_keys || !doc_chunk_offsets || !created_ticks || !contents_utf8 || !content_offsets))) {
        orr_set_error("upsert_documents: bad argument");
        return ORR_E_INVALID;
    }
    if (n_docs == 0) return ORR_OK;
    const int64_t total = doc_chunk_offsets[n_docs];
    const int dim = s->cfg.dim, slots = s->cfg.term_slots;
    {
        std::unordered_map<uint64_t, int> seen;
        for (int32_t d = 0; d < n_docs; ++d) {
            if (doc_chunk_offsets[d + 1] < doc_chunk_offsets[d]) { orr_set_error("upsert_documents: chunk offsets must not decrease"); return ORR_E_INVALID; }
            if (!seen.emplace(doc_keys[d], d).second) { orr_set_error("upsert_documents: document key %llu appears twice in one call", (unsigned long long)doc_keys[d]); return ORR_E_INVALID; }
        }
    }
    for (int64_t i = 0; i < total; ++i) {
        if (created_ticks[i] == ORR_DEAD_TICKS) { orr_set_error("upsert_documents: ticks value reserved"); return ORR_E_INVALID; }
        if (content_offsets[i + 1] < content_offsets[i]) { orr_set_error("upsert_documents: bad content offsets at chunk %lld", (long long)i); return ORR_E_INVALID; }
    }
    if (total == 0) return ORR_OK;
    // ---- 1. every chunk's distinct lower-cased tokens (and lower-cased text), on all host cores ----
    std::vector<std::vector<std::string>> toks((size_t)total);
    std::vector<std::string> lowered(s->keep_text ? (size_t)total : 0);
    std::atomic<int64_t> too_long{-1};
    try {
        parallel_chunks(total, [&](int64_t lo, int64_t hi, int) {
            for (int64_t i = lo; i < hi; ++i) {
                const char* c = contents_utf8 + content_offsets[i];
                const int64_t cl = (int64_t)(content_offsets[i + 1] - content_offsets[i]);
                toks[(size_t)i] = orr_distinct_lower_tokens(c, cl);
                if ((int64_t)toks[(size_t)i].size() > slots && !s->keep_text) too_long.store(i);
                if (s->keep_text) lowered[(size_t)i] = orr_lower_invariant(c, cl);
            }
        });
    } catch (const std::exception& e) { orr_set_error("upsert_documents: %s", e.what()); return ORR_E_OOM; }
    if (too_long.load() >= 0) {
        orr_set_error("upsert_documents: chunk %lld has more distinct tokens than the store's %d term slots (more slots, or option keep_text)",
                      (long long)too_long.load(), slots);
        return ORR_E_UNSUPPORTED;
    }
    std::vector<uint64_t> toff;
    uint64_t text_total = 0;
    if (s->keep_text) {
        toff.resize((size_t)total + 1, 0ull);
        for (int64_t i = 0; i < total; ++i) { text_total += lowered[(size_t)i].size(); toff[(size_t)i + 1] = text_total; }
    }

    // ---- 2. rows written behind rows_used; searches keep running ----
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    const int64_t first = s->rows_used;                        // only mutators move it, and this thread is the mutator
    if (first + total > s->cfg.capacity_rows) {
        orr_set_error("upsert_documents: store full (%lld + %lld > %lld rows)", (long long)first, (long long)total, (long long)s->cfg.capacity_rows);
        return ORR_E_OOM;
    }
    if (s->keep_text) {
        if (s->text_rows != first) { orr_set_error("upsert_documents: chunk text must be given for every row of the store or for none"); return ORR_E_INVALID; }
        int rc0 = ensure_text_arena(s);
        if (rc0 != ORR_OK) return rc0;
        if (s->text_used + text_total > s->text_cap) { orr_set_error("upsert_documents: text arena full (raise option text_bytes_per_row)"); return ORR_E_OOM; }
    }
    cudaStream_t st = s->mut_stream;
    Staging stage;
    int rc = stage.init(st);
    if (rc != ORR_OK) return rc;
    const size_t row_bytes = (size_t)dim * sizeof(float);
    rc = stage.send(s->d_emb + first * (int64_t)dim, (size_t)total * row_bytes, row_bytes, [&](uint8_t* dst, size_t off, size_t n) {
        const int64_t r0 = (int64_t)(off / row_bytes), nr = (int64_t)(n / row_bytes);
        if (emb) memcpy(dst, (const uint8_t*)emb + off, n); else memset(dst, 0, n);
        if (emb && has_emb)
            for (int64_t r = 0; r < nr; ++r) if (!has_emb[r0 + r]) memset(dst + (size_t)r * row_bytes, 0, row_bytes);   // :71-72 -> cosine 0
    });
    if (rc != ORR_OK) return rc;
    rc = stage.send(s->d_ticks + first, (size_t)total * 8, 8, [&](uint8_t* dst, size_t off, size_t n) { memcpy(dst, (const uint8_t*)created_ticks + off, n); });
    if (rc != ORR_OK) return rc;
    int64_t overflow_total = 0;
    std::vector<int32_t> doc_over((size_t)n_docs, 0);
    for (int32_t d = 0; d < n_docs; ++d)
        for (uint32_t i = doc_chunk_offsets[d]; i < doc_chunk_offsets[d + 1]; ++i)
            if ((int64_t)toks[i].size() > slots) { ++doc_over[(size_t)d]; ++overflow_total; }
    auto fill_terms = [&](uint8_t* dst, size_t off, size_t n, bool wide) {
        const size_t rb = (size_t)slots * (wide ? 8 : 4);
        const int64_t r0 = (int64_t)(off / rb), nr = (int64_t)(n / rb);
        memset(dst, 0, n);
        parallel_chunks(nr, [&](int64_t lo, int64_t hi, int) {
            for (int64_t r = lo; r < hi; ++r) {
                const auto& tk = toks[(size_t)(r0 + r)];
                const size_t keep = std::min(tk.size(), (size_t)slots);
                for (size_t w = 0; w < keep; ++w) {
                    const uint64_t h = orr_hash_bytes(tk[w].data(), (int64_t)tk[w].size());
                    if (wide) ((uint64_t*)dst)[(size_t)r * slots + w] = h; else ((uint32_t*)dst)[(size_t)r * slots + w] = orr_hash_low(h);
                }
            }
        });
    };
    rc = stage.send(s->d_terms32 + first * (int64_t)slots, (size_t)total * slots * 4, (size_t)slots * 4,
                    [&](uint8_t* dst, size_t off, size_t n) { fill_terms(dst, off, n, false); });
    if (rc != ORR_OK) return rc;
    rc = stage.send(s->d_terms64 + first * (int64_t)slots, (size_t)total * slots * 8, (size_t)slots * 8,
                    [&](uint8_t* dst, size_t off, size_t n) { fill_terms(dst, off, n, true); });
    if (rc != ORR_OK) return rc;
    if (s->keep_text) {
        std::vector<uint64_t> off_abs((size_t)total);
        std::vector<uint32_t> len((size_t)total);
        for (int64_t i = 0; i < total; ++i) { off_abs[(size_t)i] = s->text_used + toff[(size_t)i]; len[(size_t)i] = (uint32_t)lowered[(size_t)i].size(); }
        rc = stage.send(s->d_text_off + first, (size_t)total * 8, 8, [&](uint8_t* dst, size_t off, size_t n) { memcpy(dst, (const uint8_t*)off_abs.data() + off, n); });
        if (rc != ORR_OK) return rc;
        rc = stage.send(s->d_text_len + first, (size_t)total * 4, 4, [&](uint8_t* dst, size_t off, size_t n) { memcpy(dst, (const uint8_t*)len.data() + off, n); });
        if (rc != ORR_OK) return rc;
        int64_t cursor = 0;                                     // chunk whose text contains byte `off` (slices arrive in order)
        rc = stage.send(s->d_text + s->text_used, (size_t)text_total, 1, [&](uint8_t* dst, size_t off, size_t n) {
            size_t done = 0;
            while (done < n) {
                while (toff[(size_t)cursor + 1] <= off + done) ++cursor;
                const size_t in = (size_t)(off + done - toff[(size_t)cursor]);
                const size_t m = std::min(n - done, lowered[(size_t)cursor].size() - in);
                memcpy(dst + done, lowered[(size_t)cursor].data() + in, m);
                done += m;
            }
        });
        if (rc != ORR_OK) return rc;
    }
    ORR_CUDA_OK(cudaStreamSynchronize(st));

    // ---- 3. publish: one short exclusive section ----
    std::vector<int64_t> runs;
    int64_t* d_runs = nullptr;
    {
        std::unique_lock<std::shared_mutex> lock(s->mu);
        int64_t dead_rows = 0;
        for (int32_t d = 0; d < n_docs; ++d) {
            auto it = s->docs.find(doc_keys[d]);
            if (it == s->docs.end()) continue;
            for (auto& run : it->second) {
                runs.push_back(run.first); runs.push_back(run.second);
                dead_rows += run.second;
                for (int64_t r = run.first; r < run.first + run.second && r < (int64_t)s->h_ticks.size(); ++r) s->h_ticks[(size_t)r] = ORR_DEAD_TICKS;
            }
            s->docs.erase(it);
        }
        if (!runs.empty()) {
            rc = [&]() -> int {
                ORR_CUDA_OK(cudaMalloc(&d_runs, runs.size() * sizeof(int64_t)));
                ORR_CUDA_OK(cudaMemcpyAsync(d_runs, runs.data(), runs.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                const int nr = (int)(runs.size() / 2);
                orr_tombstone_runs_kernel<<<std::min(nr, 1024), 128, 0, st>>>(s->d_ticks, d_runs, nr);
                ORR_CUDA_OK(cudaGetLastError());
                ORR_CUDA_OK(cudaStreamSynchronize(st));
                return ORR_OK;
            }();
            cudaFree(d_runs);
            if (rc != ORR_OK) return rc;
        }
        if ((int64_t)s->h_ticks.size() == first) s->h_ticks.insert(s->h_ticks.end(), created_ticks, created_ticks + total);
        for (int32_t d = 0; d < n_docs; ++d) {
            const int64_t n = (int64_t)doc_chunk_offsets[d + 1] - doc_chunk_offsets[d];
            if (n > 0) s->docs[doc_keys[d]].push_back({first + doc_chunk_offsets[d], n});
        }
        if (s->keep_text) { s->text_used += text_total; s->text_rows += total; }
        s->rows_used += total;
        s->live_rows += total - dead_rows;
        s->version++;
    }
    if (out_rows) for (int64_t i = 0; i < total; ++i) out_rows[i] = s->cfg.row_base + (uint64_t)(first + i);

    // ---- vocabulary: the replaced documents' words leave, the new ones arrive (ids resolved on all cores) ----
    for (int32_t d = 0; d < n_docs; ++d) release_doc_words(s, doc_keys[d]);
    std::lock_guard<std::mutex> g(s->vocab_mu);
    if (!s->vocab) s->vocab = orr_vocab_new(s->cfg.device);
    std::vector<std::vector<uint32_t>> ids((size_t)total);
    parallel_chunks(total, [&](int64_t lo, int64_t hi, int) {            // read-only lookups: nobody adds words meanwhile
        for (int64_t i = lo; i < hi; ++i) {
            auto& v = ids[(size_t)i];
            v.resize(toks[(size_t)i].size());
            for (size_t w = 0; w < v.size(); ++w) v[w] = orr_vocab_find(s->vocab, toks[(size_t)i][w].data(), toks[(size_t)i][w].size());
        }
    });
    for (int32_t d = 0; d < n_docs; ++d) {
        if (doc_chunk_offsets[d + 1] == doc_chunk_offsets[d]) continue;
        std::vector<uint32_t>& mine = s->doc_words[doc_keys[d]];
        for (uint32_t i = doc_chunk_offsets[d]; i < doc_chunk_offsets[d + 1]; ++i)
            for (size_t w = 0; w < ids[i].size(); ++w) {
                uint32_t id = ids[i][w];
                if (id == 0xffffffffu) id = orr_vocab_add(s->vocab, toks[i][w].data(), toks[i][w].size(), 1u);   // a new word
                else orr_vocab_addref(s->vocab, id, 1u);
                mine.push_back(id);
            }
        if (doc_over[(size_t)d]) { s->doc_overflow[doc_keys[d]] += doc_over[(size_t)d]; }
    }
    s->overflow_rows += overflow_total;
    return ORR_OK;
}

int orr_store_fill_synthetic(orr_store* s, const orr_synth_spec* spec, uint64_t first_row, int64_t n) {
    if (!s || !spec || n < 0) { orr_set_error("fill_synthetic: bad argument"); return ORR_E_INVALID; }
    if (spec->dim != s->cfg.dim || spec->terms_per_chunk > s->cfg.term_slots || spec->gen_dim < spec->dim ||
        spec->terms_per_chunk < 0 || spec->vocab != (1 << 20)) {
        orr_set_error("fill_synthetic: spec does not match the store");
        return ORR_E_INVALID;
    }
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    std::unique_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    if (s->rows_used + n > s->cfg.capacity_rows) { orr_set_error("fill_synthetic: store full"); return ORR_E_OOM; }
    const bool with_text = s->synth_text && s->text_rows == s->rows_used;
    const uint64_t text_len = spec->terms_per_chunk > 0 ? (uint64_t)(9 * spec->terms_per_chunk - 1) : 0;
    if (with_text) {
        int rc0 = ensure_text_arena(s);
        if (rc0 != ORR_OK) return rc0;
        if (s->text_used + text_len * (uint64_t)n > s->text_cap) { orr_set_error("fill_synthetic: text arena full (raise text_bytes_per_row)"); return ORR_E_OOM; }
    }
    int rc = orr_launch_synth_fill(s->d_emb, s->d_ticks, s->d_terms32, s->d_terms64, s->cfg.dim, s->cfg.term_slots,
                                   *spec, first_row, s->rows_used, n, with_text ? s->d_text : nullptr, s->d_text_off,
                                   s->d_text_len, s->text_used, s->mut_stream);
    if (rc != ORR_OK) return rc;
    ORR_CUDA_OK(cudaStreamSynchronize(s->mut_stream));
    if (with_text) { s->text_used += text_len * (uint64_t)n; s->text_rows += n; }
    s->rows_used += n;
    s->live_rows += n;
    s->version++;
    if (s->synth_vocab && spec->terms_per_chunk > 0) {
        // the synthetic corpus draws its tokens from "t%07d", id < 2^20: register them all once (orr_search_query expands
        // query terms over the vocabulary, as it does for ingested text)
        std::lock_guard<std::mutex> g(s->vocab_mu);
        if (!s->vocab) s->vocab = orr_vocab_new(s->cfg.device);
        if (orr_vocab_words(s->vocab) == 0) {
            char w[16];
            for (uint32_t id = 0; id < (1u << 20); ++id) {
                snprintf(w, sizeof w, "t%07u", id);
                orr_vocab_add(s->vocab, w, 8, 1u);
            }
        }
    }
    return ORR_OK;
}

// ---- search --------------------------------------------------------------------------------
// The exact path end to end: E1 (every row's score key), then rounds of digit passes + gather until the radix walk has
// ended (two passes on ordinary scores).  Leaves {n_out, flags} and the hits in the context's pinned buffers.
// `kw` carries text mode's row bitmaps (or NULLs).
static int run_exact(orr_store* s, SearchCtx* c, const OrrShard& sh, const OrrProbes& pr, int64_t now_ticks,
                     int q_dim, int top_k, cudaEvent_t after_scores = nullptr, const uint32_t* kw_bits = nullptr,
                     int64_t kw_row_words = 0, int kw_terms = 0) {
    const int k = (int)std::min<int64_t>(std::max(1, top_k), std::max<int64_t>(1, s->live_rows));
    int rc = ensure_exact_buffers(s, c, k);
    if (rc != ORR_OK) return rc;
    rc = ensure_hits(c, k);
    if (rc != ORR_OK) return rc;
    OrrScratch sc = c->sc;
    sc.kw_bits = kw_bits; sc.kw_row_words = kw_row_words; sc.kw_terms = kw_terms;
    const OrrWeights w = weights_of(s);
    for (int attempt = 0; attempt < 2; ++attempt) {
        // attempt 0 may score on the 32-bit term table (a screen the gather verifies); attempt 1 = the 64-bit kernel
        const bool force_general = attempt == 1;
        const bool screened = orr_exact_keys_are_screened(sc, pr, q_dim, force_general);
        rc = orr_launch_exact_scores(sh, sc, pr, w, now_ticks, q_dim, c->stream, force_general);
        if (rc != ORR_OK) return rc;
        if (after_scores && attempt == 0) ORR_CUDA_OK(cudaEventRecord(after_scores, c->stream));
        for (int first = 1, n = 2;; first += n, n = 3) {
            rc = orr_launch_exact_select(sh, sc, k, first, n, c->stream, screened ? &pr : nullptr, screened ? &w : nullptr, now_ticks);
            if (rc != ORR_OK) return rc;
            rc = fetch_results(c, k);
            if (rc != ORR_OK) return rc;
            ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
            if (c->h_status[1] & ORR_EXACT_INTERNAL) { orr_set_error("exact path: selection histogram inconsistent"); return ORR_E_INTERNAL; }
            if (!(c->h_status[1] & ORR_EXACT_INCOMPLETE)) break;
            if (first + n > ORR_EXACT_PASSES + 3) { orr_set_error("exact path: selection did not terminate"); return ORR_E_INTERNAL; }
        }
        if (!(c->h_status[1] & ORR_EXACT_UNPROVEN)) break;
        if (attempt == 1) { orr_set_error("exact path: unproven selection from exact keys"); return ORR_E_INTERNAL; }
    }
    return ORR_OK;
}

int orr_search(orr_store* s, const float* q, int32_t q_dim, int32_t n_terms, const uint64_t* probe_hash,
               const int32_t* probe_term, int32_t n_probes, int64_t now_ticks, int32_t top_k,
               int32_t candidate_cap, orr_hit* out, int32_t* n_out) {
    const double t0 = now_ms();
    if (!s || !out || !n_out || q_dim < 0 || (q_dim > 0 && !q) || candidate_cap < 0) {
        orr_set_error("orr_search: bad argument");
        return ORR_E_INVALID;
    }
    *n_out = 0;
    OrrProbes pr;
    int rc = build_probes(n_terms, probe_hash, probe_term, n_probes, &pr);
    if (rc != ORR_OK) return rc;
    const int k = std::max(1, top_k);                                   // Math.Max(1, topK) :36
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    memset(&g_timing, 0, sizeof g_timing);
    if (s->live_rows == 0) { g_timing.wall_ms = (float)(now_ms() - t0); return ORR_OK; }   // empty store: no citations
    CtxLease lease(s);
    rc = lease.acquire();
    if (rc != ORR_OK) return rc;
    SearchCtx* c = lease.c.get();
    const OrrShard sh = shard_view(s);
    // :71-72 — a query whose length differs from the stored width scores cosine 0 everywhere
    const int eff_q_dim = (q_dim == s->cfg.dim) ? q_dim : 0;
    const int64_t kk = std::min<int64_t>(k, s->rows_used);
    rc = ensure_hits(c, (int)kk);
    if (rc != ORR_OK) return rc;
    if (eff_q_dim > 0) {
        memcpy(c->h_q, q, sizeof(float) * (size_t)q_dim);
        ORR_CUDA_OK(cudaMemcpyAsync(c->sc.q, c->h_q, sizeof(float) * (size_t)q_dim, cudaMemcpyHostToDevice, c->stream));
    }
    int path = 0;
    ORR_CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    if (candidate_cap > 0) {
        path = ORR_PATH_SUBSET;
        if (candidate_cap > ORR_SORT_MAX) { orr_set_error("candidate_cap %d > %d", candidate_cap, ORR_SORT_MAX); return ORR_E_UNSUPPORTED; }
        // the capped row list only changes when the store does: it stays in this context's device buffer between queries
        rc = ensure_cap_rows(s, c, candidate_cap, s->version);
        if (rc != ORR_OK) return rc;
        const int32_t nl = c->cap_n;
        OrrScratch sc = c->sc;
        sc.surv_rows = c->d_cap_rows;
        ORR_CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
        rc = orr_launch_rescore(sh, sc, pr, weights_of(s), now_ticks, eff_q_dim, (int)kk, nl, false, c->stream, nl);
        if (rc != ORR_OK) return rc;
        g_timing.n_survivors = nl;
    } else if (eff_q_dim == 0 || k > ORR_FUSED_MAX_K) {
        path = ORR_PATH_EXACT;                                          // scan_ms = the scoring kernel, finalize_ms = the selection
        rc = run_exact(s, c, sh, pr, now_ticks, eff_q_dim, (int)kk, c->ev[1]);
        if (rc != ORR_OK) return rc;
    } else {
        path = ORR_PATH_FUSED;
        const int M = survivors_for(k);
        rc = orr_launch_scan(sh, c->sc, pr, weights_of(s), now_ticks, M, s->sms, c->stream);
        if (rc != ORR_OK) return rc;
        ORR_CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
        rc = orr_launch_rescore(sh, c->sc, pr, weights_of(s), now_ticks, eff_q_dim, k, M, true, c->stream);
        if (rc != ORR_OK) return rc;
        g_timing.n_survivors = M;
    }
    ORR_CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
    if (path != ORR_PATH_EXACT) {                                       // the exact path has already brought its results down
        rc = fetch_results(c, (int)kk);
        if (rc != ORR_OK) return rc;
    }
    ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
    float scan_ms = 0.f, fin_ms = 0.f;
    cudaEventElapsedTime(&scan_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&fin_ms, c->ev[1], c->ev[2]);
    if (path == ORR_PATH_FUSED && (c->h_status[1] & 1)) {
        // K3 could not prove the fp32 selection safe (ties / near-ties at the survivor
        // boundary): re-run on the exact path.
        path = ORR_PATH_EXACT | ORR_PATH_ESCALATED;
        ORR_CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
        rc = run_exact(s, c, sh, pr, now_ticks, eff_q_dim, (int)kk);
        if (rc != ORR_OK) return rc;
        ORR_CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
        ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
        float extra = 0.f;
        cudaEventElapsedTime(&extra, c->ev[1], c->ev[2]);
        fin_ms += extra;
    }
    const int got = std::min<int>(c->h_status[0], (int)kk);
    memcpy(out, c->h_hits, sizeof(orr_hit) * (size_t)got);
    *n_out = got;
    g_timing.scan_ms = scan_ms;
    g_timing.finalize_ms = fin_ms;
    g_timing.total_device_ms = scan_ms + fin_ms;
    g_timing.path = path;
    g_timing.rows_scanned = (candidate_cap > 0) ? g_timing.n_survivors : s->rows_used;
    g_timing.wall_ms = (float)(now_ms() - t0);
    return ORR_OK;
}

// ---- text mode: the keyword predicate evaluated on the chunk text itself --------------------------
int orr_search_text(orr_store* s, const float* q, int32_t q_dim, int32_t n_terms, const char* terms_lower_utf8,
                    const uint32_t* term_offsets, int64_t now_ticks, int32_t top_k, int32_t candidate_cap,
                    orr_hit* out, int32_t* n_out) {
    const double t0 = now_ms();
    if (!s || !out || !n_out || q_dim < 0 || (q_dim > 0 && !q) || candidate_cap < 0 || n_terms < 0 ||
        (n_terms > 0 && (!terms_lower_utf8 || !term_offsets))) {
        orr_set_error("orr_search_text: bad argument");
        return ORR_E_INVALID;
    }
    *n_out = 0;
    if (n_terms > ORR_MAX_QUERY_TERMS) { orr_set_error("orr_search_text: %d terms > %d", n_terms, ORR_MAX_QUERY_TERMS); return ORR_E_UNSUPPORTED; }
    int max_len = 0;
    for (int32_t t = 0; t < n_terms; ++t) {
        if (term_offsets[t + 1] <= term_offsets[t]) { orr_set_error("orr_search_text: term %d is empty", t); return ORR_E_INVALID; }
        max_len = std::max<int>(max_len, (int)(term_offsets[t + 1] - term_offsets[t]));
    }
    if (n_terms > 0 && (max_len > ORR_TEXT_MAX_TERM_BYTES || term_offsets[n_terms] - term_offsets[0] > (uint32_t)ORR_TEXT_TERMS_BYTES)) {
        orr_set_error("orr_search_text: terms too long (longest %d bytes, limit %d; total limit %d)", max_len,
                      ORR_TEXT_MAX_TERM_BYTES, ORR_TEXT_TERMS_BYTES);
        return ORR_E_UNSUPPORTED;
    }
    const int k = std::max(1, top_k);
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    memset(&g_timing, 0, sizeof g_timing);
    if (s->live_rows == 0) { g_timing.wall_ms = (float)(now_ms() - t0); return ORR_OK; }
    if (n_terms > 0 && s->text_rows != s->rows_used) {
        orr_set_error("orr_search_text: the store holds chunk text for %lld of %lld rows (use orr_store_upsert_document_chunks_text)",
                      (long long)s->text_rows, (long long)s->rows_used);
        return ORR_E_INVALID;
    }
    CtxLease lease(s);
    int rc = lease.acquire();
    if (rc != ORR_OK) return rc;
    SearchCtx* c = lease.c.get();
    const OrrShard sh = shard_view(s);
    const int eff_q_dim = (q_dim == s->cfg.dim) ? q_dim : 0;
    const int64_t kk = std::min<int64_t>(k, s->rows_used);
    rc = ensure_hits(c, (int)kk);
    if (rc != ORR_OK) return rc;
    if (eff_q_dim > 0) {
        memcpy(c->h_q, q, sizeof(float) * (size_t)q_dim);
        ORR_CUDA_OK(cudaMemcpyAsync(c->sc.q, c->h_q, sizeof(float) * (size_t)q_dim, cudaMemcpyHostToDevice, c->stream));
    }
    const int64_t row_words = (s->rows_used + 31) / 32;
    bool escalated = false, exact = false;
    int32_t n_sub = 0;
    if (candidate_cap > 0) {
        if (candidate_cap > ORR_SORT_MAX) { orr_set_error("candidate_cap %d > %d", candidate_cap, ORR_SORT_MAX); return ORR_E_UNSUPPORTED; }
        rc = ensure_cap_rows(s, c, candidate_cap, s->version);
        if (rc != ORR_OK) return rc;
        n_sub = c->cap_n;
    }
    ORR_CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    OrrScratch sc = c->sc;
    if (n_terms > 0) {
        if (!c->d_terms) {
            ORR_CUDA_OK(cudaMalloc(&c->d_terms, sizeof(OrrTextTerms)));
            ORR_CUDA_OK(cudaMallocHost(&c->h_terms, sizeof(OrrTextTerms)));
        }
        const size_t need = (size_t)n_terms * (size_t)row_words;
        if (need > c->kw_bits_words) {
            cudaFree(c->d_kw_bits); c->d_kw_bits = nullptr; c->kw_bits_words = 0;
            const size_t cap_words = (size_t)ORR_MAX_QUERY_TERMS * (size_t)((s->cfg.capacity_rows + 31) / 32);
            const size_t alloc = std::min(cap_words, std::max(need, (size_t)8 * (size_t)((s->cfg.capacity_rows + 31) / 32)));
            ORR_CUDA_OK(cudaMalloc(&c->d_kw_bits, alloc * sizeof(uint32_t)));
            c->kw_bits_words = alloc;
        }
        OrrTextTerms* ht = c->h_terms;
        memset(ht, 0, sizeof *ht);
        ht->n_terms = n_terms; ht->max_len = max_len;
        for (int32_t t = 0; t <= n_terms; ++t) ht->off[t] = (uint16_t)(term_offsets[t] - term_offsets[0]);
        memcpy(ht->bytes, terms_lower_utf8 + term_offsets[0], term_offsets[n_terms] - term_offsets[0]);
        ORR_CUDA_OK(cudaMemcpyAsync(c->d_terms, ht, sizeof *ht, cudaMemcpyHostToDevice, c->stream));
        OrrTextView tv{s->d_text, s->d_text_off, s->d_text_len};
        if (candidate_cap > 0) {
            ORR_CUDA_OK(cudaMemsetAsync(c->d_kw_bits, 0, need * sizeof(uint32_t), c->stream));
            rc = orr_launch_text_bits(tv, s->rows_used, c->d_terms, n_terms, max_len, c->d_cap_rows, n_sub,
                                      c->d_kw_bits, row_words, c->stream);
        } else {
            rc = orr_launch_text_bits(tv, s->rows_used, c->d_terms, n_terms, max_len, nullptr, 0, c->d_kw_bits, row_words, c->stream);
        }
        if (rc != ORR_OK) return rc;
        sc.kw_bits = c->d_kw_bits; sc.kw_row_words = row_words; sc.kw_terms = n_terms;
    }
    OrrProbes pr;
    memset(&pr, 0, sizeof pr);
    pr.n_terms = n_terms;                                   // the denominator; matches come from the bitmaps
    ORR_CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    if (candidate_cap > 0) {
        const int32_t nl = n_sub;
        sc.surv_rows = c->d_cap_rows;
        rc = orr_launch_rescore(sh, sc, pr, weights_of(s), now_ticks, eff_q_dim, (int)kk, nl, false, c->stream, nl);
        g_timing.n_survivors = nl;
    } else {
        // every row is a candidate: the fused scan selects with the bitmaps as its keyword side (<= 32 terms),
        // K3 re-scores exactly and proves the selection; otherwise (or when the proof fails) the full fp64 pass
        bool fused = eff_q_dim > 0 && k <= ORR_FUSED_MAX_K && n_terms <= 32;
        if (fused) {
            const int M = survivors_for(k);
            rc = orr_launch_scan(sh, sc, pr, weights_of(s), now_ticks, M, s->sms, c->stream);
            if (rc != ORR_OK) return rc;
            rc = orr_launch_rescore(sh, sc, pr, weights_of(s), now_ticks, eff_q_dim, k, M, true, c->stream);
            if (rc != ORR_OK) return rc;
            if (!c->zero_copy) ORR_CUDA_OK(cudaMemcpyAsync(c->h_status, c->sc.status, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, c->stream));
            ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
            g_timing.n_survivors = M;
            if (c->h_status[1] & 1) { fused = false; escalated = true; }
        }
        if (!fused) {
            exact = true;
            rc = n_terms > 0 ? run_exact(s, c, sh, pr, now_ticks, eff_q_dim, (int)kk, nullptr, c->d_kw_bits, row_words, n_terms)
                             : run_exact(s, c, sh, pr, now_ticks, eff_q_dim, (int)kk);
        }
    }
    if (rc != ORR_OK) return rc;
    ORR_CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
    if (!exact) {                                            // the exact path has already brought its results down
        rc = fetch_results(c, (int)kk);
        if (rc != ORR_OK) return rc;
    }
    ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
    float match_ms = 0.f, fin_ms = 0.f;
    cudaEventElapsedTime(&match_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&fin_ms, c->ev[1], c->ev[2]);
    const int got = std::min<int>(c->h_status[0], (int)kk);
    memcpy(out, c->h_hits, sizeof(orr_hit) * (size_t)got);
    *n_out = got;
    g_timing.scan_ms = match_ms;                             // substring kernel
    g_timing.finalize_ms = fin_ms;                           // fused scan + K3 (or the full fp64 pass) + ordering
    g_timing.total_device_ms = match_ms + fin_ms;
    g_timing.path = ORR_PATH_TEXT | (escalated ? ORR_PATH_ESCALATED : 0);
    g_timing.rows_scanned = (candidate_cap > 0) ? g_timing.n_survivors : s->rows_used;
    g_timing.wall_ms = (float)(now_ms() - t0);
    return ORR_OK;
}

// ---- query string in, hits out: KeywordScore's query side + vocabulary expansion + search, in one call ----------------
// RecallSearchService.cs:95-108 (split, lower-case, distinct, stop words with the all-stop-words fallback) runs here, the
// terms are expanded over the live vocabulary on the GPU (orr_vocab.cu: which words contain each term, :110-111), and the
// scan is orr_search's.  keyword_mode 0 = auto (hashed probes; text mode when a term sits in more than 128 words, when
// there are more than 64 terms, or while a chunk with more distinct tokens than term_slots is live), 1 = hashed only,
// 2 = text only.
int orr_search_query(orr_store* s, const char* query_utf8, int32_t query_len, const float* q, int32_t q_dim, int64_t now_ticks,
                     int32_t top_k, int32_t candidate_cap, int32_t keyword_mode, orr_hit* out, int32_t* n_out) {
    if (!s || !out || !n_out || query_len < 0 || (query_len > 0 && !query_utf8) || keyword_mode < 0 || keyword_mode > 2) {
        orr_set_error("orr_search_query: bad argument");
        return ORR_E_INVALID;
    }
    *n_out = 0;
    std::vector<std::string> terms;
    try {
        std::vector<std::string> raw = orr_distinct_lower_tokens(query_utf8, query_len);          // :95-98
        if (raw.empty()) { orr_set_error("Query is required."); return ORR_E_INVALID; }           // :22-23 (IsNullOrWhiteSpace)
        for (const auto& t : raw) if (!orr_is_stop_word(t)) terms.push_back(t);                   // :103-105
        if (terms.empty()) terms = raw;                                                           // :107-108
    } catch (const std::exception& e) { orr_set_error("orr_search_query: %s", e.what()); return ORR_E_OOM; }
    const int nt = (int)terms.size();
    const bool has_text = s->d_text && s->text_rows == s->rows_used;
    if (keyword_mode != 2 && nt <= ORR_MAX_QUERY_TERMS && s->overflow_rows == 0) {
        uint64_t ph[ORR_MAX_QUERY_PROBES];
        int32_t pt[ORR_MAX_QUERY_PROBES];
        int32_t np = 0;
        int rc = ORR_OK;
        {
            std::lock_guard<std::mutex> g(s->vocab_mu);
            if (s->vocab) rc = orr_vocab_expand(s->vocab, terms, ph, pt, ORR_MAX_QUERY_PROBES, &np);
            else if (s->live_rows > 0) {
                orr_set_error("orr_search_query: the store has no vocabulary (rows must arrive through orr_store_upsert_document_texts)");
                return ORR_E_INVALID;
            }
        }
        if (rc != ORR_OK && rc != ORR_E_UNSUPPORTED) return rc;
        if (rc == ORR_OK && np <= ORR_MAX_QUERY_PROBES)
            return orr_search(s, q, q_dim, nt, ph, pt, np, now_ticks, top_k, candidate_cap, out, n_out);   // np may be 0: no live word holds any term
        if (keyword_mode == 1 || !has_text) {
            if (rc == ORR_OK) orr_set_error("orr_search_query: the terms expand to %d vocabulary words (limit %d) and the store keeps no chunk text", np, ORR_MAX_QUERY_PROBES);
            return ORR_E_UNSUPPORTED;
        }
    } else if (keyword_mode == 1 || !has_text) {
        orr_set_error("orr_search_query: %d terms / %lld over-long chunks need text mode, and the store keeps no chunk text (option keep_text)",
                      nt, (long long)s->overflow_rows);
        return ORR_E_UNSUPPORTED;
    }
    std::string blob;
    std::vector<uint32_t> off((size_t)nt + 1, 0u);
    for (int t = 0; t < nt; ++t) { blob += terms[(size_t)t]; off[(size_t)t + 1] = (uint32_t)blob.size(); }
    return orr_search_text(s, q, q_dim, nt, blob.data(), off.data(), now_ticks, top_k, candidate_cap, out, n_out);
}

// The keyword side of a query as orr_search would receive it (inspection / tests): |terms| after A-2 filtering and the
// (hash, term) probes of the vocabulary expansion.  *n_probes may exceed cap (nothing beyond cap is written).
int orr_expand_query(orr_store* s, const char* query_utf8, int32_t query_len, uint64_t* probe_hash, int32_t* probe_term,
                     int32_t cap, int32_t* n_terms, int32_t* n_probes) {
    if (!s || !n_terms || !n_probes || query_len < 0 || (query_len > 0 && !query_utf8) || cap < 0 || (cap > 0 && (!probe_hash || !probe_term))) {
        orr_set_error("orr_expand_query: bad argument");
        return ORR_E_INVALID;
    }
    *n_terms = 0; *n_probes = 0;
    std::vector<std::string> terms;
    try {
        std::vector<std::string> raw = orr_distinct_lower_tokens(query_utf8, query_len);
        for (const auto& t : raw) if (!orr_is_stop_word(t)) terms.push_back(t);
        if (terms.empty()) terms = raw;
    } catch (const std::exception& e) { orr_set_error("orr_expand_query: %s", e.what()); return ORR_E_OOM; }
    *n_terms = (int32_t)terms.size();
    std::lock_guard<std::mutex> g(s->vocab_mu);
    if (!s->vocab || terms.empty()) return ORR_OK;
    return orr_vocab_expand(s->vocab, terms, probe_hash, probe_term, cap, n_probes);
}

// ---- snapshot / warm load ----------------------------------------------
// The HBM store is volatile; a snapshot is its byte image: header, then emb / ticks / terms32 / terms64
// (/ text_off / text_len / text) for rows [0, rows_used) and the document -> row-run table.  Row ids are
// preserved, so the host's row -> chunk map stays valid.  Device <-> file goes through a pinned bounce buffer.
namespace {
struct SnapHeader {
    char     magic[8];           // "ORRSNAP1"
    int32_t  abi_version, dim, term_slots, has_text;
    int64_t  rows_used, live_rows;
    uint64_t row_base, text_used, n_docs, n_runs;
    double   w_cos, w_kw, w_rec, recency_days;
};
constexpr size_t SNAP_BOUNCE = (size_t)64 << 20;

int snap_write(FILE* f, const void* dev, size_t bytes, void* bounce) {
    const uint8_t* p = (const uint8_t*)dev;
    while (bytes) {
        const size_t n = std::min(bytes, SNAP_BOUNCE);
        ORR_CUDA_OK(cudaMemcpy(bounce, p, n, cudaMemcpyDeviceToHost));
        if (fwrite(bounce, 1, n, f) != n) { orr_set_error("snapshot: short write"); return ORR_E_INVALID; }
        p += n; bytes -= n;
    }
    return ORR_OK;
}
int snap_read(FILE* f, void* dev, size_t bytes, void* bounce) {
    uint8_t* p = (uint8_t*)dev;
    while (bytes) {
        const size_t n = std::min(bytes, SNAP_BOUNCE);
        if (fread(bounce, 1, n, f) != n) { orr_set_error("snapshot: truncated file"); return ORR_E_INVALID; }
        ORR_CUDA_OK(cudaMemcpy(p, bounce, n, cudaMemcpyHostToDevice));
        p += n; bytes -= n;
    }
    return ORR_OK;
}
}  // namespace

int orr_store_save(orr_store* s, const char* path) {
    if (!s || !path) { orr_set_error("orr_store_save: NULL argument"); return ORR_E_INVALID; }
    std::shared_lock<std::shared_mutex> lock(s->mu);           // searches may continue; mutators wait
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    FILE* f = fopen(path, "wb");
    if (!f) { orr_set_error("orr_store_save: cannot open %s", path); return ORR_E_INVALID; }
    void* bounce = nullptr;
    int rc = [&]() -> int {
        ORR_CUDA_OK(cudaMallocHost(&bounce, SNAP_BOUNCE));
        SnapHeader h{};
        memcpy(h.magic, "ORRSNAP1", 8);
        h.abi_version = ORR_ABI_VERSION; h.dim = s->cfg.dim; h.term_slots = s->cfg.term_slots;
        h.has_text = (s->d_text && s->text_rows == s->rows_used && s->rows_used > 0) ? 1 : 0;
        h.rows_used = s->rows_used; h.live_rows = s->live_rows; h.row_base = s->cfg.row_base;
        h.text_used = h.has_text ? s->text_used : 0;
        h.n_docs = s->docs.size();
        for (auto& d : s->docs) h.n_runs += d.second.size();
        h.w_cos = s->cfg.w_cos; h.w_kw = s->cfg.w_kw; h.w_rec = s->cfg.w_rec; h.recency_days = s->cfg.recency_days;
        if (fwrite(&h, sizeof h, 1, f) != 1) { orr_set_error("snapshot: short write"); return ORR_E_INVALID; }
        const size_t n = (size_t)s->rows_used;
        int r;
        if ((r = snap_write(f, s->d_emb, n * s->cfg.dim * sizeof(float), bounce)) != ORR_OK) return r;
        if ((r = snap_write(f, s->d_ticks, n * sizeof(int64_t), bounce)) != ORR_OK) return r;
        if ((r = snap_write(f, s->d_terms32, n * s->cfg.term_slots * sizeof(uint32_t), bounce)) != ORR_OK) return r;
        if ((r = snap_write(f, s->d_terms64, n * s->cfg.term_slots * sizeof(uint64_t), bounce)) != ORR_OK) return r;
        if (h.has_text) {
            if ((r = snap_write(f, s->d_text_off, n * sizeof(uint64_t), bounce)) != ORR_OK) return r;
            if ((r = snap_write(f, s->d_text_len, n * sizeof(uint32_t), bounce)) != ORR_OK) return r;
            if ((r = snap_write(f, s->d_text, (size_t)s->text_used, bounce)) != ORR_OK) return r;
        }
        std::vector<uint64_t> table;
        table.reserve((size_t)(2 * h.n_docs + 2 * h.n_runs));
        for (auto& d : s->docs) {
            table.push_back(d.first); table.push_back((uint64_t)d.second.size());
            for (auto& run : d.second) { table.push_back((uint64_t)run.first); table.push_back((uint64_t)run.second); }
        }
        if (!table.empty() && fwrite(table.data(), sizeof(uint64_t), table.size(), f) != table.size()) {
            orr_set_error("snapshot: short write");
            return ORR_E_INVALID;
        }
        // trailer (absent in snapshots written before the vocabulary existed): "ORRVOC01", the vocabulary, the
        // document -> word ids table and the per-document counts of over-long chunks
        std::lock_guard<std::mutex> g(s->vocab_mu);
        std::vector<uint8_t> vb;
        if (s->vocab) orr_vocab_serialize(s->vocab, &vb);
        std::vector<uint64_t> t2;
        t2.push_back((uint64_t)vb.size());
        t2.push_back((uint64_t)s->doc_words.size());
        t2.push_back((uint64_t)s->doc_overflow.size());
        if (fwrite("ORRVOC01", 1, 8, f) != 8 || fwrite(t2.data(), 8, t2.size(), f) != t2.size() ||
            (!vb.empty() && fwrite(vb.data(), 1, vb.size(), f) != vb.size())) { orr_set_error("snapshot: short write"); return ORR_E_INVALID; }
        for (auto& d : s->doc_words) {
            const uint64_t hdr[2] = {d.first, (uint64_t)d.second.size()};
            if (fwrite(hdr, 8, 2, f) != 2 || (!d.second.empty() && fwrite(d.second.data(), 4, d.second.size(), f) != d.second.size())) {
                orr_set_error("snapshot: short write");
                return ORR_E_INVALID;
            }
        }
        for (auto& d : s->doc_overflow) {
            const uint64_t rec[2] = {d.first, (uint64_t)d.second};
            if (fwrite(rec, 8, 2, f) != 2) { orr_set_error("snapshot: short write"); return ORR_E_INVALID; }
        }
        return ORR_OK;
    }();
    cudaFreeHost(bounce);
    if (fclose(f) != 0 && rc == ORR_OK) { orr_set_error("orr_store_save: close failed"); rc = ORR_E_INVALID; }
    return rc;
}

int orr_store_load(orr_store* s, const char* path) {
    if (!s || !path) { orr_set_error("orr_store_load: NULL argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    std::unique_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    if (s->rows_used != 0) { orr_set_error("orr_store_load: the store is not empty"); return ORR_E_INVALID; }
    { const int wrc = wait_device_searches(s); if (wrc != ORR_OK) return wrc; }
    FILE* f = fopen(path, "rb");
    if (!f) { orr_set_error("orr_store_load: cannot open %s", path); return ORR_E_INVALID; }
    void* bounce = nullptr;
    auto body = [&]() -> int {
        SnapHeader h{};
        if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "ORRSNAP1", 8) != 0) { orr_set_error("orr_store_load: %s is not a snapshot", path); return ORR_E_INVALID; }
        if (h.abi_version != ORR_ABI_VERSION || h.dim != s->cfg.dim || h.term_slots != s->cfg.term_slots) {
            orr_set_error("orr_store_load: snapshot is dim %d / %d slots / abi %d, store is dim %d / %d slots / abi %d", h.dim,
                          h.term_slots, h.abi_version, s->cfg.dim, s->cfg.term_slots, ORR_ABI_VERSION);
            return ORR_E_INVALID;
        }
        if (h.rows_used < 0 || h.rows_used > s->cfg.capacity_rows) { orr_set_error("orr_store_load: %lld rows exceed the capacity %lld", (long long)h.rows_used, (long long)s->cfg.capacity_rows); return ORR_E_OOM; }
        ORR_CUDA_OK(cudaMallocHost(&bounce, SNAP_BOUNCE));
        const size_t n = (size_t)h.rows_used;
        int r;
        if ((r = snap_read(f, s->d_emb, n * s->cfg.dim * sizeof(float), bounce)) != ORR_OK) return r;
        if ((r = snap_read(f, s->d_ticks, n * sizeof(int64_t), bounce)) != ORR_OK) return r;
        if ((r = snap_read(f, s->d_terms32, n * s->cfg.term_slots * sizeof(uint32_t), bounce)) != ORR_OK) return r;
        if ((r = snap_read(f, s->d_terms64, n * s->cfg.term_slots * sizeof(uint64_t), bounce)) != ORR_OK) return r;
        if (h.has_text) {
            const uint64_t cap_bytes = std::max<uint64_t>((uint64_t)((double)s->cfg.capacity_rows * s->text_bytes_per_row), h.text_used) + 4096;
            if (!s->d_text) {
                ORR_CUDA_OK(cudaMalloc(&s->d_text, cap_bytes + 64));
                ORR_CUDA_OK(cudaMalloc(&s->d_text_off, sizeof(uint64_t) * (size_t)s->cfg.capacity_rows));
                ORR_CUDA_OK(cudaMalloc(&s->d_text_len, sizeof(uint32_t) * (size_t)s->cfg.capacity_rows));
                s->text_cap = cap_bytes;
            } else if (h.text_used > s->text_cap) { orr_set_error("orr_store_load: text arena too small"); return ORR_E_OOM; }
            if ((r = snap_read(f, s->d_text_off, n * sizeof(uint64_t), bounce)) != ORR_OK) return r;
            if ((r = snap_read(f, s->d_text_len, n * sizeof(uint32_t), bounce)) != ORR_OK) return r;
            if ((r = snap_read(f, s->d_text, (size_t)h.text_used, bounce)) != ORR_OK) return r;
            s->text_used = h.text_used; s->text_rows = h.rows_used;
        }
        // the document table: sized from the header, so the header is checked against the file and the rows first
        const long at = ftell(f);
        fseek(f, 0, SEEK_END);
        const long file_end = ftell(f);
        fseek(f, at, SEEK_SET);
        const uint64_t left = (at >= 0 && file_end >= at) ? (uint64_t)(file_end - at) : 0;
        if (h.n_docs > left / 16 || h.n_runs > left / 16 || 16 * (h.n_docs + h.n_runs) > left || h.n_runs > (uint64_t)h.rows_used ||
            h.live_rows < 0 || h.live_rows > h.rows_used) {
            orr_set_error("orr_store_load: corrupt snapshot header (%llu documents, %llu runs, %lld live of %lld rows)",
                          (unsigned long long)h.n_docs, (unsigned long long)h.n_runs, (long long)h.live_rows, (long long)h.rows_used);
            return ORR_E_INVALID;
        }
        if (h.row_base != s->cfg.row_base) {
            orr_set_error("orr_store_load: snapshot row_base %llu != store row_base %llu (row ids would not be preserved)",
                          (unsigned long long)h.row_base, (unsigned long long)s->cfg.row_base);
            return ORR_E_INVALID;
        }
        std::vector<uint64_t> table((size_t)(2 * h.n_docs + 2 * h.n_runs));
        if (!table.empty() && fread(table.data(), sizeof(uint64_t), table.size(), f) != table.size()) { orr_set_error("snapshot: truncated file"); return ORR_E_INVALID; }
        std::unordered_map<uint64_t, std::vector<std::pair<int64_t, int64_t>>> docs;
        size_t i = 0;
        uint64_t runs_seen = 0, rows_in_runs = 0;
        for (uint64_t d = 0; d < h.n_docs; ++d) {
            if (i + 2 > table.size()) { orr_set_error("snapshot: corrupt document table"); return ORR_E_INVALID; }
            const uint64_t key = table[i++], nr = table[i++];
            if (nr > h.n_runs - runs_seen || i + 2 * nr > table.size()) { orr_set_error("snapshot: corrupt document table"); return ORR_E_INVALID; }
            auto& runs = docs[key];
            for (uint64_t k = 0; k < nr; ++k) {
                const uint64_t first = table[i], count = table[i + 1];
                i += 2;
                if (first > (uint64_t)h.rows_used || count > (uint64_t)h.rows_used - first) { orr_set_error("snapshot: a document run lies outside the stored rows"); return ORR_E_INVALID; }
                runs.push_back({(int64_t)first, (int64_t)count});
                rows_in_runs += count;
            }
            runs_seen += nr;
        }
        if (runs_seen != h.n_runs || rows_in_runs > (uint64_t)h.live_rows) { orr_set_error("snapshot: document table does not match the header"); return ORR_E_INVALID; }
        // optional trailer: vocabulary, document -> words, over-long chunk counts
        char vmagic[8];
        std::unordered_map<uint64_t, std::vector<uint32_t>> doc_words;
        std::unordered_map<uint64_t, int32_t> doc_overflow;
        std::vector<uint8_t> vb;
        bool have_vocab = false;
        int64_t overflow_rows = 0;
        if (fread(vmagic, 1, 8, f) == 8) {
            if (memcmp(vmagic, "ORRVOC01", 8) != 0) { orr_set_error("snapshot: unknown trailer"); return ORR_E_INVALID; }
            uint64_t t2[3];
            const long at2 = ftell(f);
            const uint64_t left2 = (at2 >= 0 && file_end >= at2) ? (uint64_t)(file_end - at2) : 0;
            if (fread(t2, 8, 3, f) != 3 || t2[0] > left2 || t2[1] > left2 / 16 || t2[2] > left2 / 16) { orr_set_error("snapshot: corrupt vocabulary trailer"); return ORR_E_INVALID; }
            vb.resize((size_t)t2[0]);
            if (!vb.empty() && fread(vb.data(), 1, vb.size(), f) != vb.size()) { orr_set_error("snapshot: truncated file"); return ORR_E_INVALID; }
            for (uint64_t d = 0; d < t2[1]; ++d) {
                uint64_t hdr[2];
                if (fread(hdr, 8, 2, f) != 2 || hdr[1] > left2 / 4) { orr_set_error("snapshot: corrupt document-words table"); return ORR_E_INVALID; }
                std::vector<uint32_t>& ids = doc_words[hdr[0]];
                ids.resize((size_t)hdr[1]);
                if (!ids.empty() && fread(ids.data(), 4, ids.size(), f) != ids.size()) { orr_set_error("snapshot: truncated file"); return ORR_E_INVALID; }
            }
            for (uint64_t d = 0; d < t2[2]; ++d) {
                uint64_t rec[2];
                if (fread(rec, 8, 2, f) != 2 || rec[1] > (uint64_t)h.rows_used) { orr_set_error("snapshot: corrupt overflow table"); return ORR_E_INVALID; }
                doc_overflow[rec[0]] = (int32_t)rec[1];
                overflow_rows += (int64_t)rec[1];
            }
            have_vocab = true;
        }
        {
            std::lock_guard<std::mutex> g(s->vocab_mu);
            if (have_vocab && !vb.empty()) {
                if (!s->vocab) s->vocab = orr_vocab_new(s->cfg.device);
                if (orr_vocab_deserialize(s->vocab, vb.data(), vb.size()) != ORR_OK) { orr_set_error("snapshot: corrupt vocabulary"); return ORR_E_INVALID; }
                const uint64_t nw = orr_vocab_words(s->vocab);
                for (auto& d : doc_words) for (uint32_t id : d.second) if (id >= nw) { orr_vocab_clear(s->vocab); orr_set_error("snapshot: corrupt document-words table"); return ORR_E_INVALID; }
            } else if (s->vocab) {
                orr_vocab_clear(s->vocab);
            }
            s->doc_words.swap(doc_words);
            s->doc_overflow.swap(doc_overflow);
            s->overflow_rows = overflow_rows;
        }
        s->docs.swap(docs);
        s->rows_used = h.rows_used; s->live_rows = h.live_rows;
        s->h_ticks.clear();
        s->version++;
        return ORR_OK;
    };
    int rc;
    try { rc = body(); }                                       // a corrupt file must not throw across the C ABI
    catch (const std::exception& e) { orr_set_error("orr_store_load: %s", e.what()); rc = ORR_E_INVALID; }
    if (rc != ORR_OK) { s->text_used = 0; s->text_rows = 0; }  // the store stays empty
    if (bounce) cudaFreeHost(bounce);
    fclose(f);
    return rc;
}

// ---- compaction of tombstones ---------------------------------------------
// Replace-by-document and delete only tombstone rows; the scan still reads them.  Compaction squeezes the
// live rows to the front IN ORDER (so the stable row-order tie-break of the reference is unchanged) and
// tells the host which old row every new row was.  Output segments are gathered into a bounded temp buffer
// and copied back: output segment [o, o+m) only ever reads source rows >= o, and later segments only read
// rows >= o+m, so the in-place move is safe.
namespace {
__global__ void orr_gather_rows_kernel(const uint32_t* src, const uint32_t* idx, int64_t first, int64_t m, int words_per_row,
                                       uint32_t* tmp) {
    const int64_t total = m * words_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words_per_row;
        const int w = (int)(i - r * words_per_row);
        tmp[i] = src[(int64_t)idx[first + r] * words_per_row + w];
    }
}
__global__ void orr_copy_text_kernel(const uint8_t* old_text, const uint64_t* old_off, const uint32_t* len, const uint64_t* new_off,
                                     int64_t n, uint8_t* new_text) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = gw; r < n; r += W) {
        const uint8_t* a = old_text + old_off[r];
        uint8_t* b = new_text + new_off[r];
        const uint32_t l = len[r];
        for (uint32_t i = lane; i < l; i += 32) b[i] = a[i];
    }
}
int compact_array(void* base, size_t row_bytes, const uint32_t* d_idx, int64_t n_live, void* tmp, size_t tmp_bytes, cudaStream_t st) {
    const int words = (int)(row_bytes / 4);
    const int64_t seg = std::max<int64_t>(1, (int64_t)(tmp_bytes / row_bytes));
    for (int64_t o = 0; o < n_live; o += seg) {
        const int64_t m = std::min(seg, n_live - o);
        const int64_t total = m * words;
        const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
        orr_gather_rows_kernel<<<grid, 256, 0, st>>>((const uint32_t*)base, d_idx, o, m, words, (uint32_t*)tmp);
        ORR_CUDA_OK(cudaGetLastError());
        ORR_CUDA_OK(cudaMemcpyAsync((uint8_t*)base + (size_t)o * row_bytes, tmp, (size_t)m * row_bytes, cudaMemcpyDeviceToDevice, st));
    }
    return ORR_OK;
}
}  // namespace

int orr_store_compact(orr_store* s, uint64_t* old_rows_out, int64_t out_cap, int64_t* n_live_out) {
    if (!s || !n_live_out) { orr_set_error("orr_store_compact: NULL argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> mutator(s->mutator_mu);
    std::unique_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    { const int wrc = wait_device_searches(s); if (wrc != ORR_OK) return wrc; }   // rows are about to move under any in-flight scan
    if ((int64_t)s->h_ticks.size() < s->rows_used) {
        const size_t have = s->h_ticks.size();
        s->h_ticks.resize((size_t)s->rows_used);
        ORR_CUDA_OK(cudaMemcpy(s->h_ticks.data() + have, s->d_ticks + have, sizeof(int64_t) * (s->h_ticks.size() - have), cudaMemcpyDeviceToHost));
    }
    std::vector<uint32_t> idx;
    idx.reserve((size_t)s->live_rows);
    std::vector<int64_t> live_before((size_t)s->rows_used + 1, 0);
    for (int64_t r = 0; r < s->rows_used; ++r) {
        live_before[(size_t)r] = (int64_t)idx.size();
        if (s->h_ticks[(size_t)r] != ORR_DEAD_TICKS) idx.push_back((uint32_t)r);
    }
    const int64_t n_live = (int64_t)idx.size();
    *n_live_out = n_live;
    if (old_rows_out) {
        if (out_cap < n_live) { orr_set_error("orr_store_compact: old_rows_out holds %lld, %lld rows are live", (long long)out_cap, (long long)n_live); return ORR_E_INVALID; }
        for (int64_t i = 0; i < n_live; ++i) old_rows_out[i] = s->cfg.row_base + idx[(size_t)i];
    }
    if (n_live == s->rows_used) return ORR_OK;                       // nothing to squeeze
    cudaStream_t st = s->mut_stream;
    uint32_t* d_idx = nullptr; void* tmp = nullptr;
    uint8_t* new_text = nullptr; uint64_t* d_new_off = nullptr;
    const size_t tmp_bytes = std::max<size_t>((size_t)128 << 20, (size_t)s->cfg.dim * 4 * 64);
    int rc = [&]() -> int {
        if (n_live > 0) {
            ORR_CUDA_OK(cudaMalloc(&d_idx, sizeof(uint32_t) * (size_t)n_live));
            ORR_CUDA_OK(cudaMalloc(&tmp, tmp_bytes));
            ORR_CUDA_OK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(uint32_t) * (size_t)n_live, cudaMemcpyHostToDevice, st));
            int r;
            if ((r = compact_array(s->d_emb, (size_t)s->cfg.dim * 4, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
            if ((r = compact_array(s->d_ticks, 8, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
            if ((r = compact_array(s->d_terms32, (size_t)s->cfg.term_slots * 4, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
            if ((r = compact_array(s->d_terms64, (size_t)s->cfg.term_slots * 8, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
            if (s->d_text && s->text_rows == s->rows_used) {
                if ((r = compact_array(s->d_text_off, 8, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
                if ((r = compact_array(s->d_text_len, 4, d_idx, n_live, tmp, tmp_bytes, st)) != ORR_OK) return r;
                // squeeze the byte arena too: new offsets are the running sum of the live rows' lengths
                std::vector<uint32_t> len((size_t)n_live);
                ORR_CUDA_OK(cudaStreamSynchronize(st));
                ORR_CUDA_OK(cudaMemcpy(len.data(), s->d_text_len, sizeof(uint32_t) * (size_t)n_live, cudaMemcpyDeviceToHost));
                std::vector<uint64_t> noff((size_t)n_live);
                uint64_t used = 0;
                for (int64_t i = 0; i < n_live; ++i) { noff[(size_t)i] = used; used += len[(size_t)i]; }
                ORR_CUDA_OK(cudaMalloc(&new_text, s->text_cap + 64));
                ORR_CUDA_OK(cudaMalloc(&d_new_off, sizeof(uint64_t) * (size_t)n_live));
                ORR_CUDA_OK(cudaMemcpyAsync(d_new_off, noff.data(), sizeof(uint64_t) * (size_t)n_live, cudaMemcpyHostToDevice, st));
                orr_copy_text_kernel<<<148 * 8, 256, 0, st>>>(s->d_text, s->d_text_off, s->d_text_len, d_new_off, n_live, new_text);
                ORR_CUDA_OK(cudaGetLastError());
                ORR_CUDA_OK(cudaMemcpyAsync(s->d_text_off, d_new_off, sizeof(uint64_t) * (size_t)n_live, cudaMemcpyDeviceToDevice, st));
                ORR_CUDA_OK(cudaStreamSynchronize(st));
                std::swap(s->d_text, new_text);
                s->text_used = used;
            }
            ORR_CUDA_OK(cudaStreamSynchronize(st));
        }
        return ORR_OK;
    }();
    cudaFree(d_idx); cudaFree(tmp); cudaFree(new_text); cudaFree(d_new_off);
    if (rc != ORR_OK) return rc;
    // host bookkeeping: every run still in the table is fully live (tombstoning erases its document)
    for (auto& d : s->docs)
        for (auto& run : d.second) run.first = live_before[(size_t)run.first];
    std::vector<int64_t> nt((size_t)n_live);
    for (int64_t i = 0; i < n_live; ++i) nt[(size_t)i] = s->h_ticks[idx[(size_t)i]];
    s->h_ticks.swap(nt);
    if (s->text_rows == s->rows_used) s->text_rows = n_live;
    if (n_live == 0) s->text_used = 0;
    s->rows_used = n_live;
    s->live_rows = n_live;
    s->version++;
    if (s->batch) { std::lock_guard<std::mutex> g(s->batch->mu); s->batch->planes_rows = 0; s->batch->mid_rows = 0; }   // bf16 planes follow the rows
    return ORR_OK;
}

// diagnostic: the fp32 score the fused scan computes for every row (what its selection and tau are built from)
int orr_debug_scan_scores(orr_store* s, const float* q, int32_t q_dim, int32_t n_terms, const uint64_t* probe_hash,
                          const int32_t* probe_term, int32_t n_probes, int64_t now_ticks, float* out, int64_t out_cap) {
    if (!s || !q || !out || q_dim != s->cfg.dim) { orr_set_error("orr_debug_scan_scores: bad argument"); return ORR_E_INVALID; }
    OrrProbes pr;
    int rc = build_probes(n_terms, probe_hash, probe_term, n_probes, &pr);
    if (rc != ORR_OK) return rc;
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    if (out_cap < s->rows_used) { orr_set_error("orr_debug_scan_scores: out holds %lld, store has %lld rows", (long long)out_cap, (long long)s->rows_used); return ORR_E_INVALID; }
    if (s->rows_used == 0) return ORR_OK;
    CtxLease lease(s);
    rc = lease.acquire();
    if (rc != ORR_OK) return rc;
    SearchCtx* c = lease.c.get();
    float* d_dense = nullptr;
    ORR_CUDA_OK(cudaMalloc(&d_dense, sizeof(float) * (size_t)s->rows_used));
    rc = [&]() -> int {
        ORR_CUDA_OK(cudaMemcpyAsync(c->sc.q, q, sizeof(float) * (size_t)q_dim, cudaMemcpyHostToDevice, c->stream));
        OrrScratch sc = c->sc;
        sc.scan_dense = d_dense;
        int r = orr_launch_scan(shard_view(s), sc, pr, weights_of(s), now_ticks, 64, s->sms, c->stream);
        if (r != ORR_OK) return r;
        ORR_CUDA_OK(cudaMemcpyAsync(out, d_dense, sizeof(float) * (size_t)s->rows_used, cudaMemcpyDeviceToHost, c->stream));
        ORR_CUDA_OK(cudaStreamSynchronize(c->stream));
        return ORR_OK;
    }();
    cudaFree(d_dense);
    return rc;
}

int orr_search_device(orr_store* s, const float* q_dev, int32_t q_dim, int32_t n_terms,
                      const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
                      int64_t now_ticks, int32_t top_k, orr_hit* out_dev, int32_t* status_dev,
                      void* cuda_stream) {
    if (!s || !q_dev || !out_dev || !status_dev) { orr_set_error("orr_search_device: NULL argument"); return ORR_E_INVALID; }
    if (q_dim != s->cfg.dim) { orr_set_error("orr_search_device: q_dim %d != store dim %d", q_dim, s->cfg.dim); return ORR_E_INVALID; }
    const int k = std::max(1, top_k);
    if (k > ORR_FUSED_MAX_K) { orr_set_error("orr_search_device: top_k %d > %d", k, ORR_FUSED_MAX_K); return ORR_E_UNSUPPORTED; }
    OrrProbes pr;
    int rc = build_probes(n_terms, probe_hash, probe_term, n_probes, &pr);
    if (rc != ORR_OK) return rc;
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (s->live_rows == 0) { ORR_CUDA_OK(cudaMemsetAsync(status_dev, 0, 2 * sizeof(int32_t), st)); return ORR_OK; }
    std::lock_guard<std::mutex> g(s->dev_mu);
    // lease a scratch context nobody's kernels are using: two calls on two streams must not share tickets / candidate buffers
    SearchCtx* c = nullptr;
    for (auto& cand : s->dev_pool) {
        if (cand->busy && cudaEventQuery(cand->done) == cudaSuccess) cand->busy = false;
        if (!cand->busy) { c = cand.get(); break; }
    }
    cudaGetLastError();                                              // cudaErrorNotReady from the queries above is not an error
    if (!c) {
        if (s->dev_pool.size() >= 64) {                              // 64 searches in flight: wait for the oldest instead of growing
            c = s->dev_pool.front().get();
            ORR_CUDA_OK(cudaEventSynchronize(c->done));
            c->busy = false;
        } else {
            std::unique_ptr<SearchCtx> fresh;
            rc = make_ctx(s, fresh, false);
            if (rc != ORR_OK) return rc;
            ORR_CUDA_OK(cudaEventCreateWithFlags(&fresh->done, cudaEventDisableTiming));
            c = fresh.get();
            s->dev_pool.push_back(std::move(fresh));
        }
    }
    OrrScratch sc = c->sc;
    sc.q = const_cast<float*>(q_dev);
    sc.hits = out_dev;
    sc.status = status_dev;
    const OrrShard sh = shard_view(s);
    const int M = survivors_for(k);
    ORR_CUDA_OK(cudaEventRecord(c->ev[0], st));
    rc = orr_launch_scan(sh, sc, pr, weights_of(s), now_ticks, M, s->sms, st);
    if (rc == ORR_OK) {
        cudaEventRecord(c->ev[1], st);
        rc = orr_launch_rescore(sh, sc, pr, weights_of(s), now_ticks, q_dim, k, M, true, st);
    }
    // whatever was enqueued keeps the context busy until it has run, also on a failed launch
    c->busy = true;
    ORR_CUDA_OK(cudaEventRecord(c->ev[2], st));
    ORR_CUDA_OK(cudaEventRecord(c->done, st));
    if (rc != ORR_OK) return rc;
    s->dev_last = c;
    s->dev_rows = s->rows_used;
    return ORR_OK;
}

// compact / load move rows: no orr_search_device kernel may still be reading them (the shared lock of the enqueue is long
// gone by the time the kernels run).  Called with the exclusive lock held, so no new device search can be enqueued.
int wait_device_searches(orr_store* s) {
    std::lock_guard<std::mutex> g(s->dev_mu);
    for (auto& c : s->dev_pool)
        if (c->busy) { ORR_CUDA_OK(cudaEventSynchronize(c->done)); c->busy = false; }
    return ORR_OK;
}

int orr_search_device_timing(orr_store* s, orr_timing* out) {
    if (!s || !out) { orr_set_error("orr_search_device_timing: NULL argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(s->dev_mu);
    memset(out, 0, sizeof *out);
    if (!s->dev_last) { orr_set_error("orr_search_device_timing: no orr_search_device call yet"); return ORR_E_INVALID; }
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    SearchCtx* c = s->dev_last;
    ORR_CUDA_OK(cudaEventSynchronize(c->ev[2]));
    ORR_CUDA_OK(cudaEventElapsedTime(&out->scan_ms, c->ev[0], c->ev[1]));
    ORR_CUDA_OK(cudaEventElapsedTime(&out->finalize_ms, c->ev[1], c->ev[2]));
    out->total_device_ms = out->scan_ms + out->finalize_ms;
    out->path = ORR_PATH_FUSED;
    out->rows_scanned = s->dev_rows;
    return ORR_OK;
}

// ---- batched path ------------------------------------------------------------------------------
static int batch_prepare(orr_store* s, BatchState* bs, int batch_padded, int k, bool need_mid) {
    const int dim = s->cfg.dim;
    const size_t cap = (size_t)s->cfg.capacity_rows;
    if (!bs->stream) {
        ORR_CUDA_OK(cudaStreamCreateWithFlags(&bs->stream, cudaStreamNonBlocking));
        for (auto& e : bs->ev) ORR_CUDA_OK(cudaEventCreate(&e));
    }
    if (!bs->ehi) {
        const size_t cap_pad = (cap + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE * ORR_BATCH_TILE;
        ORR_CUDA_OK(cudaMalloc(&bs->ehi, (size_t)orr_batch_plane_elems((int64_t)cap, dim) * 2));
        ORR_CUDA_OK(cudaMalloc(&bs->rowaux, cap_pad * sizeof(float)));
        bs->planes_rows = 0;
    }
    if (bs->planes_rows < s->rows_used) {        // rows appended since the last batch
        int rc = orr_batch_build_planes(s->d_emb, bs->ehi, nullptr, bs->planes_rows, s->rows_used - bs->planes_rows, dim,
                                        (float)s->cfg.w_cos, bs->stream);
        if (rc != ORR_OK) return rc;
        bs->planes_rows = s->rows_used;
    }
    // the mid plane (the other half of the split) costs as much HBM as the hi plane and is only read by bf16x3
    // passes: it is allocated and built the first time one runs (auto mode: the first cascade)
    if (need_mid) {
        if (!bs->emid) { ORR_CUDA_OK(cudaMalloc(&bs->emid, (size_t)orr_batch_plane_elems((int64_t)cap, dim) * 2)); bs->mid_rows = 0; }
        if (bs->mid_rows < s->rows_used) {
            int rc = orr_batch_build_planes(s->d_emb, nullptr, bs->emid, bs->mid_rows, s->rows_used - bs->mid_rows, dim,
                                            (float)s->cfg.w_cos, bs->stream);
            if (rc != ORR_OK) return rc;
            bs->mid_rows = s->rows_used;
        }
    }
    if (batch_padded > bs->bcap || k > bs->kcap) {
        void** ptrs[] = {(void**)&bs->q, &bs->qhi, &bs->qmid, (void**)&bs->qscale, (void**)&bs->thr, (void**)&bs->kww,
                         (void**)&bs->qterm, &bs->cand, (void**)&bs->cand_count, (void**)&bs->hits, (void**)&bs->status,
                         (void**)&bs->probes, (void**)&bs->qbad};
        for (void** p : ptrs) { cudaFree(*p); *p = nullptr; }
        cudaFreeHost(bs->h_qterm); cudaFreeHost(bs->h_kww); cudaFreeHost(bs->h_probes); cudaFreeHost(bs->h_status);
        bs->h_qterm = nullptr; bs->h_kww = nullptr; bs->h_probes = nullptr; bs->h_status = nullptr;
        const size_t B = (size_t)std::max(batch_padded, bs->bcap);
        const size_t K = (size_t)std::max(k, bs->kcap);
        ORR_CUDA_OK(cudaMalloc(&bs->q, B * dim * sizeof(float)));
        ORR_CUDA_OK(cudaMalloc(&bs->qhi, B * dim * 2));
        ORR_CUDA_OK(cudaMalloc(&bs->qmid, B * dim * 2));
        ORR_CUDA_OK(cudaMalloc(&bs->qscale, B * sizeof(float)));
        ORR_CUDA_OK(cudaMalloc(&bs->thr, B * sizeof(float)));
        ORR_CUDA_OK(cudaMalloc(&bs->qbad, B * sizeof(int32_t)));
        ORR_CUDA_OK(cudaMalloc(&bs->kww, B * sizeof(float)));
        ORR_CUDA_OK(cudaMalloc(&bs->qterm, B * ORR_BATCH_TERMS * sizeof(int32_t)));
        ORR_CUDA_OK(cudaMalloc(&bs->cand, B * BATCH_CAND_CAP * 8));
        ORR_CUDA_OK(cudaMalloc(&bs->cand_count, B * sizeof(uint32_t)));
        ORR_CUDA_OK(cudaMalloc(&bs->hits, B * K * sizeof(orr_hit)));
        ORR_CUDA_OK(cudaMalloc(&bs->status, B * 2 * sizeof(int32_t)));
        ORR_CUDA_OK(cudaMalloc(&bs->probes, B * sizeof(OrrBatchProbes)));
        ORR_CUDA_OK(cudaMallocHost(&bs->h_qterm, B * ORR_BATCH_TERMS * sizeof(int32_t)));
        ORR_CUDA_OK(cudaMallocHost(&bs->h_kww, B * sizeof(float)));
        ORR_CUDA_OK(cudaMallocHost(&bs->h_probes, B * sizeof(OrrBatchProbes)));
        ORR_CUDA_OK(cudaMallocHost(&bs->h_status, B * 2 * sizeof(int32_t)));
        if (!bs->h_table) ORR_CUDA_OK(cudaMallocHost(&bs->h_table, BATCH_TABLE_SLOTS * sizeof(uint2)));
        bs->bcap = (int)B; bs->kcap = (int)K;
    }
    return ORR_OK;
}

// rows re-scored exactly per query: the GEMM screen must keep every row whose screen score is within
// eps of the k-th best, so the bf16-only screen (eps ~ 5.5e-3) keeps a deeper list than bf16x3 (eps ~ 1e-4)
static int batch_survivors(int k, int passes) {
    if (passes == 1) return std::min(ORR_BATCH_MAX_SURV, std::max(128, (4 * k + 64 + 63) / 64 * 64));
    return std::min(256, std::max(64, (k + std::max(32, k) + 31) / 32 * 32));
}
static double batch_eps(const OrrWeights& w, int passes) {
    const double base = (double)ORR_BATCH_EPS * (std::fabs(w.w_cos) + std::fabs(w.w_kw) + std::fabs(w.w_rec));
    // bf16 operands: each product is off by < 2*2^-8 + 2^-16 relative, so |d cos| < 2^-7 by Cauchy-Schwarz
    return passes == 1 ? base + 0.0079 * std::fabs(w.w_cos) : base;
}

// the GEMM path proper; `redo` receives the queries whose selection could not be proven safe
// out_dev / n_out_dev (both or neither): the hits stay in HBM — `out` is then not written
static int batch_gemm_path(orr_store* s, int32_t batch, const float* q, const int32_t* n_terms,
                           const uint64_t* probe_hash, const uint32_t* probe_offsets, int64_t now_ticks, int32_t top_k,
                           orr_hit* out, int32_t* n_out, int passes, std::vector<int32_t>* redo,
                           orr_hit* out_dev = nullptr, int32_t* n_out_dev = nullptr) {
    std::call_once(s->batch_once, [&] { s->batch.reset(new BatchState()); });    // searches hold the lock SHARED: create once
    BatchState* bs = s->batch.get();
    std::lock_guard<std::mutex> g(bs->mu);
    const int dim = s->cfg.dim, k = std::max(1, top_k);
    const int bp = (batch + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE * ORR_BATCH_TILE;
    int rc = batch_prepare(s, bs, bp, k, passes == 3);
    if (rc != ORR_OK) return rc;
    cudaStream_t st = bs->stream;
    const OrrShard sh = shard_view(s);
    const OrrWeights w = weights_of(s);
    const int64_t rows = s->rows_used;
    const int64_t rows_pad = (rows + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE * ORR_BATCH_TILE;
    const int M = batch_survivors(k, passes);

    ORR_CUDA_OK(cudaMemcpyAsync(bs->q, q, sizeof(float) * (size_t)batch * dim, cudaMemcpyHostToDevice, st));
    ORR_CUDA_OK(cudaEventRecord(bs->ev[3], st));               // queries resident in HBM from here on
    rc = orr_batch_prep_queries(bs->q, bs->qhi, bs->qmid, bs->qscale, bs->qbad, batch, bp, dim, st);
    if (rc != ORR_OK) return rc;
    rc = orr_batch_build_rowrec(s->d_ticks, bs->rowaux, rows, rows_pad, now_ticks, w, st);
    if (rc != ORR_OK) return rc;

    // ---- keyword side: distinct batch terms -> ids, bitmaps over rows, per-query id lists ----
    bool any_terms = false;
    for (int32_t b = 0; b < batch; ++b) any_terms |= (n_terms && n_terms[b] > 0);
    const int64_t row_words = rows_pad / 32;
    if (any_terms) {
        int64_t total_terms = 0;
        for (int32_t b = 0; b < batch; ++b) total_terms += std::max(0, n_terms[b]);
        if (total_terms > BATCH_MAX_TERM_IDS) { orr_set_error("batch path: %lld query terms in one launch", (long long)total_terms); return ORR_E_INTERNAL; }
        // Term bitmaps persist across batches (keyed by the 32-bit term hash) until the store mutates:
        // a batch only builds the bitmaps of terms it is the first to ask for.
        if (bs->term_slot.mask == 0) bs->term_slot.reset((size_t)ORR_BATCH_MAX_TERM_SLOTS + BATCH_MAX_TERM_IDS);
        if (bs->term_version != s->version || bs->term_row_words != row_words) {
            bs->term_slot.clear(); bs->term_slots_used = 0;
            bs->term_version = s->version; bs->term_row_words = row_words;
        }
        int32_t* qterm = bs->h_qterm;
        float* kww = bs->h_kww;
        OrrBatchProbes* hp = bs->h_probes;
        std::fill(qterm, qterm + (size_t)bp * ORR_BATCH_TERMS, -1);
        std::fill(kww, kww + bp, 0.f);
        memset(hp, 0, sizeof(OrrBatchProbes) * (size_t)batch);
        std::vector<uint32_t>& missing = bs->missing;
        int32_t first_new = bs->term_slots_used;
        // one pass over the batch's terms: known terms resolve to their slot, unseen ones take the next slots in order
        auto assign_slots = [&]() {
            missing.clear();
            first_new = bs->term_slots_used;
            for (int32_t b = 0; b < batch; ++b) {
                const int32_t nt = n_terms[b];
                hp[b].n_terms = nt;
                if (nt > 0) kww[b] = (float)(w.w_kw / (double)nt);
                for (int32_t t = 0; t < nt; ++t) {
                    const uint64_t h = probe_hash[probe_offsets[b] + t] ? probe_hash[probe_offsets[b] + t] : 1ULL;
                    hp[b].h64[t] = h;
                    const uint32_t h32 = orr_hash_low(h);
                    int32_t slot = bs->term_slot.get(h32);
                    if (slot < 0) {
                        slot = bs->term_slots_used++;
                        bs->term_slot.put(h32, slot);
                        missing.push_back(h32);
                    }
                    qterm[(size_t)b * ORR_BATCH_TERMS + t] = slot;
                }
            }
        };
        assign_slots();
        if (bs->term_slots_used > bs->term_slots_cap) {
            // out of slots: drop the cache (every term of this batch is unseen again); grow the pool if this batch
            // alone does not fit
            bs->term_slot.clear(); bs->term_slots_used = 0;
            assign_slots();
            const size_t n_distinct = missing.size();
            if ((int64_t)n_distinct > bs->term_slots_cap) {
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                // sized for the store's CAPACITY (32 B per slot and 256-row tile), so appended rows never outgrow it
                const size_t slot_bytes = (size_t)((s->cfg.capacity_rows + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE) * 32;
                const size_t have = (size_t)bs->term_slots_cap * slot_bytes;
                size_t budget = std::min<size_t>((free_b + have) / 4, (size_t)12 << 30);
                size_t want = std::max<size_t>(8 * n_distinct, 16384);   // room for several batches before the pool recycles
                if (want * slot_bytes > budget) want = std::max<size_t>(n_distinct, budget / slot_bytes);
                want = std::min<size_t>(want, ORR_BATCH_MAX_TERM_SLOTS);
                cudaFree(bs->term_bits); bs->term_bits = nullptr; bs->term_slots_cap = 0;
                ORR_CUDA_OK(cudaMalloc(&bs->term_bits, want * slot_bytes));
                bs->term_slots_cap = (int32_t)want;
            }
        }
        if (!missing.empty()) {
            int table_slots = 256;
            while (table_slots < 2 * (int64_t)missing.size() && table_slots < BATCH_TABLE_SLOTS) table_slots <<= 1;   // load <= 0.5: the kernel's filter keeps most hashes away from it
            uint2* table = bs->h_table;
            std::fill(table, table + table_slots, make_uint2(0u, 0u));
            for (size_t i = 0; i < missing.size(); ++i) {
                const uint32_t h = missing[i];
                uint32_t pos = (h * 0x9E3779B1u) & (uint32_t)(table_slots - 1);
                while (table[pos].x != 0u) pos = (pos + 1) & (uint32_t)(table_slots - 1);
                table[pos] = make_uint2(h, (uint32_t)(first_new + (int32_t)i));
            }
            if (!bs->table) ORR_CUDA_OK(cudaMalloc(&bs->table, BATCH_TABLE_SLOTS * 8));
            ORR_CUDA_OK(cudaMemcpyAsync(bs->table, table, (size_t)table_slots * 8, cudaMemcpyHostToDevice, st));
            rc = orr_batch_launch_term_bits(s->d_terms32, s->cfg.term_slots, rows, bs->table, table_slots, bs->term_bits,
                                            bs->term_slots_cap, first_new, (int)missing.size(), st);
            if (rc != ORR_OK) return rc;
        }
        ORR_CUDA_OK(cudaMemcpyAsync(bs->qterm, qterm, (size_t)bp * ORR_BATCH_TERMS * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        ORR_CUDA_OK(cudaMemcpyAsync(bs->kww, kww, (size_t)bp * sizeof(float), cudaMemcpyHostToDevice, st));
        ORR_CUDA_OK(cudaMemcpyAsync(bs->probes, hp, sizeof(OrrBatchProbes) * (size_t)batch, cudaMemcpyHostToDevice, st));
        g_batch_terms_built = (int32_t)missing.size();
    }

    OrrBatchGemm gm{};
    gm.qhi = bs->qhi; gm.qmid = bs->qmid; gm.ehi = bs->ehi; gm.emid = bs->emid; gm.rowaux = bs->rowaux;
    gm.qscale = bs->qscale; gm.thr = bs->thr; gm.cand = bs->cand; gm.cand_count = bs->cand_count; gm.cand_cap = BATCH_CAND_CAP;
    gm.term_bits = any_terms ? bs->term_bits : nullptr; gm.slot_cap = bs->term_slots_cap;
    gm.q_term_ids = any_terms ? bs->qterm : nullptr; gm.q_kw_w = any_terms ? bs->kww : nullptr;
    gm.rows = rows; gm.dim = dim; gm.batch_padded = bp; gm.sms = s->sms; gm.passes = passes;

    // ---- sampling pass: dense scores of every stride-th row tile -> per-query thresholds ----
    // thr[b] = the rstar-th best sampled score, rstar ~ ORR_BATCH_SAMPLE_HITS, chosen so that about
    // `target` rows of the whole store pass it (cand_cap leaves 4-16x head room for the estimate's noise)
    const int target = std::min(4 * M, BATCH_CAND_CAP / 4);
    const int64_t all_tiles = rows_pad / ORR_BATCH_TILE;
    const int64_t want_tiles = std::max<int64_t>(8, ((int64_t)ORR_BATCH_SAMPLE_HITS * rows / target + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE);
    const int stride = (int)std::max<int64_t>(1, all_tiles / std::min(all_tiles, want_tiles));
    const int64_t s_tiles = (all_tiles + stride - 1) / stride;
    const int64_t n_s = s_tiles * ORR_BATCH_TILE;
    if ((size_t)bp * (size_t)n_s > bs->dense_elems) {
        cudaFree(bs->dense); bs->dense = nullptr; bs->dense_elems = 0;
        ORR_CUDA_OK(cudaMalloc(&bs->dense, (size_t)bp * (size_t)n_s * sizeof(float)));
        bs->dense_elems = (size_t)bp * (size_t)n_s;
    }
    gm.dense = bs->dense; gm.dense_ld = n_s; gm.tile_stride = stride;
    gm.dense_half = 1;                           // fp16 is plenty for a threshold estimate: half the stores and reads
    ORR_CUDA_OK(cudaEventRecord(bs->ev[0], st));
    rc = orr_batch_launch_gemm(gm, st);
    if (rc != ORR_OK) return rc;
    const int rstar = stride == 1 ? target : (int)std::max<int64_t>(1, (int64_t)target * n_s / rows_pad);
    rc = orr_batch_launch_threshold(bs->dense, 1, n_s, (int)n_s, rstar, bs->thr, batch, bp, st);
    if (rc != ORR_OK) return rc;

    // ---- main pass: every row tile, candidates above the thresholds ----
    ORR_CUDA_OK(cudaMemsetAsync(bs->cand_count, 0, sizeof(uint32_t) * (size_t)bp, st));
    gm.dense = nullptr; gm.dense_ld = 0; gm.dense_half = 0; gm.tile_stride = 1;
    ORR_CUDA_OK(cudaEventRecord(bs->ev[1], st));
    rc = orr_batch_launch_gemm(gm, st);
    if (rc != ORR_OK) return rc;
    ORR_CUDA_OK(cudaEventRecord(bs->ev[2], st));

    // ---- per-query finalize: survivors, exact fp64 re-score, order, bound check ----
    const double eps = batch_eps(w, passes);
    rc = orr_batch_launch_finalize(sh, bs->q, dim, any_terms ? bs->probes : nullptr, w, now_ticks, bs->cand, bs->cand_count,
                                   bs->thr, BATCH_CAND_CAP, M, top_k, k, eps, bs->qbad, bs->hits, bs->status, batch, st);
    if (rc != ORR_OK) return rc;
    int32_t* st_host = bs->h_status;
    ORR_CUDA_OK(cudaEventRecord(bs->ev[4], st));
    ORR_CUDA_OK(cudaMemcpyAsync(st_host, bs->status, sizeof(int32_t) * 2 * (size_t)batch, cudaMemcpyDeviceToHost, st));
    if (out_dev) {
        ORR_CUDA_OK(cudaMemcpyAsync(out_dev, bs->hits, sizeof(orr_hit) * (size_t)batch * k, cudaMemcpyDeviceToDevice, st));
        ORR_CUDA_OK(cudaMemcpy2DAsync(n_out_dev, sizeof(int32_t), bs->status, 2 * sizeof(int32_t), sizeof(int32_t), (size_t)batch,
                                      cudaMemcpyDeviceToDevice, st));       // status is {n_out, flags} per query
    } else {
        ORR_CUDA_OK(cudaMemcpyAsync(out, bs->hits, sizeof(orr_hit) * (size_t)batch * k, cudaMemcpyDeviceToHost, st));
    }
    ORR_CUDA_OK(cudaEventRecord(bs->ev[5], st));
    ORR_CUDA_OK(cudaStreamSynchronize(st));
    float ms_sample = 0.f, ms_main = 0.f, ms_all = 0.f;
    cudaEventElapsedTime(&ms_sample, bs->ev[0], bs->ev[1]);
    cudaEventElapsedTime(&ms_main, bs->ev[1], bs->ev[2]);
    cudaEventElapsedTime(&ms_all, bs->ev[3], bs->ev[4]);
    if (getenv("ORR_BATCH_TRACE")) {
        float ms_prep = 0.f, ms_fin = 0.f, ms_d2h = 0.f;
        cudaEventElapsedTime(&ms_prep, bs->ev[3], bs->ev[0]);
        cudaEventElapsedTime(&ms_fin, bs->ev[2], bs->ev[4]);
        cudaEventElapsedTime(&ms_d2h, bs->ev[4], bs->ev[5]);
        fprintf(stderr, "[orr batch] B=%d k=%d passes=%d M=%d stride=%d terms_built=%d: prep %.3f  sample %.3f  main %.3f  finalize %.3f  d2h %.3f ms\n",
                batch, k, passes, M, stride, g_batch_terms_built, ms_prep, ms_sample, ms_main, ms_fin, ms_d2h);
    }
    for (int32_t b = 0; b < batch; ++b) {
        if (n_out) n_out[b] = st_host[(size_t)2 * b];
        if (st_host[(size_t)2 * b + 1] != 0) redo->push_back(b);
    }
    g_timing.scan_ms = ms_main;                  // the main tcgen05 pass
    g_timing.finalize_ms = ms_all - ms_main;     // query/row prep, term bitmaps, sampling pass, thresholds, exact re-rank
    g_timing.total_device_ms = ms_all;           // queries in HBM -> hits in HBM
    g_timing.path = ORR_PATH_BATCH;
    g_timing.n_survivors = M;
    g_timing.rows_scanned = rows;
    return ORR_OK;
}

int orr_search_batch(orr_store* s, int32_t batch, const float* q, int32_t q_dim, const int32_t* n_terms,
                     const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
                     int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out) {
    const double t0 = now_ms();
    if (!s || batch < 0 || !out || !n_out || (batch > 0 && q_dim > 0 && !q)) { orr_set_error("orr_search_batch: bad argument"); return ORR_E_INVALID; }
    if (batch == 0) return ORR_OK;
    const int k = std::max(1, top_k);
    {   // one launch handles <= ORR_BATCH_MAX_QUERIES queries and <= BATCH_MAX_TERM_IDS query terms: slice larger batches
        int64_t terms_sum = 0;
        int32_t cut = batch;
        for (int32_t b = 0; b < batch; ++b) {
            terms_sum += n_terms ? std::max(0, n_terms[b]) : 0;
            if (b >= ORR_BATCH_MAX_QUERIES || terms_sum > BATCH_MAX_TERM_IDS) { cut = b; break; }
        }
        if (cut < batch && cut > 0) {
            for (int32_t b0 = 0; b0 < batch;) {
                int64_t sum = 0;
                int32_t b1 = b0;
                while (b1 < batch && b1 - b0 < ORR_BATCH_MAX_QUERIES) {
                    const int64_t nt = n_terms ? std::max(0, n_terms[b1]) : 0;
                    if (b1 > b0 && sum + nt > BATCH_MAX_TERM_IDS) break;
                    sum += nt; ++b1;
                }
                int rc = orr_search_batch(s, b1 - b0, q ? q + (int64_t)b0 * q_dim : nullptr, q_dim, n_terms ? n_terms + b0 : nullptr,
                                          probe_hash, probe_term, probe_offsets ? probe_offsets + b0 : nullptr, now_ticks, top_k,
                                          out + (int64_t)b0 * k, n_out + b0);
                if (rc != ORR_OK) return rc;
                b0 = b1;
            }
            g_timing.wall_ms = (float)(now_ms() - t0);
            return ORR_OK;
        }
    }
    // the tcgen05 path takes: a query embedding of the store's width, dim % 64 == 0, k <= 128,
    // <= 16 terms per query given as identity probes; anything else runs query by query
    bool gemm_ok = (q_dim == s->cfg.dim) && (s->cfg.dim % 64 == 0) && k <= 128 && batch >= 8 && s->live_rows > 0 &&
                   getenv("ORR_BATCH_LOOP") == nullptr;
    if (gemm_ok && n_terms) {
        for (int32_t b = 0; b < batch && gemm_ok; ++b) {
            const int32_t nt = n_terms[b];
            if (nt < 0 || nt > ORR_BATCH_TERMS) gemm_ok = false;
            if (nt > 0) {
                if (!probe_offsets || !probe_hash) gemm_ok = false;
                else if ((int32_t)(probe_offsets[b + 1] - probe_offsets[b]) != nt) gemm_ok = false;
                else if (probe_term) for (int32_t t = 0; t < nt; ++t) if (probe_term[probe_offsets[b] + t] != t) gemm_ok = false;
            }
        }
    }
    std::vector<int32_t> redo;
    if (gemm_ok) {
        std::shared_lock<std::shared_mutex> lock(s->mu);
        ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
        memset(&g_timing, 0, sizeof g_timing);
        // auto: screen with one bf16 pass (3x less tensor work, deeper candidate lists); queries whose selection the
        // bound check cannot prove go through bf16x3.  A screen that fails for most of a batch (scores packed closer
        // than the bf16 error bound) is skipped for the next 16 batches.
        const bool automatic = s->batch_passes == 0;
        int passes = automatic ? 1 : s->batch_passes;
        if (automatic && s->batch_hold.load() > 0) { passes = 3; s->batch_hold.fetch_sub(1); }
        int rc = batch_gemm_path(s, batch, q, n_terms, probe_hash, probe_offsets, now_ticks, top_k, out, n_out, passes, &redo);
        if (rc != ORR_OK) return rc;
        if (automatic && passes == 1 && redo.size() * 2 > (size_t)batch) s->batch_hold = 16;
        if (automatic && passes == 1 && redo.size() >= 8) {
            // queries the bf16 screen could not prove safe go through the split-precision GEMM as a
            // smaller batch before anything falls back to the per-query path
            const orr_timing first = g_timing;
            const int32_t nb = (int32_t)redo.size();
            std::vector<float> q2((size_t)nb * q_dim);
            std::vector<int32_t> nt2((size_t)nb, 0);
            std::vector<uint32_t> off2((size_t)nb + 1, 0u);
            std::vector<uint64_t> ph2;
            for (int32_t i = 0; i < nb; ++i) {
                const int32_t b = redo[(size_t)i];
                memcpy(q2.data() + (size_t)i * q_dim, q + (int64_t)b * q_dim, sizeof(float) * (size_t)q_dim);
                if (n_terms && n_terms[b] > 0) {
                    nt2[(size_t)i] = n_terms[b];
                    ph2.insert(ph2.end(), probe_hash + probe_offsets[b], probe_hash + probe_offsets[b + 1]);
                }
                off2[(size_t)i + 1] = (uint32_t)ph2.size();
            }
            std::vector<orr_hit> out2((size_t)nb * k);
            std::vector<int32_t> n2((size_t)nb, 0), redo2;
            rc = batch_gemm_path(s, nb, q2.data(), nt2.data(), ph2.empty() ? nullptr : ph2.data(), off2.data(), now_ticks,
                                 top_k, out2.data(), n2.data(), 3, &redo2);
            if (rc == ORR_E_OOM) {
                // no room for the second bf16 plane: the unproven queries take the per-query path below instead
                cudaGetLastError();
                redo2.clear();
                for (int32_t i = 0; i < nb; ++i) redo2.push_back(i);
            } else if (rc != ORR_OK) {
                return rc;
            }
            std::vector<int32_t> still;
            size_t r2 = 0;
            for (int32_t i = 0; i < nb; ++i) {
                const int32_t b = redo[(size_t)i];
                if (r2 < redo2.size() && redo2[r2] == i) { still.push_back(b); ++r2; continue; }
                memcpy(out + (int64_t)b * k, out2.data() + (size_t)i * k, sizeof(orr_hit) * (size_t)k);
                n_out[b] = n2[(size_t)i];
            }
            redo.swap(still);
            g_timing.scan_ms += first.scan_ms; g_timing.finalize_ms += first.finalize_ms;
            g_timing.total_device_ms += first.total_device_ms;
            g_timing.path = ORR_PATH_BATCH | ORR_PATH_ESCALATED;
        }
    } else {
        for (int32_t b = 0; b < batch; ++b) redo.push_back(b);
    }
    const orr_timing batch_timing = g_timing;
    for (int32_t b : redo) {                       // per-query path (exact escalation included)
        const uint32_t p0 = probe_offsets ? probe_offsets[b] : 0, p1 = probe_offsets ? probe_offsets[b + 1] : 0;
        int rc = orr_search(s, q ? q + (int64_t)b * q_dim : nullptr, q_dim, n_terms ? n_terms[b] : 0,
                            probe_hash ? probe_hash + p0 : nullptr, probe_term ? probe_term + p0 : nullptr,
                            (int32_t)(p1 - p0), now_ticks, top_k, 0, out + (int64_t)b * k, n_out + b);
        if (rc != ORR_OK) return rc;
    }
    if (gemm_ok) {
        g_timing = batch_timing;
        g_timing.n_survivors = (int32_t)redo.size() | (batch_timing.n_survivors << 16);   // low half: queries re-run singly
    }
    g_timing.wall_ms = (float)(now_ms() - t0);
    return ORR_OK;
}

// orr_search_batch with the answers left in HBM: the row-sharded form feeds them straight into the all-gather.  Takes what the
// tcgen05 path takes in ONE launch and returns ORR_E_UNSUPPORTED — nothing usable written — when the batch has to go
// through orr_search_batch instead (shape outside the path, or a query whose selection could not be proven).
int orr_search_batch_device(orr_store* s, int32_t batch, const float* q, int32_t q_dim, const int32_t* n_terms,
                            const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
                            int64_t now_ticks, int32_t top_k, orr_hit* out_dev, int32_t* n_out_dev) {
    const double t0 = now_ms();
    if (!s || batch <= 0 || !out_dev || !n_out_dev || !q) { orr_set_error("orr_search_batch_device: bad argument"); return ORR_E_INVALID; }
    const int k = std::max(1, top_k);
    bool ok = (q_dim == s->cfg.dim) && (s->cfg.dim % 64 == 0) && k <= 128 && batch >= 8 && batch <= ORR_BATCH_MAX_QUERIES;
    int64_t terms_sum = 0;
    if (ok && n_terms) {
        for (int32_t b = 0; b < batch && ok; ++b) {
            const int32_t nt = n_terms[b];
            if (nt < 0 || nt > ORR_BATCH_TERMS) ok = false;
            if (nt > 0) {
                if (!probe_offsets || !probe_hash) ok = false;
                else if ((int32_t)(probe_offsets[b + 1] - probe_offsets[b]) != nt) ok = false;
                else if (probe_term) for (int32_t t = 0; t < nt; ++t) if (probe_term[probe_offsets[b] + t] != t) ok = false;
            }
            terms_sum += std::max(0, nt);
        }
    }
    if (!ok || terms_sum > BATCH_MAX_TERM_IDS) { orr_set_error("orr_search_batch_device: batch outside the single-launch tcgen05 path"); return ORR_E_UNSUPPORTED; }
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    if (s->live_rows <= 0) { orr_set_error("orr_search_batch_device: empty store"); return ORR_E_UNSUPPORTED; }
    memset(&g_timing, 0, sizeof g_timing);
    int passes = s->batch_passes == 0 ? 1 : s->batch_passes;
    if (s->batch_passes == 0 && s->batch_hold.load() > 0) { passes = 3; s->batch_hold.fetch_sub(1); }   // as orr_search_batch's auto mode
    std::vector<int32_t> redo;
    int rc = batch_gemm_path(s, batch, q, n_terms, probe_hash, probe_offsets, now_ticks, top_k, nullptr, nullptr, passes, &redo,
                             out_dev, n_out_dev);
    if (rc != ORR_OK) return rc;
    g_timing.wall_ms = (float)(now_ms() - t0);
    if (s->batch_passes == 0 && passes == 1 && redo.size() * 2 > (size_t)batch) s->batch_hold = 16;   // the bf16 screen fails on this corpus: skip it for a while
    if (!redo.empty()) {
        orr_set_error("orr_search_batch_device: %d of %d queries not proven by the screen; run the batch through orr_search_batch",
                      (int)redo.size(), batch);
        return ORR_E_UNSUPPORTED;
    }
    return ORR_OK;
}

// diagnostic: the raw fused GEMM scores (no keyword term) of every stride-th row tile
int orr_debug_batch_scores(orr_store* s, int32_t batch, const float* q, int32_t q_dim, int64_t now_ticks,
                           int32_t tile_stride, float* out, int64_t out_ld) {
    if (!s || !q || !out || batch < 1 || batch > ORR_BATCH_MAX_QUERIES || q_dim != s->cfg.dim || tile_stride < 1) { orr_set_error("orr_debug_batch_scores: bad argument"); return ORR_E_INVALID; }
    std::shared_lock<std::shared_mutex> lock(s->mu);
    ORR_CUDA_OK(cudaSetDevice(s->cfg.device));
    std::call_once(s->batch_once, [&] { s->batch.reset(new BatchState()); });
    BatchState* bs = s->batch.get();
    std::lock_guard<std::mutex> g(bs->mu);
    const int dim = s->cfg.dim;
    const int bp = (batch + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE * ORR_BATCH_TILE;
    int rc = batch_prepare(s, bs, bp, 1, s->batch_passes != 1);
    if (rc != ORR_OK) return rc;
    cudaStream_t st = bs->stream;
    const int64_t rows = s->rows_used, rows_pad = (rows + ORR_BATCH_TILE - 1) / ORR_BATCH_TILE * ORR_BATCH_TILE;
    const int64_t s_tiles = (rows_pad / ORR_BATCH_TILE + tile_stride - 1) / tile_stride, n_s = s_tiles * ORR_BATCH_TILE;
    if (out_ld < n_s) { orr_set_error("orr_debug_batch_scores: out_ld %lld < %lld", (long long)out_ld, (long long)n_s); return ORR_E_INVALID; }
    ORR_CUDA_OK(cudaMemcpyAsync(bs->q, q, sizeof(float) * (size_t)batch * dim, cudaMemcpyHostToDevice, st));
    if ((rc = orr_batch_prep_queries(bs->q, bs->qhi, bs->qmid, bs->qscale, bs->qbad, batch, bp, dim, st)) != ORR_OK) return rc;
    if ((rc = orr_batch_build_rowrec(s->d_ticks, bs->rowaux, rows, rows_pad, now_ticks, weights_of(s), st)) != ORR_OK) return rc;
    if ((size_t)bp * (size_t)n_s > bs->dense_elems) {
        cudaFree(bs->dense); bs->dense = nullptr; bs->dense_elems = 0;
        ORR_CUDA_OK(cudaMalloc(&bs->dense, (size_t)bp * (size_t)n_s * sizeof(float)));
        bs->dense_elems = (size_t)bp * (size_t)n_s;
    }
    OrrBatchGemm gm{};
    gm.qhi = bs->qhi; gm.qmid = bs->qmid; gm.ehi = bs->ehi; gm.emid = bs->emid; gm.rowaux = bs->rowaux;
    gm.qscale = bs->qscale; gm.thr = bs->thr; gm.cand = bs->cand; gm.cand_count = bs->cand_count; gm.cand_cap = BATCH_CAND_CAP;
    gm.rows = rows; gm.dim = dim; gm.batch_padded = bp; gm.sms = s->sms;
    gm.dense = bs->dense; gm.dense_ld = n_s; gm.tile_stride = tile_stride; gm.passes = s->batch_passes == 1 ? 1 : 3;
    if ((rc = orr_batch_launch_gemm(gm, st)) != ORR_OK) return rc;
    ORR_CUDA_OK(cudaMemcpy2DAsync(out, (size_t)out_ld * sizeof(float), bs->dense, (size_t)n_s * sizeof(float),
                                  (size_t)n_s * sizeof(float), (size_t)batch, cudaMemcpyDeviceToHost, st));
    ORR_CUDA_OK(cudaStreamSynchronize(st));
    return ORR_OK;
}

// host merge of per-shard lists with the reference tie chain (score desc / NaN last,
// ticks desc, row asc); lists are independent so a k-way pick is enough
int orr_merge_hits(const orr_hit* lists, const int32_t* list_len, int32_t n_lists, int32_t list_stride,
                   int32_t top_k, orr_hit* out, int32_t* n_out) {
    if (!lists || !list_len || !out || !n_out || n_lists < 0 || list_stride < 0) { orr_set_error("orr_merge_hits: bad argument"); return ORR_E_INVALID; }
    const int k = std::max(1, top_k);
    auto before = [](const orr_hit& x, const orr_hit& y) {
        const bool xn = std::isnan(x.score), yn = std::isnan(y.score);
        if (xn != yn) return yn;
        if (!xn && x.score != y.score) return x.score > y.score;
        if (x.created_ticks != y.created_ticks) return x.created_ticks > y.created_ticks;
        return x.row < y.row;
    };
    // every list is already in reference order: a k-way pick of the best head, k times
    std::vector<int32_t> head((size_t)n_lists, 0);
    int got = 0;
    while (got < k) {
        int best = -1;
        for (int32_t l = 0; l < n_lists; ++l) {
            if (head[(size_t)l] >= list_len[l] || head[(size_t)l] >= list_stride) continue;
            if (best < 0 || before(lists[(int64_t)l * list_stride + head[(size_t)l]], lists[(int64_t)best * list_stride + head[(size_t)best]])) best = l;
        }
        if (best < 0) break;
        out[got++] = lists[(int64_t)best * list_stride + head[(size_t)best]++];
    }
    *n_out = got;
    return ORR_OK;
}

int orr_merge_hits_device(int32_t device, const orr_hit* lists_dev, const int32_t* status_dev, int32_t n_lists,
                          int32_t list_stride, int32_t top_k, orr_hit* out_dev, int32_t* out_status_dev,
                          void* cuda_stream) {
    if (!lists_dev || !status_dev || !out_dev || !out_status_dev) { orr_set_error("orr_merge_hits_device: NULL argument"); return ORR_E_INVALID; }
    ORR_CUDA_OK(cudaSetDevice(device));
    return orr_launch_merge(lists_dev, status_dev, n_lists, list_stride, top_k, out_dev, out_status_dev,
                            (cudaStream_t)cuda_stream);
}

int orr_merge_hits_batch_device(int32_t device, const orr_hit* lists_dev, const int32_t* n_dev, int32_t n_lists, int32_t batch,
                                int32_t k, orr_hit* out_dev, int32_t* n_out_dev, void* cuda_stream) {
    if (!lists_dev || !n_dev || !out_dev || !n_out_dev) { orr_set_error("orr_merge_hits_batch_device: NULL argument"); return ORR_E_INVALID; }
    ORR_CUDA_OK(cudaSetDevice(device));
    return orr_launch_merge_batch(lists_dev, n_dev, n_lists, batch, k, out_dev, n_out_dev, (cudaStream_t)cuda_stream);
}

int orr_last_timing(orr_timing* out) {
    if (!out) return ORR_E_INVALID;
    *out = g_timing;
    return ORR_OK;
}

}  // extern "C"
