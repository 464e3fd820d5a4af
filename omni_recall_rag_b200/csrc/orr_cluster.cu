// orr_cluster.cu — one host process driving the GPUs of a box ("single process, one
// stream per device").  The .NET API is a single process, so its multi-GPU form is not one rank per GPU but
// one orr_cluster: N orr_store shards (one per device), N exchange buffers attached to each other
// (cudaDeviceEnablePeerAccess), and per query
//     for every device d:  q -> HBM(d) ; orr_search_device(d) ; orr_xchg_allgather_merge(d)      (all async)
//     device 0:            merged hits -> pinned host ; synchronise stream 0
// i.e. 3 kernels per device, all launched from the calling thread, no NCCL, no host thread per GPU.
// Built on the public C ABI only (orr_store_*, orr_search_device, orr_xchg_*).
#include <algorithm>
#include <cstring>
#include <thread>

#include "orr_internal.h"

namespace {
constexpr int CLUSTER_MAX = ORR_XCHG_MAX_WORLD;
constexpr int CLUSTER_ROW_SHIFT = 40;          // global row id = shard << 40 | local row
constexpr int CLUSTER_RING = 3;                // queries in flight in orr_cluster_search_many (< the exchange's 4 slots)
}

struct orr_cluster {
    int n = 0;
    int dim = 0;
    int max_k = 0;
    int devices[CLUSTER_MAX] = {};
    orr_store* store[CLUSTER_MAX] = {};
    orr_xchg* xchg[CLUSTER_MAX] = {};
    cudaStream_t stream[CLUSTER_MAX] = {};
    float* d_q[CLUSTER_MAX] = {};
    orr_hit* d_hits[CLUSTER_MAX] = {};
    int32_t* d_status[CLUSTER_MAX] = {};
    orr_hit* d_out[CLUSTER_MAX] = {};
    int32_t* d_out_status[CLUSTER_MAX] = {};
    float* h_q = nullptr;                       // pinned
    orr_hit* h_out = nullptr;                   // pinned
    int32_t* h_status = nullptr;                // pinned
    // orr_cluster_search_many: a ring of CLUSTER_RING queries in flight (allocated on first use)
    bool ring_ready = false;
    cudaStream_t xstream[CLUSTER_MAX] = {};     // the exchange + merge of query i runs here while stream[d] scans query i+1
    float* r_q[CLUSTER_MAX][CLUSTER_RING] = {};
    orr_hit* r_hits[CLUSTER_MAX][CLUSTER_RING] = {};
    int32_t* r_status[CLUSTER_MAX][CLUSTER_RING] = {};
    orr_hit* r_out[CLUSTER_MAX][CLUSTER_RING] = {};
    int32_t* r_out_status[CLUSTER_MAX][CLUSTER_RING] = {};
    cudaEvent_t ev_scan[CLUSTER_MAX][CLUSTER_RING] = {};
    cudaEvent_t ev_done[CLUSTER_MAX][CLUSTER_RING] = {};
    float* rh_q = nullptr;                      // pinned [RING][dim]
    orr_hit* rh_out = nullptr;                  // pinned [RING][max_k]
    int32_t* rh_status = nullptr;               // pinned [RING][2]
    std::unordered_map<uint64_t, int> doc_shard;
    std::mutex mu;                              // one search / mutation at a time per cluster
};

extern "C" {

void orr_cluster_destroy(orr_cluster* c) {
    if (!c) return;
    for (int d = 0; d < c->n; ++d) {
        cudaSetDevice(c->devices[d]);
        if (c->stream[d]) cudaStreamSynchronize(c->stream[d]);
    }
    for (int d = 0; d < c->n; ++d) {
        cudaSetDevice(c->devices[d]);
        if (c->xchg[d]) orr_xchg_destroy(c->xchg[d]);
        cudaFree(c->d_q[d]); cudaFree(c->d_hits[d]); cudaFree(c->d_status[d]); cudaFree(c->d_out[d]); cudaFree(c->d_out_status[d]);
        if (c->xstream[d]) { cudaStreamSynchronize(c->xstream[d]); cudaStreamDestroy(c->xstream[d]); }
        for (int r = 0; r < CLUSTER_RING; ++r) {
            cudaFree(c->r_q[d][r]); cudaFree(c->r_hits[d][r]); cudaFree(c->r_status[d][r]); cudaFree(c->r_out[d][r]); cudaFree(c->r_out_status[d][r]);
            if (c->ev_scan[d][r]) cudaEventDestroy(c->ev_scan[d][r]);
            if (c->ev_done[d][r]) cudaEventDestroy(c->ev_done[d][r]);
        }
        if (c->stream[d]) cudaStreamDestroy(c->stream[d]);
        if (c->store[d]) orr_store_destroy(c->store[d]);
    }
    cudaFreeHost(c->h_q); cudaFreeHost(c->h_out); cudaFreeHost(c->h_status);
    cudaFreeHost(c->rh_q); cudaFreeHost(c->rh_out); cudaFreeHost(c->rh_status);
    delete c;
}

int orr_cluster_create(const orr_config* cfg, const int32_t* devices, int32_t n_devices, int32_t max_top_k, orr_cluster** out) {
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > CLUSTER_MAX || max_top_k < 1 || max_top_k > ORR_FUSED_MAX_K ||
        (int64_t)n_devices * max_top_k > ORR_SORT_MAX) {
        orr_set_error("orr_cluster_create: bad argument (%d devices, max_top_k %d)", n_devices, max_top_k);
        return ORR_E_INVALID;
    }
    *out = nullptr;
    std::unique_ptr<orr_cluster, void (*)(orr_cluster*)> c(new orr_cluster(), orr_cluster_destroy);
    c->n = n_devices; c->dim = cfg->dim; c->max_k = max_top_k;
    for (int d = 0; d < n_devices; ++d) {
        c->devices[d] = devices[d];
        orr_config sc = *cfg;
        sc.device = devices[d];
        sc.row_base = (uint64_t)d << CLUSTER_ROW_SHIFT;
        int rc = orr_store_create(&sc, &c->store[d]);
        if (rc != ORR_OK) return rc;
        ORR_CUDA_OK(cudaSetDevice(devices[d]));
        ORR_CUDA_OK(cudaStreamCreateWithFlags(&c->stream[d], cudaStreamNonBlocking));
        ORR_CUDA_OK(cudaMalloc(&c->d_q[d], sizeof(float) * (size_t)cfg->dim));
        ORR_CUDA_OK(cudaMalloc(&c->d_hits[d], sizeof(orr_hit) * (size_t)max_top_k));
        ORR_CUDA_OK(cudaMalloc(&c->d_status[d], sizeof(int32_t) * 2));
        ORR_CUDA_OK(cudaMalloc(&c->d_out[d], sizeof(orr_hit) * (size_t)max_top_k));
        ORR_CUDA_OK(cudaMalloc(&c->d_out_status[d], sizeof(int32_t) * 2));
        rc = orr_xchg_create(devices[d], n_devices, d, max_top_k, &c->xchg[d]);
        if (rc != ORR_OK) return rc;
    }
    for (int d = 0; d < n_devices; ++d)
        for (int p = 0; p < n_devices; ++p)
            if (p != d) { int rc = orr_xchg_attach_peer(c->xchg[d], p, c->xchg[p]); if (rc != ORR_OK) return rc; }
    ORR_CUDA_OK(cudaMallocHost(&c->h_q, sizeof(float) * (size_t)cfg->dim));
    ORR_CUDA_OK(cudaMallocHost(&c->h_out, sizeof(orr_hit) * (size_t)max_top_k));
    ORR_CUDA_OK(cudaMallocHost(&c->h_status, sizeof(int32_t) * 2));
    *out = c.release();
    return ORR_OK;
}

int32_t orr_cluster_size(const orr_cluster* c) { return c ? c->n : 0; }
orr_store* orr_cluster_shard(orr_cluster* c, int32_t i) { return (c && i >= 0 && i < c->n) ? c->store[i] : nullptr; }

int64_t orr_cluster_count(const orr_cluster* c) {
    int64_t n = 0;
    if (c) for (int d = 0; d < c->n; ++d) n += orr_store_count(c->store[d]);
    return n;
}

// replace-by-document: a document lives on ONE shard (so replace/delete touch one GPU): the shard that already
// holds it, else the one with the fewest rows in use
int orr_cluster_upsert_document_chunks(orr_cluster* c, uint64_t doc_key, int32_t n, const float* emb, const uint8_t* has_emb,
                                       const int64_t* created_ticks, const uint64_t* term_hashes, const uint32_t* term_offsets,
                                       const char* text_lower_utf8, const uint64_t* text_offsets, uint64_t* out_rows) {
    if (!c) { orr_set_error("orr_cluster_upsert_document_chunks: NULL cluster"); return ORR_E_INVALID; }
    if (n == 0) return ORR_OK;
    std::lock_guard<std::mutex> g(c->mu);
    int shard = -1;
    auto it = c->doc_shard.find(doc_key);
    if (it != c->doc_shard.end()) shard = it->second;
    else {
        int64_t best = INT64_MAX;
        for (int d = 0; d < c->n; ++d) {
            const int64_t used = orr_store_rows_used(c->store[d]);
            if (used < best) { best = used; shard = d; }
        }
    }
    int rc = text_lower_utf8
        ? orr_store_upsert_document_chunks_text(c->store[shard], doc_key, n, emb, has_emb, created_ticks, term_hashes, term_offsets,
                                                text_lower_utf8, text_offsets, out_rows)
        : orr_store_upsert_document_chunks(c->store[shard], doc_key, n, emb, has_emb, created_ticks, term_hashes, term_offsets, out_rows);
    if (rc == ORR_OK) c->doc_shard[doc_key] = shard;
    return rc;
}

int orr_cluster_delete_document(orr_cluster* c, uint64_t doc_key) {
    if (!c) { orr_set_error("orr_cluster_delete_document: NULL cluster"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->doc_shard.find(doc_key);
    if (it == c->doc_shard.end()) return ORR_OK;
    const int rc = orr_store_delete_document(c->store[it->second], doc_key);
    if (rc == ORR_OK) c->doc_shard.erase(it);
    return rc;
}

// bench/test: rows [first_row + d * n_per_shard, ...) of the synthetic corpus on shard d
int orr_cluster_fill_synthetic(orr_cluster* c, const orr_synth_spec* spec, uint64_t first_row, int64_t n_per_shard) {
    if (!c || !spec) { orr_set_error("orr_cluster_fill_synthetic: NULL argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(c->mu);
    for (int d = 0; d < c->n; ++d) {
        const int rc = orr_store_fill_synthetic(c->store[d], spec, first_row + (uint64_t)d * (uint64_t)n_per_shard, n_per_shard);
        if (rc != ORR_OK) return rc;
    }
    return ORR_OK;
}

}  // extern "C"

// after a partial launch (device d failed after devices < d had already pushed their lists): let the launched exchange
// kernels run into their time-out, then give every exchange buffer a common sequence base again
static void cluster_recover(orr_cluster* c) {
    uint32_t top = 0;
    for (int d = 0; d < c->n; ++d) {
        cudaSetDevice(c->devices[d]);
        cudaStreamSynchronize(c->stream[d]);
        if (c->xstream[d]) cudaStreamSynchronize(c->xstream[d]);
        top = std::max(top, orr_xchg_sequence(c->xchg[d]));
    }
    cudaGetLastError();
    for (int d = 0; d < c->n; ++d) orr_xchg_resync(c->xchg[d], (top + ORR_XCHG_SLOTS) & 0x7fffffffu);
}

// runs fn(d) for every shard on its own host thread (the shards' searches are independent and each saturates its GPU)
template <class F>
static int for_each_shard_parallel(orr_cluster* c, F fn) {
    std::vector<int> rcs((size_t)c->n, ORR_OK);
    std::vector<std::string> errs((size_t)c->n);
    std::vector<std::thread> th;
    for (int d = 1; d < c->n; ++d)
        th.emplace_back([&, d] { rcs[(size_t)d] = fn(d); if (rcs[(size_t)d] != ORR_OK) errs[(size_t)d] = orr_last_error(); });
    rcs[0] = fn(0);
    for (auto& t : th) t.join();
    for (int d = 0; d < c->n; ++d)
        if (rcs[(size_t)d] != ORR_OK) { if (d > 0) orr_set_error("shard %d: %s", d, errs[(size_t)d].c_str()); return rcs[(size_t)d]; }
    return ORR_OK;
}

extern "C" {

// one query, the cluster's mutex held by the caller
static int cluster_search_locked(orr_cluster* c, const float* q, int32_t q_dim, int32_t n_terms, const uint64_t* probe_hash,
                                 const int32_t* probe_term, int32_t n_probes, int64_t now_ticks, int32_t top_k, orr_hit* out,
                                 int32_t* n_out) {
    *n_out = 0;
    const int k = std::max(1, top_k);
    bool fused = q_dim == c->dim && k <= c->max_k;
    if (fused) {
        memcpy(c->h_q, q, sizeof(float) * (size_t)q_dim);
        int rc = ORR_OK;
        for (int d = 0; d < c->n && rc == ORR_OK; ++d) {
            rc = [&]() -> int {
                ORR_CUDA_OK(cudaSetDevice(c->devices[d]));
                ORR_CUDA_OK(cudaMemcpyAsync(c->d_q[d], c->h_q, sizeof(float) * (size_t)q_dim, cudaMemcpyHostToDevice, c->stream[d]));
                int r = orr_search_device(c->store[d], c->d_q[d], q_dim, n_terms, probe_hash, probe_term, n_probes, now_ticks, top_k,
                                          c->d_hits[d], c->d_status[d], c->stream[d]);
                if (r != ORR_OK) return r;
                return orr_xchg_allgather_merge(c->xchg[d], c->d_hits[d], c->d_status[d], top_k, c->d_out[d], c->d_out_status[d], c->stream[d]);
            }();
        }
        if (rc != ORR_OK) {
            const std::string msg = orr_last_error();
            cluster_recover(c);                                    // the devices that did launch are waiting for the one that did not
            orr_set_error("%s", msg.c_str());
            return rc;
        }
        ORR_CUDA_OK(cudaSetDevice(c->devices[0]));
        ORR_CUDA_OK(cudaMemcpyAsync(c->h_status, c->d_out_status[0], sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, c->stream[0]));
        ORR_CUDA_OK(cudaMemcpyAsync(c->h_out, c->d_out[0], sizeof(orr_hit) * (size_t)k, cudaMemcpyDeviceToHost, c->stream[0]));
        ORR_CUDA_OK(cudaStreamSynchronize(c->stream[0]));
        if (c->h_status[1] & ORR_STATUS_XCHG_TIMEOUT) {
            cluster_recover(c);
            orr_set_error("orr_cluster_search: a shard never published its list");
            return ORR_E_CUDA;
        }
        if (c->h_status[1] == 0) {
            const int got = std::min(c->h_status[0], k);
            memcpy(out, c->h_out, sizeof(orr_hit) * (size_t)got);
            *n_out = got;
            return ORR_OK;
        }
        // a shard could not prove its fp32 selection: fall through to the per-shard host path (which escalates)
        for (int d = 0; d < c->n; ++d) { ORR_CUDA_OK(cudaSetDevice(c->devices[d])); ORR_CUDA_OK(cudaStreamSynchronize(c->stream[d])); }
    }
    // no query embedding (the reference's default configuration), a query of another width, k beyond the fused path:
    // every shard runs orr_search (-> exact path) on its own host thread, the lists are merged on the host
    std::vector<orr_hit> lists((size_t)c->n * k);
    std::vector<int32_t> lens((size_t)c->n, 0);
    const int rc = for_each_shard_parallel(c, [&](int d) {
        return orr_search(c->store[d], q, q_dim, n_terms, probe_hash, probe_term, n_probes, now_ticks, top_k, 0,
                          lists.data() + (size_t)d * k, &lens[(size_t)d]);
    });
    if (rc != ORR_OK) return rc;
    return orr_merge_hits(lists.data(), lens.data(), c->n, k, top_k, out, n_out);
}

static int cluster_ring_init(orr_cluster* c) {
    if (c->ring_ready) return ORR_OK;
    for (int d = 0; d < c->n; ++d) {
        ORR_CUDA_OK(cudaSetDevice(c->devices[d]));
        ORR_CUDA_OK(cudaStreamCreateWithFlags(&c->xstream[d], cudaStreamNonBlocking));
        for (int r = 0; r < CLUSTER_RING; ++r) {
            ORR_CUDA_OK(cudaMalloc(&c->r_q[d][r], sizeof(float) * (size_t)c->dim));
            ORR_CUDA_OK(cudaMalloc(&c->r_hits[d][r], sizeof(orr_hit) * (size_t)c->max_k));
            ORR_CUDA_OK(cudaMalloc(&c->r_status[d][r], sizeof(int32_t) * 2));
            ORR_CUDA_OK(cudaMalloc(&c->r_out[d][r], sizeof(orr_hit) * (size_t)c->max_k));
            ORR_CUDA_OK(cudaMalloc(&c->r_out_status[d][r], sizeof(int32_t) * 2));
            ORR_CUDA_OK(cudaEventCreateWithFlags(&c->ev_scan[d][r], cudaEventDisableTiming));
            ORR_CUDA_OK(cudaEventCreateWithFlags(&c->ev_done[d][r], cudaEventDisableTiming));
        }
    }
    ORR_CUDA_OK(cudaMallocHost(&c->rh_q, sizeof(float) * (size_t)c->dim * CLUSTER_RING));
    ORR_CUDA_OK(cudaMallocHost(&c->rh_out, sizeof(orr_hit) * (size_t)c->max_k * CLUSTER_RING));
    ORR_CUDA_OK(cudaMallocHost(&c->rh_status, sizeof(int32_t) * 2 * CLUSTER_RING));
    c->ring_ready = true;
    return ORR_OK;
}

int orr_cluster_search(orr_cluster* c, const float* q, int32_t q_dim, int32_t n_terms, const uint64_t* probe_hash,
                       const int32_t* probe_term, int32_t n_probes, int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out) {
    if (!c || !out || !n_out || q_dim < 0 || (q_dim > 0 && !q)) { orr_set_error("orr_cluster_search: bad argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(c->mu);
    return cluster_search_locked(c, q, q_dim, n_terms, probe_hash, probe_term, n_probes, now_ticks, top_k, out, n_out);
}

// A run of single queries, pipelined (the throughput form; the one-process counterpart of search_device_pipelined): the
// exchange + merge of query i runs on a side stream of every device while that device already scans query i + 1, and the
// merged hits of query i come down from device 0 behind it — CLUSTER_RING queries in flight, each with its own buffers.
// Every query is still the complete search: the hits are those of n_queries calls of orr_cluster_search.
int orr_cluster_search_many(orr_cluster* c, int32_t n_queries, const float* q, int32_t q_dim, const int32_t* n_terms,
                            const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
                            int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out) {
    if (!c || n_queries < 0 || !out || !n_out || q_dim < 0 || (n_queries > 0 && q_dim > 0 && !q)) {
        orr_set_error("orr_cluster_search_many: bad argument");
        return ORR_E_INVALID;
    }
    const int k = std::max(1, top_k);
    std::lock_guard<std::mutex> g(c->mu);
    auto terms_of = [&](int32_t i, int32_t* nt, const uint64_t** ph, const int32_t** pt, int32_t* np) {
        *nt = n_terms ? n_terms[i] : 0;
        const uint32_t p0 = probe_offsets ? probe_offsets[i] : 0u, p1 = probe_offsets ? probe_offsets[i + 1] : 0u;
        *ph = probe_hash ? probe_hash + p0 : nullptr;
        *pt = probe_term ? probe_term + p0 : nullptr;
        *np = (int32_t)(p1 - p0);
    };
    auto single = [&](int32_t i) {
        int32_t nt, np; const uint64_t* ph; const int32_t* pt;
        terms_of(i, &nt, &ph, &pt, &np);
        return cluster_search_locked(c, q ? q + (int64_t)i * q_dim : nullptr, q_dim, nt, ph, pt, np, now_ticks, top_k,
                                     out + (int64_t)i * k, n_out + i);
    };
    if (q_dim != c->dim || k > c->max_k || c->n < 1) {            // not the fused path: one query at a time
        for (int32_t i = 0; i < n_queries; ++i) { const int rc = single(i); if (rc != ORR_OK) return rc; }
        return ORR_OK;
    }
    int rc = cluster_ring_init(c);
    if (rc != ORR_OK) return rc;
    std::vector<int32_t> redo;
    bool timed_out = false;
    auto finish = [&](int32_t i) -> int {                          // query i's merged hits have reached pinned memory
        const int s = i % CLUSTER_RING;
        ORR_CUDA_OK(cudaEventSynchronize(c->ev_done[0][s]));
        const int32_t* st = c->rh_status + 2 * s;
        n_out[i] = 0;
        if (st[1] & ORR_STATUS_XCHG_TIMEOUT) { timed_out = true; return ORR_OK; }
        if (st[1] != 0) { redo.push_back(i); return ORR_OK; }      // a shard could not prove its fp32 selection
        const int got = std::min(st[0], k);
        memcpy(out + (int64_t)i * k, c->rh_out + (size_t)s * c->max_k, sizeof(orr_hit) * (size_t)got);
        n_out[i] = got;
        return ORR_OK;
    };
    for (int32_t i = 0; i < n_queries && rc == ORR_OK; ++i) {
        const int s = i % CLUSTER_RING;
        if (i >= CLUSTER_RING) { rc = finish(i - CLUSTER_RING); if (rc != ORR_OK) break; }
        int32_t nt, np; const uint64_t* ph; const int32_t* pt;
        terms_of(i, &nt, &ph, &pt, &np);
        float* hq = c->rh_q + (size_t)s * c->dim;
        memcpy(hq, q + (int64_t)i * q_dim, sizeof(float) * (size_t)q_dim);
        for (int d = 0; d < c->n && rc == ORR_OK; ++d) {
            rc = [&]() -> int {
                ORR_CUDA_OK(cudaSetDevice(c->devices[d]));
                // slot s of this device is free once the exchange of query i - RING has read its local list
                if (i >= CLUSTER_RING) ORR_CUDA_OK(cudaStreamWaitEvent(c->stream[d], c->ev_done[d][s], 0));
                ORR_CUDA_OK(cudaMemcpyAsync(c->r_q[d][s], hq, sizeof(float) * (size_t)q_dim, cudaMemcpyHostToDevice, c->stream[d]));
                int r = orr_search_device(c->store[d], c->r_q[d][s], q_dim, nt, ph, pt, np, now_ticks, top_k, c->r_hits[d][s],
                                          c->r_status[d][s], c->stream[d]);
                if (r != ORR_OK) return r;
                ORR_CUDA_OK(cudaEventRecord(c->ev_scan[d][s], c->stream[d]));
                ORR_CUDA_OK(cudaStreamWaitEvent(c->xstream[d], c->ev_scan[d][s], 0));
                r = orr_xchg_allgather_merge(c->xchg[d], c->r_hits[d][s], c->r_status[d][s], top_k, c->r_out[d][s], c->r_out_status[d][s],
                                             c->xstream[d]);
                if (r != ORR_OK) return r;
                if (d == 0) {
                    ORR_CUDA_OK(cudaMemcpyAsync(c->rh_status + 2 * s, c->r_out_status[0][s], sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, c->xstream[0]));
                    ORR_CUDA_OK(cudaMemcpyAsync(c->rh_out + (size_t)s * c->max_k, c->r_out[0][s], sizeof(orr_hit) * (size_t)k,
                                                cudaMemcpyDeviceToHost, c->xstream[0]));
                }
                ORR_CUDA_OK(cudaEventRecord(c->ev_done[d][s], c->xstream[d]));
                return ORR_OK;
            }();
        }
    }
    if (rc != ORR_OK) {
        const std::string msg = orr_last_error();
        cluster_recover(c);                                        // devices that did launch wait for the one that did not
        orr_set_error("%s", msg.c_str());
        return rc;
    }
    for (int32_t i = std::max(0, n_queries - CLUSTER_RING); i < n_queries; ++i) { rc = finish(i); if (rc != ORR_OK) return rc; }
    for (int d = 0; d < c->n; ++d) {                               // the other devices' side streams end with the same queries
        ORR_CUDA_OK(cudaSetDevice(c->devices[d]));
        ORR_CUDA_OK(cudaStreamSynchronize(c->xstream[d]));
    }
    if (timed_out) {
        cluster_recover(c);
        orr_set_error("orr_cluster_search_many: a shard never published its list");
        return ORR_E_CUDA;
    }
    for (int32_t i : redo) { rc = single(i); if (rc != ORR_OK) return rc; }
    return ORR_OK;
}

// Batched queries over the shards: every shard runs orr_search_batch (tcgen05 contraction + exact re-rank, with its own
// cascade / re-run handling) on its own host thread — the hits come down once per shard, in parallel over the GPUs' own
// PCIe links — and each query's N sorted lists are merged by a k-way pick under the reference tie chain, the queries
// split over the same threads.
int orr_cluster_search_batch(orr_cluster* c, int32_t batch, const float* q, int32_t q_dim, const int32_t* n_terms,
                             const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
                             int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out) {
    if (!c || batch < 0 || !out || !n_out || (batch > 0 && q_dim > 0 && !q)) { orr_set_error("orr_cluster_search_batch: bad argument"); return ORR_E_INVALID; }
    if (batch == 0) return ORR_OK;
    const int k = std::max(1, top_k);
    std::lock_guard<std::mutex> g(c->mu);
    std::vector<orr_hit> lists((size_t)c->n * (size_t)batch * k);
    std::vector<int32_t> lens((size_t)c->n * (size_t)batch, 0);
    int rc = for_each_shard_parallel(c, [&](int d) {
        return orr_search_batch(c->store[d], batch, q, q_dim, n_terms, probe_hash, probe_term, probe_offsets, now_ticks, top_k,
                                lists.data() + (size_t)d * batch * k, lens.data() + (size_t)d * batch);
    });
    if (rc != ORR_OK) return rc;
    return for_each_shard_parallel(c, [&](int d) {
        std::vector<orr_hit> mine((size_t)c->n * k);
        std::vector<int32_t> ml((size_t)c->n);
        for (int32_t b = d; b < batch; b += c->n) {
            for (int l = 0; l < c->n; ++l) {
                ml[(size_t)l] = lens[(size_t)l * batch + b];
                memcpy(mine.data() + (size_t)l * k, lists.data() + ((size_t)l * batch + b) * k, sizeof(orr_hit) * (size_t)k);
            }
            const int r = orr_merge_hits(mine.data(), ml.data(), c->n, k, top_k, out + (size_t)b * k, n_out + b);
            if (r != ORR_OK) return r;
        }
        return (int)ORR_OK;
    });
}

}  // extern "C"
