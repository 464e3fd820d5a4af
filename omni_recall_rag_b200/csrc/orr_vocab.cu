// orr_vocab.cu — the live vocabulary and the substring expansion of query terms (the host half of
// `contentLower.Contains(term, Ordinal)`, RecallSearchService.cs:110-111).
//
// The hashed keyword index stores, per chunk, the hashes of its distinct lower-cased white-space tokens.  The
// reference's predicate is SUBSTRING containment; because query terms hold no white space and chunk text is
// words joined by single spaces (SlidingWindowTextChunker.cs:29), "content contains t" == "some word of the chunk
// contains t".  A query term therefore expands into the vocabulary words that contain it, and the scan probes the
// chunk's term set for any of those words' hashes (orr_search's probe_term).
//
// Round 1 did that expansion with a linear scan over a host dictionary: O(V) string searches per term and request
// (~100x the GPU's scan time at a 1M-word vocabulary).  Here the vocabulary's bytes live in HBM next to the store:
//   host   word -> id map, per-word reference count (live chunks holding it) and hash; arena of the words' bytes
//   HBM    the same arena + offsets, appended lazily (only the tail added since the last expansion is uploaded)
//   orr_vocab_match_kernel   one thread per word, the query's <= 64 terms in shared memory: emits (word id, term)
//                            for every word that contains the term (8 MB at 1M words: microseconds)
// The host filters the pairs by reference count (words whose last chunk was deleted stay in the arena until it is
// rebuilt) and maps ids to hashes.  Expansions are cached per term until the vocabulary changes.
#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>

#include "orr_internal.h"

namespace {

constexpr int VOCAB_MAX_PAIRS = 16384;       // (word, term) matches one expansion can report
constexpr int VOCAB_THREADS = 256;

struct VocabTerms {                          // the query terms as the kernel sees them
    int32_t  n_terms;
    int32_t  min_len;
    uint16_t off[ORR_MAX_QUERY_TERMS + 1];
    uint8_t  bytes[ORR_TEXT_TERMS_BYTES];
};

__global__ void __launch_bounds__(VOCAB_THREADS) orr_vocab_match_kernel(const uint8_t* bytes, const uint32_t* off, uint32_t n_words,
                                                                        const VocabTerms* terms_g, uint2* pairs, uint32_t* count,
                                                                        uint32_t cap) {
    __shared__ VocabTerms tt;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(terms_g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tt);
        for (int i = threadIdx.x; i < (int)(sizeof(VocabTerms) / 4); i += VOCAB_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    for (uint32_t w = blockIdx.x * VOCAB_THREADS + threadIdx.x; w < n_words; w += gridDim.x * VOCAB_THREADS) {
        const uint32_t b = off[w], len = off[w + 1] - b;
        if ((int)len < tt.min_len) continue;
        const uint8_t* p = bytes + b;
        for (int t = 0; t < tt.n_terms; ++t) {
            const int tl = tt.off[t + 1] - tt.off[t];
            if (tl > (int)len) continue;
            const uint8_t* pat = tt.bytes + tt.off[t];
            const uint8_t c0 = pat[0];
            bool hit = false;
            for (int s = 0; s + tl <= (int)len && !hit; ++s) {
                if (p[s] != c0) continue;
                int j = 1;
                while (j < tl && p[s + j] == pat[j]) ++j;
                hit = (j == tl);
            }
            if (hit) {
                const uint32_t slot = atomicAdd(count, 1u);
                if (slot < cap) pairs[slot] = make_uint2(w, (uint32_t)t);
            }
        }
    }
}

}  // namespace

struct OrrVocab {
    int device = 0;
    std::unordered_map<std::string, uint32_t> index;
    std::vector<uint32_t> refs;
    std::vector<uint64_t> hash;
    std::string bytes;
    std::vector<uint32_t> off{0u};
    uint64_t version = 0;                    // bumps whenever a word appears or its count crosses zero
    uint64_t live_words = 0;
    // device mirror
    uint8_t* d_bytes = nullptr; size_t d_bytes_cap = 0, d_bytes_used = 0;
    uint32_t* d_off = nullptr; size_t d_off_cap = 0; uint32_t d_words = 0;
    VocabTerms* d_terms = nullptr; VocabTerms* h_terms = nullptr;
    uint2* d_pairs = nullptr; uint2* h_pairs = nullptr;
    uint32_t* d_count = nullptr; uint32_t* h_count = nullptr;
    cudaStream_t stream = nullptr;
    // expansion cache: term -> ids of the live words containing it, valid for `cache_version`
    std::unordered_map<std::string, std::vector<uint32_t>> cache;
    uint64_t cache_version = ~0ull;
};

OrrVocab* orr_vocab_new(int device) {
    OrrVocab* v = new OrrVocab();
    v->device = device;
    return v;
}

void orr_vocab_free(OrrVocab* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree(v->d_bytes); cudaFree(v->d_off); cudaFree(v->d_terms); cudaFree(v->d_pairs); cudaFree(v->d_count);
    cudaFreeHost(v->h_terms); cudaFreeHost(v->h_pairs); cudaFreeHost(v->h_count);
    delete v;
}

uint64_t orr_vocab_live_words(const OrrVocab* v) { return v->live_words; }
uint64_t orr_vocab_words(const OrrVocab* v) { return v->refs.size(); }

uint32_t orr_vocab_add(OrrVocab* v, const char* word, size_t len, uint32_t count) {
    std::string key(word, len);
    auto it = v->index.find(key);
    uint32_t id;
    if (it == v->index.end()) {
        id = (uint32_t)v->refs.size();
        v->index.emplace(std::move(key), id);
        v->refs.push_back(0u);
        v->hash.push_back(orr_hash_bytes(word, (int64_t)len));
        v->bytes.append(word, len);
        v->off.push_back((uint32_t)v->bytes.size());
    } else {
        id = it->second;
    }
    if (v->refs[id] == 0u && count > 0u) { v->live_words++; v->version++; }
    v->refs[id] += count;
    return id;
}

// read-only lookup (safe from several threads while nobody adds words): the word's id, or 0xffffffff
uint32_t orr_vocab_find(const OrrVocab* v, const char* word, size_t len) {
    auto it = v->index.find(std::string(word, len));
    return it == v->index.end() ? 0xffffffffu : it->second;
}

// `count` more chunks hold word `id`
void orr_vocab_addref(OrrVocab* v, uint32_t id, uint32_t count) {
    if (id >= v->refs.size() || count == 0u) return;
    if (v->refs[id] == 0u) { v->live_words++; v->version++; }
    v->refs[id] += count;
}

void orr_vocab_release(OrrVocab* v, uint32_t id) {
    if (id >= v->refs.size() || v->refs[id] == 0u) return;
    if (--v->refs[id] == 0u) { v->live_words--; v->version++; }
}

void orr_vocab_clear(OrrVocab* v) {
    v->index.clear(); v->refs.clear(); v->hash.clear(); v->bytes.clear(); v->off.assign(1, 0u);
    v->live_words = 0; v->version++; v->d_words = 0; v->d_bytes_used = 0; v->cache.clear();
}

// serialisation for the snapshot: [n_words u64][bytes_len u64][refs u32 x n][off u32 x (n+1)][bytes]
void orr_vocab_serialize(const OrrVocab* v, std::vector<uint8_t>* out) {
    const uint64_t n = v->refs.size(), bl = v->bytes.size();
    out->resize(16 + 4 * n + 4 * (n + 1) + bl);
    uint8_t* p = out->data();
    memcpy(p, &n, 8); memcpy(p + 8, &bl, 8); p += 16;
    memcpy(p, v->refs.data(), 4 * n); p += 4 * n;
    memcpy(p, v->off.data(), 4 * (n + 1)); p += 4 * (n + 1);
    memcpy(p, v->bytes.data(), bl);
}
int orr_vocab_deserialize(OrrVocab* v, const uint8_t* p, size_t len) {
    if (len < 16) return ORR_E_INVALID;
    uint64_t n = 0, bl = 0;
    memcpy(&n, p, 8); memcpy(&bl, p + 8, 8);
    if (n > 0xfffffff0ull || bl > 0xfffffff0ull || 16 + 4 * n + 4 * (n + 1) + bl != len) return ORR_E_INVALID;
    const uint32_t* refs = reinterpret_cast<const uint32_t*>(p + 16);
    const uint32_t* off = reinterpret_cast<const uint32_t*>(p + 16 + 4 * n);
    const char* bytes = reinterpret_cast<const char*>(p + 16 + 4 * n + 4 * (n + 1));
    if (off[0] != 0 || off[n] != bl) return ORR_E_INVALID;
    for (uint64_t i = 0; i < n; ++i) if (off[i + 1] < off[i]) return ORR_E_INVALID;
    orr_vocab_clear(v);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t id = orr_vocab_add(v, bytes + off[i], off[i + 1] - off[i], refs[i]);
        if (id != (uint32_t)i) return ORR_E_INVALID;                          // duplicate word in the file
    }
    return ORR_OK;
}

static int vocab_sync_device(OrrVocab* v) {
    if (!v->stream) {
        ORR_CUDA_OK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
        ORR_CUDA_OK(cudaMalloc(&v->d_terms, sizeof(VocabTerms)));
        ORR_CUDA_OK(cudaMallocHost(&v->h_terms, sizeof(VocabTerms)));
        ORR_CUDA_OK(cudaMalloc(&v->d_pairs, sizeof(uint2) * VOCAB_MAX_PAIRS));
        ORR_CUDA_OK(cudaMallocHost(&v->h_pairs, sizeof(uint2) * VOCAB_MAX_PAIRS));
        ORR_CUDA_OK(cudaMalloc(&v->d_count, sizeof(uint32_t)));
        ORR_CUDA_OK(cudaMallocHost(&v->h_count, sizeof(uint32_t)));
    }
    const uint32_t n = (uint32_t)v->refs.size();
    if (n == v->d_words) return ORR_OK;
    if (v->bytes.size() > v->d_bytes_cap) {                                   // grow: new arena, old bytes copied on the device
        const size_t cap = std::max<size_t>(2 * v->bytes.size(), (size_t)1 << 20);
        uint8_t* nb = nullptr;
        ORR_CUDA_OK(cudaMalloc(&nb, cap));
        if (v->d_bytes_used) ORR_CUDA_OK(cudaMemcpyAsync(nb, v->d_bytes, v->d_bytes_used, cudaMemcpyDeviceToDevice, v->stream));
        ORR_CUDA_OK(cudaStreamSynchronize(v->stream));
        cudaFree(v->d_bytes);
        v->d_bytes = nb; v->d_bytes_cap = cap;
    }
    if ((size_t)n + 1 > v->d_off_cap) {
        const size_t cap = std::max<size_t>(2 * ((size_t)n + 1), (size_t)1 << 16);
        uint32_t* no = nullptr;
        ORR_CUDA_OK(cudaMalloc(&no, cap * sizeof(uint32_t)));
        if (v->d_words) ORR_CUDA_OK(cudaMemcpyAsync(no, v->d_off, ((size_t)v->d_words + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, v->stream));
        ORR_CUDA_OK(cudaStreamSynchronize(v->stream));
        cudaFree(v->d_off);
        v->d_off = no; v->d_off_cap = cap;
    }
    ORR_CUDA_OK(cudaMemcpyAsync(v->d_bytes + v->d_bytes_used, v->bytes.data() + v->d_bytes_used, v->bytes.size() - v->d_bytes_used,
                                cudaMemcpyHostToDevice, v->stream));
    ORR_CUDA_OK(cudaMemcpyAsync(v->d_off + v->d_words, v->off.data() + v->d_words, ((size_t)n - v->d_words + 1) * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, v->stream));
    ORR_CUDA_OK(cudaStreamSynchronize(v->stream));                            // the host vectors may reallocate after this call
    v->d_words = n; v->d_bytes_used = v->bytes.size();
    return ORR_OK;
}

// Expands `terms` (lower-cased, A-2 filtered) into (hash, term) probes over the live vocabulary.  *n_probes may exceed
// `cap` (nothing beyond cap is written): the caller then evaluates the query in text mode.  Caller holds the vocab lock.
int orr_vocab_expand(OrrVocab* v, const std::vector<std::string>& terms, uint64_t* probe_hash, int32_t* probe_term, int32_t cap,
                     int32_t* n_probes) {
    *n_probes = 0;
    const int nt = (int)terms.size();
    if (nt == 0 || v->refs.empty()) return ORR_OK;
    if (nt > ORR_MAX_QUERY_TERMS) { orr_set_error("vocabulary expansion: %d terms > %d", nt, ORR_MAX_QUERY_TERMS); return ORR_E_UNSUPPORTED; }
    if (v->cache_version != v->version || v->cache.size() > 65536) { v->cache.clear(); v->cache_version = v->version; }
    std::vector<const std::vector<uint32_t>*> found((size_t)nt, nullptr);
    std::vector<int> missing;
    size_t total_bytes = 0;
    for (int t = 0; t < nt; ++t) {
        auto it = v->cache.find(terms[(size_t)t]);
        if (it != v->cache.end()) found[(size_t)t] = &it->second;
        else { missing.push_back(t); total_bytes += terms[(size_t)t].size(); }
    }
    if (!missing.empty()) {
        for (int t : missing)
            if (terms[(size_t)t].empty() || terms[(size_t)t].size() > (size_t)ORR_TEXT_MAX_TERM_BYTES || total_bytes > (size_t)ORR_TEXT_TERMS_BYTES) {
                orr_set_error("vocabulary expansion: term too long (limit %d bytes, %d in total)", ORR_TEXT_MAX_TERM_BYTES, ORR_TEXT_TERMS_BYTES);
                return ORR_E_UNSUPPORTED;
            }
        ORR_CUDA_OK(cudaSetDevice(v->device));
        int rc = vocab_sync_device(v);
        if (rc != ORR_OK) return rc;
        VocabTerms* ht = v->h_terms;
        ht->n_terms = (int32_t)missing.size();
        ht->min_len = 1 << 30;
        uint32_t o = 0;
        for (size_t i = 0; i < missing.size(); ++i) {
            const std::string& s = terms[(size_t)missing[i]];
            ht->off[i] = (uint16_t)o;
            memcpy(ht->bytes + o, s.data(), s.size());
            o += (uint32_t)s.size();
            ht->min_len = std::min<int32_t>(ht->min_len, (int32_t)s.size());
        }
        ht->off[missing.size()] = (uint16_t)o;
        *v->h_count = 0u;
        ORR_CUDA_OK(cudaMemcpyAsync(v->d_terms, ht, sizeof(VocabTerms), cudaMemcpyHostToDevice, v->stream));
        ORR_CUDA_OK(cudaMemsetAsync(v->d_count, 0, sizeof(uint32_t), v->stream));
        const int grid = (int)std::min<uint32_t>((v->d_words + VOCAB_THREADS - 1) / VOCAB_THREADS, 148u * 8u);
        orr_vocab_match_kernel<<<std::max(1, grid), VOCAB_THREADS, 0, v->stream>>>(v->d_bytes, v->d_off, v->d_words, v->d_terms, v->d_pairs,
                                                                                  v->d_count, (uint32_t)VOCAB_MAX_PAIRS);
        ORR_CUDA_OK(cudaGetLastError());
        ORR_CUDA_OK(cudaMemcpyAsync(v->h_count, v->d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, v->stream));
        ORR_CUDA_OK(cudaStreamSynchronize(v->stream));
        const uint32_t got = *v->h_count;
        if (got > (uint32_t)VOCAB_MAX_PAIRS) {                               // a term inside tens of thousands of words: text mode's job
            *n_probes = (int32_t)std::min<uint32_t>(got, 0x7fffffffu);
            return ORR_OK;
        }
        if (got) {
            ORR_CUDA_OK(cudaMemcpyAsync(v->h_pairs, v->d_pairs, sizeof(uint2) * got, cudaMemcpyDeviceToHost, v->stream));
            ORR_CUDA_OK(cudaStreamSynchronize(v->stream));
        }
        std::vector<std::vector<uint32_t>> lists(missing.size());
        for (uint32_t i = 0; i < got; ++i) {
            const uint2 pr = v->h_pairs[i];
            if (pr.x < v->refs.size() && v->refs[pr.x] > 0u) lists[pr.y].push_back(pr.x);
        }
        for (size_t i = 0; i < missing.size(); ++i) {
            std::sort(lists[i].begin(), lists[i].end());                      // atomics order is arbitrary: make probes deterministic
            auto ins = v->cache.emplace(terms[(size_t)missing[i]], std::move(lists[i]));
            found[(size_t)missing[i]] = &ins.first->second;
        }
    }
    int64_t n = 0;
    for (int t = 0; t < nt; ++t) {
        for (uint32_t id : *found[(size_t)t]) {
            if (n < cap) { probe_hash[n] = v->hash[id]; probe_term[n] = t; }
            ++n;
        }
    }
    *n_probes = (int32_t)std::min<int64_t>(n, 0x7fffffff);
    return ORR_OK;
}
