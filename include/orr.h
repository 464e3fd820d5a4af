/*
 * orr.h — C ABI of liborr.so, the B200-native hybrid recall scorer.
 *
 * The reference (fchchen/omni-recall-rag, .NET 10) has NO native/FFI boundary for
 * this path: its seams are two DI-registered C# interfaces,
 *   IRecallSearchService.SearchAsync      src/OmniRecall.Api/Services/RecallSearchService.cs:6-9
 *   IIngestionStore (8 async methods)     src/OmniRecall.Api/Services/IIngestionStore.cs:5-17
 * A drop-in therefore adds two C# classes (GpuIngestionStore, GpuRecallSearchService,
 * see INTEGRATION.md and dotnet/) that P/Invoke the functions below.  Each entry point
 * cites the reference code whose work it replaces.
 *
 * Conventions
 *   - plain C types only; every buffer is caller-owned; inputs are borrowed for the
 *     duration of the call; outputs are written into caller-allocated arrays;
 *   - return 0 (ORR_OK) or a negative ORR_E_* code, never an exception/abort;
 *     orr_last_error() returns a thread-local message for the last failing call;
 *   - one orr_store per GPU shard; orr_search* may be called concurrently from any
 *     number of host threads; mutators are serialised internally (RW lock) and are
 *     safe against in-flight searches;
 *   - there is NO CPU fallback: without a usable CUDA device orr_store_create fails
 *     with ORR_E_CUDA.
 *   - "ticks" are .NET DateTime ticks (100 ns since 0001-01-01), the unit
 *     RecencyScore subtracts (RecallSearchService.cs:115-119).
 */
#ifndef ORR_H_
#define ORR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORR_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------ */
#define ORR_OK              0
#define ORR_E_INVALID      -1   /* bad argument (NULL, negative size, dim mismatch of a buffer) */
#define ORR_E_CUDA         -2   /* CUDA runtime/driver failure, no device, kernel fault         */
#define ORR_E_OOM          -3   /* HBM or pinned-host allocation failed / capacity exhausted    */
#define ORR_E_UNSUPPORTED  -4   /* shape outside what the kernels implement (see limits below)  */
#define ORR_E_INTERNAL     -5   /* invariant violated (bug)                                     */

/* ---- limits ------------------------------------------------------------------------ */
#define ORR_MAX_QUERY_TERMS   64    /* distinct query terms after A-2 filtering           */
#define ORR_MAX_QUERY_PROBES  128   /* (hash,term) probe pairs per query                  */
#define ORR_TICKS_PER_DAY     864000000000LL

typedef struct orr_store orr_store;    /* opaque: one GPU shard */

/* Scorer constants are hard-coded in the reference (RecallSearchService.cs:66,118);
 * they are carried here so the defaults are explicit.  orr_config_default() fills
 * 0.7 / 0.2 / 0.1 / 30.0. */
typedef struct orr_config {
    int32_t  abi_version;        /* ORR_ABI_VERSION                                           */
    int32_t  device;             /* CUDA ordinal of the GPU that owns this shard              */
    int32_t  dim;                /* embedding width D (fp32 per row); multiple of 4, <= 8192  */
    int32_t  term_slots;         /* hashed-term slots per chunk: 32, 64 or 128 (default 128)  */
    int64_t  capacity_rows;      /* rows of HBM reserved up front (row-major fp32[cap][D])    */
    uint64_t row_base;           /* global row id of local row 0 (row-sharded corpora)        */
    double   w_cos, w_kw, w_rec; /* RecallSearchService.cs:66                                 */
    double   recency_days;       /* RecallSearchService.cs:118                                */
} orr_config;

/* One citation, in final reference order (score desc, ticks desc, row asc):
 * what RecallSearchService.cs:34-37 leaves in `scored`.  `row` is the global row id
 * (row_base + local row) the C# shim maps back to its CosmosChunkRecord. */
typedef struct orr_hit {
    uint64_t row;
    double   score;
    int64_t  created_ticks;
} orr_hit;

/* Device/host timing of the last orr_search* call made on the calling thread. */
typedef struct orr_timing {
    float   scan_ms;        /* K1 fused scan kernel(s), CUDA events                       */
    float   finalize_ms;    /* K2 merge + exact fp64 re-score + order                     */
    float   total_device_ms;/* first launch to last kernel end                            */
    float   wall_ms;        /* host wall clock around the whole C-ABI call                */
    int32_t path;           /* ORR_PATH_*                                                 */
    int32_t n_survivors;    /* rows re-scored exactly                                     */
    int64_t rows_scanned;
} orr_timing;

#define ORR_PATH_FUSED      1   /* K1 fused fp32 scan + register top-k, K2 exact re-score     */
#define ORR_PATH_EXACT      2   /* full fp64 scan (no-embedding mode, large k, or escalation) */
#define ORR_PATH_SUBSET     3   /* candidate_cap > 0: exact scoring of the capped subset      */
#define ORR_PATH_BATCH      4   /* tcgen05 batched contraction + re-rank                      */
#define ORR_PATH_TEXT       5   /* orr_search_text: substring matching on the chunk text + exact scoring */
#define ORR_PATH_ESCALATED  0x100 /* OR-ed in when the fused path's bound check failed        */

void orr_config_default(orr_config* cfg);

/* ---- store: replaces InMemoryIngestionStore's chunk side ----------------------------
 * (src/OmniRecall.Api/Services/InMemoryIngestionStore.cs:8-9 dictionaries). */
int  orr_store_create(const orr_config* cfg, orr_store** out);
void orr_store_destroy(orr_store* s);

/* Replace-by-document, the semantics of UpsertChunksAsync
 * (InMemoryIngestionStore.cs:17-25): the document's previous rows are tombstoned and
 * the n chunks are appended in the order given (caller passes them in ChunkIndex order,
 * :23).  emb is n x dim row-major fp32 (may be NULL = no chunk has an embedding);
 * has_emb[i]==0 marks a chunk whose Embedding is null/empty/of another length
 * (RecallSearchService.cs:71-72 -> cosine 0).  term_hashes/term_offsets is a CSR of the
 * DISTINCT orr_hash_term() values of each chunk's lower-cased whitespace tokens
 * (use orr_tokenize_content); at most cfg.term_slots per chunk.  out_rows (may be
 * NULL) receives the n global row ids. */
int  orr_store_upsert_document_chunks(orr_store* s, uint64_t doc_key, int32_t n,
        const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
        const uint64_t* term_hashes, const uint32_t* term_offsets, uint64_t* out_rows);

/* The same, also keeping each chunk's lower-cased UTF-8 Content in HBM (text mode, orr_search_text):
 * chunk i's text is text_lower_utf8[text_offsets[i] .. text_offsets[i+1]).  Lower-casing
 * (ToLowerInvariant, RecallSearchService.cs:110) is the host's job.  Either every row of a store is
 * given text or none is.  Option "text_bytes_per_row" (default 1024) sizes the arena up front. */
int  orr_store_upsert_document_chunks_text(orr_store* s, uint64_t doc_key, int32_t n,
        const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
        const uint64_t* term_hashes, const uint32_t* term_offsets,
        const char* text_lower_utf8, const uint64_t* text_offsets, uint64_t* out_rows);

/* Text-level ingest — the form the C# shim uses: chunk i's Content is contents_utf8[content_offsets[i] ..
 * content_offsets[i+1]) exactly as CosmosChunkRecord.Content holds it.  The library does what KeywordScore does to
 * content (RecallSearchService.cs:110: ToLowerInvariant) plus the split into white-space tokens, hashes the distinct
 * tokens into the chunk's term set, and keeps the LIVE VOCABULARY (every distinct token with the number of live chunks
 * holding it; replace / delete release a document's words) that orr_search_query expands query terms over.  With option
 * "keep_text" = 1 (set before the first row) the lower-cased Content is also kept in HBM for text mode; a chunk with more
 * distinct tokens than cfg.term_slots is then still accepted (keyword matching goes through text mode while such a chunk
 * is live), otherwise it is refused with ORR_E_UNSUPPORTED before anything is changed.  Other arguments as
 * orr_store_upsert_document_chunks. */
int  orr_store_upsert_document_texts(orr_store* s, uint64_t doc_key, int32_t n,
        const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
        const char* contents_utf8, const uint64_t* content_offsets, uint64_t* out_rows);

/* Bulk form of the same (warm load / hydration: the Cosmos store only ever pages `SELECT TOP 300`,
 * CosmosIngestionStore.cs:178-197, a scan-everything store has to be filled from the full container): n_docs documents in
 * one call, document d owning chunks [doc_chunk_offsets[d], doc_chunk_offsets[d+1]) of the chunk arrays (each in
 * ChunkIndex order, keys distinct within a call).  Tokenising runs on all host cores, the rows are copied through pinned
 * staging buffers while searches keep running (they cannot see rows that are not published yet), and one short exclusive
 * section publishes them and tombstones the replaced documents' old rows.  Semantics per document as
 * orr_store_upsert_document_texts. */
int  orr_store_upsert_documents_texts(orr_store* s, int32_t n_docs, const uint64_t* doc_keys, const uint32_t* doc_chunk_offsets,
        const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
        const char* contents_utf8, const uint64_t* content_offsets, uint64_t* out_rows);

/* Distinct tokens held by at least one live chunk (text-level ingest only). */
int64_t orr_store_vocab_size(const orr_store* s);

/* DeleteDocumentAsync (InMemoryIngestionStore.cs:50-55): tombstones the rows. */
int  orr_store_delete_document(orr_store* s, uint64_t doc_key);

/* Squeezes tombstoned rows out (replace-by-document and delete only mark rows dead; every scan still
 * reads them).  Live rows keep their relative order, so the reference's stable row-order tie-break is
 * unchanged, but their ids change: new local row i was old row old_rows_out[i] (global ids, i < *n_live_out);
 * the host remaps its row -> chunk table.  Exclusive: waits for running searches, blocks new ones. */
int  orr_store_compact(orr_store* s, uint64_t* old_rows_out, int64_t out_cap, int64_t* n_live_out);

/* Snapshot / warm load (the HBM store is volatile): orr_store_save writes the store's byte image (rows,
 * tombstones, term tables, chunk text, document -> row table) to `path`; orr_store_load fills an EMPTY store
 * of the same dim / term_slots and capacity >= the snapshot's rows.  Row ids are preserved, so the host's
 * row -> CosmosChunkRecord map (which the host persists itself) stays valid.  Replaces re-ingesting every
 * document after a restart (the Cosmos store's `SELECT TOP 300` paging cannot hydrate a scan-everything
 * store, CosmosIngestionStore.cs:178-197). */
int  orr_store_save(orr_store* s, const char* path);
int  orr_store_load(orr_store* s, const char* path);

/* Runtime knobs.  "batch_passes": how the batched contraction SELECTS candidates; the returned hits are
 * always the exact fp64 re-score, proven complete by the bound check, so every setting returns the same hits.
 *   0 (default) = auto: one bf16 tcgen05 pass screens with a deep candidate list; queries it cannot prove
 *                 are re-run with bf16x3 split precision (fp32-grade), then singly
 *   1 = bf16 screen only (unproven queries run singly)      3 = bf16x3 split precision for every query */
int  orr_store_set_option(orr_store* s, const char* name, double value);

/* Live (non-tombstoned) rows, and rows physically occupied. */
int64_t orr_store_count(const orr_store* s);
int64_t orr_store_rows_used(const orr_store* s);

/* ---- text -> terms: the single definition of tokenising and hashing -----------------
 * orr_tokenize_query restates KeywordScore's query side (RecallSearchService.cs:95-108):
 * split on Unicode white space, lower-case, ordinal-distinct (first occurrence kept),
 * drop the 28 stop words (:13-18) unless that empties the list.  Returns the number of
 * terms in *n (0 => keyword score 0 for every chunk).
 * orr_tokenize_content produces the distinct token hashes of one chunk's Content. */
uint64_t orr_hash_term(const char* utf8_lower, int32_t len);
int  orr_tokenize_query(const char* utf8, int32_t len, uint64_t* out_hashes, int32_t cap, int32_t* n);
int  orr_tokenize_content(const char* utf8, int32_t len, uint64_t* out_hashes, int32_t cap, int32_t* n);

/* ---- search: replaces the scoring loop + ordering of SearchAsync ----------------------
 * (RecallSearchService.cs:28-37 with ScoreChunk :59-67, CosineSimilarity :69-88,
 * KeywordScore :90-113, RecencyScore :115-119).
 *   q/q_dim      query embedding in HOST memory; q_dim==0 => no embedding (cosine 0, :71)
 *   n_terms      |terms| (the denominator of :112); 0 => keyword 0
 *   probe_hash   n_probes term hashes; probe_term[i] in [0,n_terms) says which query term
 *                hash i satisfies (NULL => identity, n_probes==n_terms).  More than one
 *                probe per term lets the host express substring expansion (:111).
 *   now_ticks    the clock RecencyScore reads (:117), injected once per query
 *   top_k        clamped to max(1, top_k) (:36)
 *   candidate_cap 0 = score every live row (north-star behaviour); 300 = the
 *                reference's GetRecentChunksAsync(maxCount: 300) pre-selection (:26)
 *   out/n_out    caller array of >= max(1,top_k) hits; *n_out = hits written (0 if empty)
 */
int  orr_search(orr_store* s, const float* q, int32_t q_dim,
        int32_t n_terms, const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
        int64_t now_ticks, int32_t top_k, int32_t candidate_cap,
        orr_hit* out, int32_t* n_out);

/* Text mode: the keyword predicate of RecallSearchService.cs:110-111 evaluated literally — for every
 * chunk and query term, an ordinal substring search of the term in the chunk's lower-cased content kept
 * in HBM — followed by the exact fp64 scoring of every candidate row.  No vocabulary expansion, no probe
 * limit: this is the path for terms that are substrings of many words ("ai", "go", one letter).  Terms
 * are the lower-cased UTF-8 bytes of the A-2 filtered query terms: term t =
 * terms_lower_utf8[term_offsets[t] .. term_offsets[t+1]); <= 64 terms, <= 256 bytes each.  Other
 * arguments as orr_search.  Slower than the fused scan (every row is scored in fp64). */
int  orr_search_text(orr_store* s, const float* q, int32_t q_dim,
        int32_t n_terms, const char* terms_lower_utf8, const uint32_t* term_offsets,
        int64_t now_ticks, int32_t top_k, int32_t candidate_cap,
        orr_hit* out, int32_t* n_out);

/* The whole keyword side in one call — what IRecallSearchService.SearchAsync(query, topK) needs between the embedding
 * call and the citation build (RecallSearchService.cs:26-37): the query string is split, lower-cased, de-duplicated and
 * stop-word filtered (:95-108, all-stop-words fallback included), every term is expanded into the live vocabulary words
 * that contain it (a GPU scan of the vocabulary kept in HBM; the substring semantics of :110-111), and the probes go
 * through orr_search.  keyword_mode: 0 = auto (text mode when a term expands past ORR_MAX_QUERY_PROBES, when there are
 * more than ORR_MAX_QUERY_TERMS terms, or while an over-long chunk is live), 1 = hashed only, 2 = text mode only.
 * A blank query returns ORR_E_INVALID ("Query is required.", :22-23).  The store must have been fed through
 * orr_store_upsert_document_texts (or orr_store_fill_synthetic with option "synth_vocab"). */
int  orr_search_query(orr_store* s, const char* query_utf8, int32_t query_len, const float* q, int32_t q_dim,
        int64_t now_ticks, int32_t top_k, int32_t candidate_cap, int32_t keyword_mode,
        orr_hit* out, int32_t* n_out);

/* The keyword side orr_search_query derives from a query string, without searching: *n_terms = |terms| after the A-2
 * filtering, and the vocabulary expansion as (hash, term) probes.  *n_probes may exceed cap (nothing beyond cap is
 * written): such a query runs in text mode. */
int  orr_expand_query(orr_store* s, const char* query_utf8, int32_t query_len,
        uint64_t* probe_hash, int32_t* probe_term, int32_t cap, int32_t* n_terms, int32_t* n_probes);

/* Same work with every buffer already resident in HBM on cfg.device and no host
 * synchronisation: q_dev fp32[q_dim], probes copied at enqueue time (host arrays),
 * out_dev orr_hit[max(1,top_k)], status_dev int32[2] = {n_out, flags}; flags bit0 set
 * means the fp32 selection bound check failed and the caller must re-run through
 * orr_search (which escalates to the exact path).  Enqueued on `cuda_stream`
 * (a cudaStream_t passed as void*). */
int  orr_search_device(orr_store* s, const float* q_dev, int32_t q_dim,
        int32_t n_terms, const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
        int64_t now_ticks, int32_t top_k,
        orr_hit* out_dev, int32_t* status_dev, void* cuda_stream);

/* CUDA-event durations of the LAST orr_search_device call on this store (scan_ms = K1, finalize_ms = K3);
 * waits for that call's kernels to finish. */
int  orr_search_device_timing(orr_store* s, orr_timing* out);

/* Batched queries (tcgen05 contraction + exact re-rank).  q is batch x q_dim row-major
 * in HOST memory; terms are a CSR over queries: query b owns probes
 * [probe_offsets[b], probe_offsets[b+1]) and n_terms[b] terms.  out is
 * batch x max(1,top_k); n_out[b] hits are valid in row b. */
int  orr_search_batch(orr_store* s, int32_t batch, const float* q, int32_t q_dim,
        const int32_t* n_terms, const uint64_t* probe_hash, const int32_t* probe_term,
        const uint32_t* probe_offsets,
        int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out);

/* The same with the answers left in HBM on cfg.device: out_dev is orr_hit[batch][max(1,top_k)], n_out_dev int32[batch]
 * (row-sharded batches feed them straight into the all-gather + orr_merge_hits_batch_device; queries still come from
 * host memory).  Complete when the call returns.  Takes what the tcgen05 path takes in one launch (query width = dim,
 * dim % 64 == 0, top_k <= 128, 8 <= batch <= 1024, <= 16 identity-probe terms per query) and returns
 * ORR_E_UNSUPPORTED, with nothing usable written, when the batch must go through orr_search_batch instead (shape
 * outside the path, or a query whose selection the bound check could not prove). */
int  orr_search_batch_device(orr_store* s, int32_t batch, const float* q, int32_t q_dim,
        const int32_t* n_terms, const uint64_t* probe_hash, const int32_t* probe_term,
        const uint32_t* probe_offsets,
        int64_t now_ticks, int32_t top_k, orr_hit* out_dev, int32_t* n_out_dev);

/* Diagnostic for the batched path: the raw fused GEMM scores (w_cos*cos + w_rec*rec, no
 * keyword term; fp32) of every `tile_stride`-th 128-row tile, out[b*out_ld + i]. */
int  orr_debug_batch_scores(orr_store* s, int32_t batch, const float* q, int32_t q_dim,
        int64_t now_ticks, int32_t tile_stride, float* out, int64_t out_ld);

/* Diagnostic for the fused single-query path: the fp32 score K1 computes for every physical row (the values its
 * selection and its discard bound tau are built from): out[row], row < orr_store_rows_used.  FLT_MAX marks a row the scan
 * forces into the survivor list (magnitudes outside what fp32 ranks), -inf a tombstone.  tests use it to MEASURE the
 * bound |fp32 scan score - exact score| <= eps the selection proof assumes. */
int  orr_debug_scan_scores(orr_store* s, const float* q, int32_t q_dim,
        int32_t n_terms, const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
        int64_t now_ticks, float* out, int64_t out_cap);

/* Merge per-shard hit lists (each already in reference order) into the global top-k
 * with the same tie chain; used after the NCCL all-gather of per-GPU candidates. */
int  orr_merge_hits(const orr_hit* lists, const int32_t* list_len, int32_t n_lists,
        int32_t list_stride, int32_t top_k, orr_hit* out, int32_t* n_out);

/* Device-side form of the same merge, for the multi-GPU path: lists_dev is the
 * all-gathered [n_lists][list_stride] hit array, status_dev the all-gathered
 * [n_lists][2] {n_out, flags}; out_status_dev = {n_out, OR of flags}.  Enqueued on
 * `cuda_stream` of `device`, no host synchronisation. */
int  orr_merge_hits_device(int32_t device, const orr_hit* lists_dev, const int32_t* status_dev,
        int32_t n_lists, int32_t list_stride, int32_t top_k,
        orr_hit* out_dev, int32_t* out_status_dev, void* cuda_stream);

/* Batched form for row-sharded orr_search_batch: lists_dev is the all-gathered [n_lists][batch][k] hit array
 * (every shard's answer to every query), n_dev the [n_lists][batch] valid counts; one CTA per query writes the
 * global top-k to out_dev[batch][k] / n_out_dev[batch]. */
int  orr_merge_hits_batch_device(int32_t device, const orr_hit* lists_dev, const int32_t* n_dev, int32_t n_lists,
        int32_t batch, int32_t k, orr_hit* out_dev, int32_t* n_out_dev, void* cuda_stream);

/* ---- fused all-gather + merge over NVLink peer memory ------------------------------------------
 * The multi-GPU exchange step as ONE kernel per rank instead of an NCCL all-gather plus a merge kernel:
 * every rank pushes its exact local top-k (k x 24 B + status) into every peer's exchange buffer with stores
 * over NVLink/NVSwitch, publishes a sequence flag, waits for the other ranks' flags and merges the union with
 * the reference tie chain.  Collective semantics: every rank calls orr_xchg_allgather_merge once per query,
 * in the same order.  Status flag ORR_STATUS_XCHG_TIMEOUT is set if a peer did not arrive within 5 s.
 *   one process per GPU : orr_xchg_create; exchange the 64-byte handles of orr_xchg_get_handle out of band
 *                         (torch.distributed all_gather), orr_xchg_open_peer for every other rank (CUDA IPC)
 *   one process, N GPUs : orr_xchg_create per GPU, orr_xchg_attach_peer (cudaDeviceEnablePeerAccess)        */
typedef struct orr_xchg orr_xchg;
#define ORR_XCHG_HANDLE_BYTES   64
#define ORR_STATUS_BOUND_FAILED 1   /* status flags bit0: fp32 selection not proven, re-run through orr_search */
#define ORR_STATUS_XCHG_TIMEOUT 4   /* status flags bit2: a peer never published its list                     */
int  orr_xchg_create(int32_t device, int32_t world, int32_t rank, int32_t max_top_k, orr_xchg** out);
void orr_xchg_destroy(orr_xchg* x);
int  orr_xchg_get_handle(orr_xchg* x, void* handle_out /* ORR_XCHG_HANDLE_BYTES */);
int  orr_xchg_open_peer(orr_xchg* x, int32_t peer_rank, const void* handle);
int  orr_xchg_attach_peer(orr_xchg* x, int32_t peer_rank, orr_xchg* peer);
int  orr_xchg_allgather_merge(orr_xchg* x, const orr_hit* hits_dev, const int32_t* status_dev, int32_t top_k,
        orr_hit* out_dev, int32_t* out_status_dev, void* cuda_stream);
/* Recovery after ORR_STATUS_XCHG_TIMEOUT: the ranks drain their streams, agree out of band on a number above every
 * orr_xchg_sequence() in use, each calls orr_xchg_resync with it, and a barrier precedes the next exchange.
 * orr_xchg_set_timeout_ms bounds how long the exchange kernel waits for a missing peer (default 5000). */
uint32_t orr_xchg_sequence(orr_xchg* x);
int  orr_xchg_resync(orr_xchg* x, uint32_t next_seq_base);
int  orr_xchg_set_timeout_ms(orr_xchg* x, double ms);

/* ---- one host process, N GPUs (the .NET deployment of the row-sharded layout) ----------------------
 * An orr_cluster owns one orr_store per device (global row id = shard << 40 | local row), their exchange
 * buffers attached to each other, and one stream per device.  orr_cluster_search issues, from the calling
 * thread and without NCCL, for every device: query -> HBM, orr_search_device, orr_xchg_allgather_merge; then
 * reads the merged hits from device 0.  A document lives on one shard (the one that already holds it, else the
 * emptiest), so replace / delete touch one GPU.  Queries without an embedding, of another width, or with
 * top_k > max_top_k run through orr_search on every shard AT ONCE (one host thread per GPU) and are merged on the host.
 * One search or mutation at a time per cluster (a query saturates every GPU's HBM anyway).  If a device fails to
 * launch its part of a query, the cluster drains the others and re-synchronises its exchange buffers before returning. */
typedef struct orr_cluster orr_cluster;
struct orr_synth_spec;
int  orr_cluster_create(const orr_config* cfg /* per shard; device and row_base are overwritten */,
        const int32_t* devices, int32_t n_devices, int32_t max_top_k, orr_cluster** out);
void orr_cluster_destroy(orr_cluster* c);
int32_t    orr_cluster_size(const orr_cluster* c);
orr_store* orr_cluster_shard(orr_cluster* c, int32_t i);      /* borrowed: options, snapshot, compaction per shard */
int64_t    orr_cluster_count(const orr_cluster* c);
int  orr_cluster_upsert_document_chunks(orr_cluster* c, uint64_t doc_key, int32_t n,
        const float* emb, const uint8_t* has_emb, const int64_t* created_ticks,
        const uint64_t* term_hashes, const uint32_t* term_offsets,
        const char* text_lower_utf8 /* or NULL */, const uint64_t* text_offsets, uint64_t* out_rows);
int  orr_cluster_delete_document(orr_cluster* c, uint64_t doc_key);
int  orr_cluster_fill_synthetic(orr_cluster* c, const struct orr_synth_spec* spec, uint64_t first_row, int64_t n_per_shard);
int  orr_cluster_search(orr_cluster* c, const float* q, int32_t q_dim,
        int32_t n_terms, const uint64_t* probe_hash, const int32_t* probe_term, int32_t n_probes,
        int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out);

/* A run of n_queries SINGLE queries, pipelined — the throughput form of orr_cluster_search: the exchange + merge of query i
 * runs on a side stream of every device while that device already scans query i+1 (three queries in flight, each with its
 * own buffers).  q is n_queries x q_dim; terms as a CSR over the queries (as orr_search_batch); out is
 * n_queries x max(1,top_k).  Returns exactly the hits of n_queries orr_cluster_search calls. */
int  orr_cluster_search_many(orr_cluster* c, int32_t n_queries, const float* q, int32_t q_dim,
        const int32_t* n_terms, const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
        int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out);

/* Batched queries over the cluster: orr_search_batch on every shard (one host thread per GPU), then a k-way merge of each
 * query's per-shard lists under the reference tie chain.  Arguments as orr_search_batch. */
int  orr_cluster_search_batch(orr_cluster* c, int32_t batch, const float* q, int32_t q_dim,
        const int32_t* n_terms, const uint64_t* probe_hash, const int32_t* probe_term, const uint32_t* probe_offsets,
        int64_t now_ticks, int32_t top_k, orr_hit* out, int32_t* n_out);

const char* orr_last_error(void);
int  orr_last_timing(orr_timing* out);

/* ---- synthetic corpora (bench/test utility) ---------------------
 * Counter-based generator: every value is a pure function of (seed,row,col), computed
 * with integer arithmetic and correctly-rounded IEEE operations only, so the device
 * fill and the host generator are bit-identical.  orr_synth_rows_host produces the
 * rows on the host (oracle input); orr_store_fill_synthetic appends n rows generated
 * on the device straight into HBM. */
typedef struct orr_synth_spec {
    uint64_t seed;
    int32_t  dim;            /* stored width; rows are the first `dim` components ...     */
    int32_t  gen_dim;        /* ... of a unit vector of this width (dim==gen_dim: unit)   */
    int32_t  terms_per_chunk;/* distinct vocabulary tokens per chunk (<= term_slots)       */
    int32_t  vocab;          /* vocabulary size V, tokens "t%07d"                          */
    int64_t  now_ticks;      /* timestamps are now_ticks - U[0, 365 d)                     */
    int32_t  zero_row_ppm;   /* rows whose embedding is all-zero, parts per million        */
    int32_t  dup_row_ppm;    /* rows that duplicate an earlier row (tie stress)            */
} orr_synth_spec;

void orr_synth_spec_default(orr_synth_spec* spec, int32_t dim);
int  orr_synth_rows_host(const orr_synth_spec* spec, uint64_t first_row, int64_t n,
        float* emb, int64_t* ticks, uint32_t* term_ids /* n x terms_per_chunk */,
        uint64_t* doc_first_row /* n */);
int  orr_synth_query_host(const orr_synth_spec* spec, uint64_t query_index, uint64_t corpus_rows,
        int32_t n_terms, int32_t frequent_terms, float* q /* dim */, uint32_t* term_ids /* n_terms */);
int  orr_synth_term_text(uint32_t term_id, char* out9 /* 8 chars + NUL */);
int64_t orr_synth_row_text(const orr_synth_spec* spec, uint64_t row, char* out, int64_t cap);   /* the row's Content; returns its length */
int  orr_synth_row_info(const orr_synth_spec* spec, uint64_t row, int64_t* ticks, uint64_t* doc_first_row);
int  orr_store_fill_synthetic(orr_store* s, const orr_synth_spec* spec, uint64_t first_row, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* ORR_H_ */
