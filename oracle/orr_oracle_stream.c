/*
 * orr_oracle_stream.c — the oracle at BASELINE sizes: synthetic corpus generated block by block on the
 * host cores and scored by oracle_search (orr_oracle.c), with a running top-k under the reference tie chain.
 *
 * TEST INFRASTRUCTURE ONLY (see orr_oracle.c).  Used by tests/ (full-size parity at 1M x 3072 / 5M x 768) and by
 * bench.py's cpu_baseline / --impl reference legs, which must not map liborr.so: the corpus generator here is built
 * from the header-only definition the device fill kernel uses (omni_recall_rag_b200/csrc/orr_synth.h: integer mixing
 * + correctly rounded IEEE operations, so host and device rows are bit-identical), not from liborr.
 *
 * What is restated: nothing new — a block's rows are exactly the candidate list RecallSearchService.cs:26-33 would
 * score (chunk Content = the chunk's tokens joined by single spaces, SlidingWindowTextChunker.cs:29), each block goes
 * through oracle_search unchanged, and merging the blocks' top-k lists by (score desc / NaN last, CreatedAtUtc desc,
 * row asc) is RecallSearchService.cs:34-37 applied to the union (the global top-k is a subset of the union of the
 * blocks' top-k, and oracle_search orders a block by the same chain).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/orr.h"                              /* orr_synth_spec (plain struct) */
#include "../omni_recall_rag_b200/csrc/orr_synth.h"      /* the generator's single definition */

typedef struct oracle_hit { uint64_t row; double score; int64_t created_ticks; } oracle_hit;

int32_t oracle_search(int64_t n, int32_t dim, const float* emb, const int64_t* emb_off,
                      const char* content, const int64_t* content_off, const int64_t* ticks,
                      const uint8_t* live, const char* query, int64_t query_len,
                      const float* qvec, int32_t q_len, int64_t now_ticks, int32_t top_k,
                      int32_t candidate_cap, int32_t threads, oracle_hit* out);
int32_t oracle_max_threads(void);

void oracle_synth_spec_default(orr_synth_spec* spec, int32_t dim) {
    memset(spec, 0, sizeof *spec);
    spec->seed = 20261018ULL;
    spec->dim = dim;
    spec->gen_dim = dim < 3072 ? 3072 : dim;
    spec->terms_per_chunk = 64;
    spec->vocab = 1 << ORR_SYNTH_VOCAB_LOG2;
    spec->now_ticks = 639963072000000000LL;      /* 2026-10-18T00:00:00Z */
    spec->zero_row_ppm = 10000;
    spec->dup_row_ppm = 0;
}

static int spec_ok(const orr_synth_spec* s) {
    return s && s->dim > 0 && s->gen_dim >= s->dim && s->terms_per_chunk >= 0 && s->terms_per_chunk <= 128 &&
           s->vocab == (1 << ORR_SYNTH_VOCAB_LOG2);
}

/* ---- rows [first_row, first_row + n) on `threads` host threads ------------------------------------------ */
typedef struct gen_job {
    const orr_synth_spec* spec; uint64_t first_row; int64_t n; float* emb; int64_t* ticks; uint32_t* term_ids;
    int32_t part, parts;
} gen_job;

static void* gen_slice(void* arg) {
    const gen_job* j = (const gen_job*)arg;
    const orr_synth_spec* spec = j->spec;
    const int64_t lo = j->n * j->part / j->parts, hi = j->n * (j->part + 1) / j->parts;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t row = j->first_row + (uint64_t)i;
        const uint64_t crow = orr_synth_content_row(spec->seed, row, spec->dup_row_ppm);
        if (j->emb) {
            const int zero = orr_synth_is_zero_row(spec->seed, crow, spec->zero_row_ppm);
            const double scale = zero ? 0.0 : orr_synth_row_scale(spec->seed, crow, spec->gen_dim);
            float* out = j->emb + i * (int64_t)spec->dim;
            for (int c = 0; c < spec->dim; ++c)
                out[c] = orr_synth_scaled(zero ? 0 : orr_synth_component(spec->seed, crow, (uint32_t)c), scale);
        }
        if (j->ticks) j->ticks[i] = orr_synth_row_ticks(spec->seed, row, spec->now_ticks, spec->dup_row_ppm);
        if (j->term_ids) orr_synth_chunk_terms(spec->seed, crow, spec->terms_per_chunk, j->term_ids + i * (int64_t)spec->terms_per_chunk);
    }
    return NULL;
}

int32_t oracle_synth_rows(const orr_synth_spec* spec, uint64_t first_row, int64_t n, float* emb, int64_t* ticks,
                          uint32_t* term_ids, int32_t threads) {
    if (!spec_ok(spec) || n < 0) return -1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (n < 4 * threads) threads = 1;
    gen_job jobs[256]; pthread_t tid[256];
    for (int32_t t = 0; t < threads; ++t) {
        gen_job j = { spec, first_row, n, emb, ticks, term_ids, t, threads };
        jobs[t] = j;
        if (threads > 1) pthread_create(&tid[t], NULL, gen_slice, &jobs[t]);
    }
    if (threads == 1) gen_slice(&jobs[0]);
    else for (int32_t t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

int32_t oracle_synth_query(const orr_synth_spec* spec, uint64_t qi, uint64_t corpus_rows, int32_t n_terms,
                           int32_t frequent_terms, float* q, uint32_t* term_ids) {
    if (!spec_ok(spec) || n_terms < 0 || frequent_terms > n_terms) return -1;
    if (q) {
        uint64_t src = 0;
        const int has_src = orr_synth_query_source(spec->seed, qi, corpus_rows, &src);
        const uint64_t src_c = has_src ? orr_synth_content_row(spec->seed, src, spec->dup_row_ppm) : 0;
        const int use_src = has_src && !orr_synth_is_zero_row(spec->seed, src_c, spec->zero_row_ppm);
        int64_t ss = 0;
        for (int c = 0; c < spec->gen_dim; ++c) {
            const int64_t v = orr_synth_query_component(spec->seed, qi, (uint32_t)c, use_src, src_c);
            ss += v * v;
        }
        const double scale = ss == 0 ? 0.0 : 1.0 / __builtin_sqrt((double)ss);
        for (int c = 0; c < spec->dim; ++c)
            q[c] = orr_synth_scaled(orr_synth_query_component(spec->seed, qi, (uint32_t)c, use_src, src_c), scale);
    }
    if (term_ids) {
        for (int32_t s = 0; s < n_terms; ++s) {
            for (uint32_t attempt = 0;; ++attempt) {
                const uint64_t h = orr_rng(spec->seed + 1, ORR_STREAM_QTERM, qi, ((uint64_t)s << 32) | attempt);
                const uint32_t t = s < frequent_terms ? orr_synth_frequent_token(h) : orr_synth_zipf_token(h);
                int dup = 0;
                for (int32_t k = 0; k < s; ++k) dup |= (term_ids[k] == t);
                if (!dup) { term_ids[s] = t; break; }
            }
        }
    }
    return 0;
}

/* 1 if query qi is "planted" (a corpus row + noise); *src = that row */
int32_t oracle_synth_query_source(const orr_synth_spec* spec, uint64_t qi, uint64_t corpus_rows, uint64_t* src) {
    uint64_t s = 0;
    const int has = orr_synth_query_source(spec->seed, qi, corpus_rows, &s);
    if (src) *src = s;
    return has;
}

/* chunk Content of synthetic rows: the tokens "t%07d" joined by single spaces (SlidingWindowTextChunker.cs:29).
 * blob needs n * max(0, 9*tpc - 1) bytes, off n + 1 entries. */
int64_t oracle_synth_contents(const uint32_t* term_ids, int64_t n, int32_t tpc, char* blob, int64_t* off) {
    int64_t o = 0;
    for (int64_t i = 0; i < n; ++i) {
        off[i] = o;
        for (int32_t s = 0; s < tpc; ++s) {
            if (s) blob[o++] = ' ';
            uint32_t v = term_ids[i * (int64_t)tpc + s];
            blob[o] = 't';
            for (int d = 7; d >= 1; --d) { blob[o + d] = (char)('0' + v % 10u); v /= 10u; }
            o += 8;
        }
    }
    off[n] = o;
    return o;
}

/* ---- ordering of the running list: RecallSearchService.cs:34-35 + stable fallback (global row asc) ------- */
static int cmp_hit(const void* a, const void* b) {
    const oracle_hit* x = (const oracle_hit*)a; const oracle_hit* y = (const oracle_hit*)b;
    const int xn = isnan(x->score), yn = isnan(y->score);
    if (xn || yn) { if (xn != yn) return xn - yn; }
    else if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->created_ticks != y->created_ticks) return x->created_ticks > y->created_ticks ? -1 : 1;
    return (x->row < y->row) ? -1 : (x->row > y->row);
}

/*
 * SearchAsync's scoring + ordering over synthetic rows [first_row, first_row + n_rows) for n_queries queries at
 * once (the corpus is generated once per block and shared by the queries).
 *   queries/query_off   UTF-8 query strings, query i = bytes [query_off[i], query_off[i+1])
 *   qvecs               n_queries x q_len fp32 (q_len == 0: no query embedding)
 *   with_emb            0: rows carry no embedding at all (the reference's default NoOp provider)
 *   out                 n_queries x max(1, top_k) hits, rows are GLOBAL synthetic row ids; n_out per query
 * Returns 0, or -1 on a bad argument / allocation failure.
 */
int32_t oracle_search_streamed(const orr_synth_spec* spec, uint64_t first_row, int64_t n_rows, int64_t block_rows,
                               int32_t n_queries, const char* queries, const int64_t* query_off, const float* qvecs,
                               int32_t q_len, int32_t with_emb, int64_t now_ticks, int32_t top_k, int32_t threads,
                               oracle_hit* out, int32_t* n_out) {
    if (!spec_ok(spec) || n_rows < 0 || n_queries < 0 || block_rows < 1) return -1;
    const int32_t k = top_k < 1 ? 1 : top_k;
    if (threads < 1) threads = oracle_max_threads();
    const int32_t tpc = spec->terms_per_chunk;
    const int64_t text_per_row = tpc > 0 ? 9 * (int64_t)tpc - 1 : 0;
    if (block_rows > n_rows && n_rows > 0) block_rows = n_rows;
    float* emb = with_emb ? (float*)malloc(sizeof(float) * (size_t)block_rows * (size_t)spec->dim) : NULL;
    int64_t* ticks = (int64_t*)malloc(sizeof(int64_t) * (size_t)block_rows);
    uint32_t* tids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)block_rows * (size_t)(tpc > 0 ? tpc : 1));
    char* blob = (char*)malloc((size_t)(block_rows * text_per_row + 16));
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(block_rows + 1));
    oracle_hit* run = (oracle_hit*)malloc(sizeof(oracle_hit) * (size_t)n_queries * 2 * (size_t)k);   /* [q][2k]: kept + new */
    if ((with_emb && !emb) || !ticks || !tids || !blob || !off || !run) { free(emb); free(ticks); free(tids); free(blob); free(off); free(run); return -1; }
    for (int32_t q = 0; q < n_queries; ++q) n_out[q] = 0;
    for (int64_t b0 = 0; b0 < n_rows; b0 += block_rows) {
        const int64_t nb = n_rows - b0 < block_rows ? n_rows - b0 : block_rows;
        oracle_synth_rows(spec, first_row + (uint64_t)b0, nb, emb, ticks, tids, threads);
        oracle_synth_contents(tids, nb, tpc, blob, off);
        for (int32_t q = 0; q < n_queries; ++q) {
            oracle_hit* mine = run + (size_t)q * 2 * k;
            const int32_t got = oracle_search(nb, spec->dim, emb, NULL, blob, off, ticks, NULL,
                                              queries + query_off[q], query_off[q + 1] - query_off[q],
                                              qvecs ? qvecs + (size_t)q * q_len : NULL, qvecs ? q_len : 0, now_ticks, k, 0,
                                              threads, mine + n_out[q]);
            if (got < 0) { free(emb); free(ticks); free(tids); free(blob); free(off); free(run); return -1; }
            for (int32_t i = 0; i < got; ++i) mine[n_out[q] + i].row += first_row + (uint64_t)b0;
            int32_t tot = n_out[q] + got;
            qsort(mine, (size_t)tot, sizeof(oracle_hit), cmp_hit);
            n_out[q] = tot < k ? tot : k;
        }
    }
    for (int32_t q = 0; q < n_queries; ++q)
        memcpy(out + (size_t)q * k, run + (size_t)q * 2 * k, sizeof(oracle_hit) * (size_t)n_out[q]);
    free(emb); free(ticks); free(tids); free(blob); free(off); free(run);
    return 0;
}
