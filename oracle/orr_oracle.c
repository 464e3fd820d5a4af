/*
 * orr_oracle.c — CPU restatement of the reference's hybrid recall scorer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under omni_recall_rag_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker and as the timed CPU baseline.
 *
 * The reference is C# (.NET 10) and cannot be compiled or run in this image (no dotnet /
 * mono / csc), so this is a "port" oracle.  PINNING: it is checked against every fixture
 * the reference's own tests hold for this path (tests/golden/reference_fixtures.json,
 * from tests/OmniRecall.Api.Tests/Services/RecallSearchServiceTests.cs:9-49,51-117,
 * Endpoints/RecallEndpointTests.cs:11-30, Endpoints/ChatEndpointTests.cs:26-100); those
 * tests assert top-1 identity only, so score values beyond that are pinned only by this
 * restatement and by an independent numpy restatement (oracle/oracle_np.py).
 *
 * All citations are to /root/reference/src/OmniRecall.Api/Services/.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math: the arithmetic
 * must stay IEEE, fp32 products widened to fp64 and accumulated sequentially).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct oracle_hit {
    uint64_t row;
    double   score;
    int64_t  created_ticks;
} oracle_hit;

/* ---- UTF-8 helpers -------------------------------------------------------------------- */
static int utf8_decode(const unsigned char* s, int64_t n, uint32_t* cp) {
    if (n <= 0) return 0;
    unsigned c = s[0];
    if (c < 0x80) { *cp = c; return 1; }
    if ((c >> 5) == 6 && n >= 2) { *cp = ((c & 0x1F) << 6) | (s[1] & 0x3F); return 2; }
    if ((c >> 4) == 14 && n >= 3) { *cp = ((c & 0x0F) << 12) | ((s[1] & 0x3F) << 6) | (s[2] & 0x3F); return 3; }
    if ((c >> 3) == 30 && n >= 4) {
        *cp = ((c & 0x07) << 18) | ((s[1] & 0x3F) << 12) | ((s[2] & 0x3F) << 6) | (s[3] & 0x3F);
        return 4;
    }
    *cp = c; return 1; /* invalid byte: pass through */
}
static int utf8_encode(uint32_t cp, unsigned char* out) {
    if (cp < 0x80) { out[0] = (unsigned char)cp; return 1; }
    if (cp < 0x800) { out[0] = 0xC0 | (cp >> 6); out[1] = 0x80 | (cp & 0x3F); return 2; }
    if (cp < 0x10000) { out[0] = 0xE0 | (cp >> 12); out[1] = 0x80 | ((cp >> 6) & 0x3F); out[2] = 0x80 | (cp & 0x3F); return 3; }
    out[0] = 0xF0 | (cp >> 18); out[1] = 0x80 | ((cp >> 12) & 0x3F); out[2] = 0x80 | ((cp >> 6) & 0x3F); out[3] = 0x80 | (cp & 0x3F);
    return 4;
}
/* char.IsWhiteSpace: the separator set of string.Split((char[])null) and
 * string.IsNullOrWhiteSpace (RecallSearchService.cs:92,95). */
static int is_ws(uint32_t c) {
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F ||
           c == 0x205F || c == 0x3000;
}
/* ToLowerInvariant (RecallSearchService.cs:96,110): simple case mapping for ASCII,
 * Latin-1, Latin Extended-A, Greek and Cyrillic; other code points pass through. */
static uint32_t lower_cp(uint32_t c) {
    if (c >= 'A' && c <= 'Z') return c + 32;
    if (c < 0xC0) return c;
    if (c >= 0xC0 && c <= 0xDE && c != 0xD7) return c + 32;
    if (c >= 0x100 && c <= 0x137) return (c & 1) ? c : (c == 0x130 ? 0x69 : c + 1);
    if (c >= 0x139 && c <= 0x148) return (c & 1) ? c + 1 : c;
    if (c >= 0x14A && c <= 0x177) return (c & 1) ? c : c + 1;
    if (c == 0x178) return 0xFF;
    if (c >= 0x179 && c <= 0x17E) return (c & 1) ? c + 1 : c;
    if (c >= 0x391 && c <= 0x3A9 && c != 0x3A2) return c + 32;
    if (c >= 0x400 && c <= 0x40F) return c + 80;
    if (c >= 0x410 && c <= 0x42F) return c + 32;
    return c;
}
/* lower-case a UTF-8 span into out (capacity >= 4*n/ (min 1)); returns bytes written */
static int64_t lower_utf8(const char* s, int64_t n, char* out) {
    int64_t i = 0, o = 0;
    while (i < n) {
        const unsigned char b = (unsigned char)s[i];
        if (b < 0x80) { out[o++] = (char)((b >= 'A' && b <= 'Z') ? b + 32 : b); i += 1; continue; }  /* ASCII fast path */
        uint32_t cp; int k = utf8_decode((const unsigned char*)s + i, n - i, &cp);
        if (k == 1 && ((unsigned char)s[i]) >= 0x80) { out[o++] = s[i]; i += 1; continue; }
        o += utf8_encode(lower_cp(cp), (unsigned char*)out + o);
        i += k;
    }
    return o;
}
static int all_ws(const char* s, int64_t n) {
    int64_t i = 0;
    while (i < n) {
        uint32_t cp; int k = utf8_decode((const unsigned char*)s + i, n - i, &cp);
        if (!is_ws(cp)) return 0;
        i += k;
    }
    return 1;
}

/* ---- A-2 query terms (RecallSearchService.cs:95-108) ----------------------------------- */
static const char* const STOP_WORDS[28] = { /* :13-18 */
    "a", "an", "and", "are", "as", "at", "be", "by", "for", "from", "how", "in", "is",
    "it", "of", "on", "or", "that", "the", "to", "was", "what", "when", "where", "which",
    "who", "why", "with" };

typedef struct term_list {
    char*    buf;      /* lower-cased terms, NUL separated */
    int64_t* off;      /* offset of each term in buf       */
    int64_t* len;
    int32_t  n;
} term_list;

static void term_list_free(term_list* t) { free(t->buf); free(t->off); free(t->len); memset(t, 0, sizeof *t); }

static int is_stop(const char* s, int64_t n) {
    for (int i = 0; i < 28; ++i)
        if ((int64_t)strlen(STOP_WORDS[i]) == n && memcmp(STOP_WORDS[i], s, (size_t)n) == 0) return 1;
    return 0;
}

/* returns |terms| after stop-word filtering with the "all stop words" fallback (:107-108) */
static int32_t query_terms(const char* q, int64_t qn, term_list* out) {
    memset(out, 0, sizeof *out);
    if (qn <= 0 || all_ws(q, qn)) return 0;                         /* :92 */
    out->buf = (char*)malloc((size_t)(2 * qn + 16));
    int64_t max_terms = qn / 1 + 1;
    out->off = (int64_t*)malloc(sizeof(int64_t) * (size_t)max_terms);
    out->len = (int64_t*)malloc(sizeof(int64_t) * (size_t)max_terms);
    int64_t i = 0, o = 0; int32_t n = 0;
    while (i < qn) {                                                /* Split + lower + Distinct :95-98 */
        uint32_t cp; int k = utf8_decode((const unsigned char*)q + i, qn - i, &cp);
        if (is_ws(cp)) { i += k; continue; }
        int64_t start = i;
        while (i < qn) {
            k = utf8_decode((const unsigned char*)q + i, qn - i, &cp);
            if (is_ws(cp)) break;
            i += k;
        }
        int64_t ln = lower_utf8(q + start, i - start, out->buf + o);
        int dup = 0;
        for (int32_t t = 0; t < n && !dup; ++t)
            dup = (out->len[t] == ln && memcmp(out->buf + out->off[t], out->buf + o, (size_t)ln) == 0);
        if (!dup) { out->off[n] = o; out->len[n] = ln; ++n; out->buf[o + ln] = 0; o += ln + 1; }
    }
    if (n == 0) { out->n = 0; return 0; }                           /* :100-101 */
    int32_t kept = 0;                                               /* :103-105 */
    for (int32_t t = 0; t < n; ++t) kept += !is_stop(out->buf + out->off[t], out->len[t]);
    if (kept > 0 && kept < n) {
        int32_t w = 0;
        for (int32_t t = 0; t < n; ++t)
            if (!is_stop(out->buf + out->off[t], out->len[t])) { out->off[w] = out->off[t]; out->len[w] = out->len[t]; ++w; }
        n = w;
    }                                                               /* kept==0 -> rawTerms :107-108 */
    out->n = n;
    return n;
}

static const char* find_sub(const char* hay, int64_t hn, const char* needle, int64_t nn) {
    if (nn == 0) return hay;
    if (nn > hn) return NULL;
    return (const char*)memmem(hay, (size_t)hn, needle, (size_t)nn);
}

/* KeywordScore (:90-113) with the query side hoisted out of the per-chunk loop */
static double keyword_score(const term_list* terms, const char* content, int64_t cn, char* scratch) {
    if (terms->n == 0) return 0.0;
    if (cn <= 0 || all_ws(content, cn)) return 0.0;                 /* :92 */
    int64_t ln = lower_utf8(content, cn, scratch);                  /* :110 */
    int32_t matches = 0;
    for (int32_t t = 0; t < terms->n; ++t)                          /* :111 */
        matches += find_sub(scratch, ln, terms->buf + terms->off[t], terms->len[t]) != NULL;
    return (double)matches / (double)terms->n;                      /* :112 */
}

/* CosineSimilarity (:69-88): fp32 products, widened, accumulated sequentially in fp64 */
double oracle_cosine(const float* a, int32_t na, const float* b, int32_t nb) {
    if (na == 0 || b == NULL || nb == 0 || na != nb) return 0.0;    /* :71-72 */
    double dot = 0.0, normA = 0.0, normB = 0.0;
    for (int32_t i = 0; i < na; ++i) {                              /* :77-82 */
        volatile float pab = a[i] * b[i];
        volatile float paa = a[i] * a[i];
        volatile float pbb = b[i] * b[i];
        dot += (double)pab;
        normA += (double)paa;
        normB += (double)pbb;
    }
    if (normA <= 0.0 || normB <= 0.0) return 0.0;                   /* :84-85 */
    return dot / (sqrt(normA) * sqrt(normB));                       /* :87 */
}

/* same arithmetic without the volatile round trips (-ffp-contract=off and SSE2 keep each
 * product an fp32 mulss); used by the timed baseline, checked equal to oracle_cosine */
static double cosine_fast(const float* a, int32_t na, const float* b, int32_t nb) {
    if (na == 0 || b == NULL || nb == 0 || na != nb) return 0.0;
    double dot = 0.0, normA = 0.0, normB = 0.0;
    for (int32_t i = 0; i < na; ++i) {
        float pab = a[i] * b[i], paa = a[i] * a[i], pbb = b[i] * b[i];
        dot += (double)pab; normA += (double)paa; normB += (double)pbb;
    }
    if (normA <= 0.0 || normB <= 0.0) return 0.0;
    return dot / (sqrt(normA) * sqrt(normB));
}

/* RecencyScore (:115-119) with the clock injected */
double oracle_recency(int64_t now_ticks, int64_t created_ticks) {
    double age_days = (double)(now_ticks - created_ticks) / 864000000000.0;  /* TimeSpan.TotalDays */
    if (!(age_days > 0.0)) age_days = 0.0;                          /* Math.Max(0d, .) */
    return exp(-age_days / 30.0);
}

/* ScoreChunk (:59-67), left to right, no contraction */
double oracle_fuse(double cosv, double kw, double rec) {
    volatile double a = cosv * 0.7;
    volatile double b = kw * 0.2;
    volatile double c = rec * 0.1;
    volatile double ab = a + b;
    return ab + c;
}

double oracle_keyword(const char* query, int64_t qn, const char* content, int64_t cn) {
    term_list t; query_terms(query, qn, &t);
    char* scratch = (char*)malloc((size_t)(2 * (cn > 0 ? cn : 0) + 16));
    double k = keyword_score(&t, content, cn, scratch);
    free(scratch); term_list_free(&t);
    return k;
}

/* writes the lower-cased, filtered query terms NUL-separated into out; returns count */
int32_t oracle_query_terms(const char* query, int64_t qn, char* out, int64_t cap) {
    term_list t; int32_t n = query_terms(query, qn, &t);
    int64_t o = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (o + t.len[i] + 1 > cap) { n = -1; break; }
        memcpy(out + o, t.buf + t.off[i], (size_t)t.len[i]); out[o + t.len[i]] = 0; o += t.len[i] + 1;
    }
    term_list_free(&t);
    return n;
}

/* ---- ordering --------------------------------------------------------------------------
 * Comparer<double>.Default: NaN is less than everything and equal to itself, so under
 * OrderByDescending NaN sorts last (:34).  Returns <0 if x ranks before y. */
typedef struct scored { double score; int64_t ticks; int64_t pos; uint64_t row; } scored;

static int cmp_double_desc(double x, double y) {
    int xn = isnan(x), yn = isnan(y);
    if (xn || yn) return xn - yn;            /* NaN after non-NaN; NaN==NaN */
    return (x > y) ? -1 : (x < y) ? 1 : 0;
}
static int ranks_before(const scored* x, const scored* y) { /* :34-35 then stable (A-6) */
    int c = cmp_double_desc(x->score, y->score);
    if (c) return c;
    if (x->ticks != y->ticks) return x->ticks > y->ticks ? -1 : 1;
    return (x->pos < y->pos) ? -1 : (x->pos > y->pos) ? 1 : 0;
}
static int cmp_scored(const void* a, const void* b) { return ranks_before((const scored*)a, (const scored*)b); }

typedef struct cand { int64_t ticks; int64_t row; } cand;
static int cmp_cand(const void* a, const void* b) { /* OrderByDescending(CreatedAtUtc), stable: InMemoryIngestionStore.cs:61 */
    const cand* x = (const cand*)a; const cand* y = (const cand*)b;
    if (x->ticks != y->ticks) return x->ticks > y->ticks ? -1 : 1;
    return (x->row < y->row) ? -1 : (x->row > y->row);
}

/* keep the best k of a stream under ranks_before: binary heap with the WORST kept at top */
static void heap_sift_down(scored* h, int32_t n, int32_t i) {
    for (;;) {
        int32_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && ranks_before(&h[w], &h[l]) < 0) w = l;
        if (r < n && ranks_before(&h[w], &h[r]) < 0) w = r;
        if (w == i) return;
        scored t = h[i]; h[i] = h[w]; h[w] = t; i = w;
    }
}

/* the scoring loop (:28-33) over a slice of the candidate list; `threads` slices run on
 * pthreads (1 = the reference's sequential LINQ pipeline) */
typedef struct score_job {
    const cand* cands; int64_t nc; int32_t dim; const float* emb; const int64_t* emb_off;
    const char* content; const int64_t* content_off; const int64_t* ticks; const term_list* terms;
    const float* qvec; int32_t q_len; int64_t now_ticks; int64_t max_content; double* scores;
    int32_t part, parts;
} score_job;

static void* score_slice(void* arg) {
    const score_job* j = (const score_job*)arg;
    int64_t lo = j->nc * j->part / j->parts, hi = j->nc * (j->part + 1) / j->parts;
    char* scratch = (char*)malloc((size_t)(2 * j->max_content + 16));
    for (int64_t c = lo; c < hi; ++c) {
        int64_t i = j->cands[c].row;
        const float* b = NULL; int32_t bl = 0;
        if (j->emb) {
            if (j->emb_off) { b = j->emb + j->emb_off[i]; bl = (int32_t)(j->emb_off[i + 1] - j->emb_off[i]); }
            else { b = j->emb + i * (int64_t)j->dim; bl = j->dim; }
        }
        double cs = cosine_fast(j->qvec, j->q_len, b, bl);
        double kw = j->content_off
            ? keyword_score(j->terms, j->content + j->content_off[i], j->content_off[i + 1] - j->content_off[i], scratch)
            : 0.0;
        double rc = oracle_recency(j->now_ticks, j->ticks[i]);
        j->scores[c] = oracle_fuse(cs, kw, rc);
    }
    free(scratch);
    return NULL;
}

static void run_score_jobs(const score_job* proto, int32_t threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads == 1 || proto->nc < 2 * threads) { score_job j = *proto; j.part = 0; j.parts = 1; score_slice(&j); return; }
    pthread_t tid[256]; score_job jobs[256];
    for (int32_t t = 0; t < threads; ++t) {
        jobs[t] = *proto; jobs[t].part = t; jobs[t].parts = threads;
        pthread_create(&tid[t], NULL, score_slice, &jobs[t]);
    }
    for (int32_t t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
}

/*
 * SearchAsync's scoring + ordering (:26-37) over rows given as flat arrays.
 *   emb/emb_off   ragged fp32 embeddings: row i owns emb[emb_off[i] .. emb_off[i+1]);
 *                 emb_off==NULL => every row has exactly `dim` values at emb + i*dim;
 *                 emb==NULL => no row has an embedding
 *   content/content_off  UTF-8 chunk texts (row i = bytes [off[i], off[i+1]))
 *   live          optional 0/1 per row (deleted rows are simply absent in the reference)
 *   candidate_cap 300 = reference (GetRecentChunksAsync(maxCount:300), :26 ->
 *                 InMemoryIngestionStore.cs:57-65, Take(Math.Max(1,maxCount)));
 *                 0 = score every row (the north-star extension)
 *   threads       pthreads for the scoring loop (1 = the reference's sequential LINQ)
 * Returns hits written (Take(Math.Max(1, topK)), :36).
 */
int32_t oracle_search(int64_t n, int32_t dim, const float* emb, const int64_t* emb_off,
                      const char* content, const int64_t* content_off, const int64_t* ticks,
                      const uint8_t* live, const char* query, int64_t query_len,
                      const float* qvec, int32_t q_len, int64_t now_ticks, int32_t top_k,
                      int32_t candidate_cap, int32_t threads, oracle_hit* out) {
    if (n < 0 || top_k > 0x3fffffff) return -1;
    int32_t k = top_k < 1 ? 1 : top_k;
    /* A-1 candidate list */
    int64_t nc = 0;
    cand* cands = (cand*)malloc(sizeof(cand) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i)
        if (!live || live[i]) { cands[nc].ticks = ticks[i]; cands[nc].row = i; ++nc; }
    if (candidate_cap != 0) {
        int64_t cap = candidate_cap < 1 ? 1 : candidate_cap;
        qsort(cands, (size_t)nc, sizeof(cand), cmp_cand);
        if (nc > cap) nc = cap;
    }
    term_list terms; query_terms(query, query_len, &terms);
    int64_t max_content = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t l = content_off ? content_off[i + 1] - content_off[i] : 0;
        if (l > max_content) max_content = l;
    }
    double* scores = (double*)malloc(sizeof(double) * (size_t)(nc > 0 ? nc : 1));
    score_job job = { cands, nc, dim, emb, emb_off, content, content_off, ticks, &terms, qvec, q_len,
                      now_ticks, max_content, scores, 0, 1 };
    run_score_jobs(&job, threads);                                  /* :28-33 */
    /* A-6 selection: best k under (score desc, ticks desc, candidate position asc) */
    scored* heap = (scored*)malloc(sizeof(scored) * (size_t)k);
    int32_t hn = 0;
    for (int64_t c = 0; c < nc; ++c) {
        scored s; s.score = scores[c]; s.ticks = cands[c].ticks;
        /* candidate position: after the stable ticks-desc sort it is (ticks desc, row asc);
         * within equal (score,ticks) that is row asc, so the row is an equivalent key */
        s.pos = cands[c].row; s.row = (uint64_t)cands[c].row;
        if (hn < k) {
            heap[hn++] = s;
            if (hn == k) for (int32_t i = k / 2 - 1; i >= 0; --i) heap_sift_down(heap, k, i);
        } else if (ranks_before(&s, &heap[0]) < 0) {
            heap[0] = s; heap_sift_down(heap, k, 0);
        }
    }
    qsort(heap, (size_t)hn, sizeof(scored), cmp_scored);
    for (int32_t i = 0; i < hn; ++i) { out[i].row = heap[i].row; out[i].score = heap[i].score; out[i].created_ticks = heap[i].ticks; }
    free(heap); free(scores); free(cands); term_list_free(&terms);
    return hn;
}

/* per-row scores for every row (no ordering): lets tests compare component-wise */
int32_t oracle_score_rows(int64_t n, int32_t dim, const float* emb, const int64_t* emb_off,
                          const char* content, const int64_t* content_off, const int64_t* ticks,
                          const char* query, int64_t query_len, const float* qvec, int32_t q_len,
                          int64_t now_ticks, double* out_score, double* out_cos, double* out_kw,
                          double* out_rec) {
    term_list terms; query_terms(query, query_len, &terms);
    int64_t max_content = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t l = content_off ? content_off[i + 1] - content_off[i] : 0;
        if (l > max_content) max_content = l;
    }
    char* scratch = (char*)malloc((size_t)(2 * max_content + 16));
    for (int64_t i = 0; i < n; ++i) {
        const float* b = NULL; int32_t bl = 0;
        if (emb) {
            if (emb_off) { b = emb + emb_off[i]; bl = (int32_t)(emb_off[i + 1] - emb_off[i]); }
            else { b = emb + i * (int64_t)dim; bl = dim; }
        }
        double cs = oracle_cosine(qvec, q_len, b, bl);
        double kw = content_off
            ? keyword_score(&terms, content + content_off[i], content_off[i + 1] - content_off[i], scratch) : 0.0;
        double rc = oracle_recency(now_ticks, ticks[i]);
        if (out_cos) out_cos[i] = cs;
        if (out_kw) out_kw[i] = kw;
        if (out_rec) out_rec[i] = rc;
        out_score[i] = oracle_fuse(cs, kw, rc);
    }
    free(scratch); term_list_free(&terms);
    return 0;
}

/* Math.Round(score, 4) (:51): round-half-to-even of score*1e4, then /1e4 */
double oracle_round4(double x) {
    if (isnan(x) || isinf(x)) return x;
    double p = x * 10000.0;
    double r = nearbyint(p);            /* default rounding mode = to nearest even */
    return r / 10000.0;
}

/* TextSnippetHelper.BuildSnippet (TextSnippetHelper.cs:5-11) over UTF-8 with `max_chars`
 * counted in UTF-16 code units as .NET does; returns bytes written (no NUL). */
int64_t oracle_snippet(const char* content, int64_t n, int32_t max_chars, char* out, int64_t cap) {
    /* replace \n,\r by space, then Trim() (Unicode white space both ends) */
    int64_t s = 0, e = n;
    while (s < e) { uint32_t cp; int k = utf8_decode((const unsigned char*)content + s, e - s, &cp); if (!is_ws(cp)) break; s += k; }
    while (e > s) {
        int64_t p = e - 1; while (p > s && (((unsigned char)content[p]) & 0xC0) == 0x80) --p;
        uint32_t cp; utf8_decode((const unsigned char*)content + p, e - p, &cp);
        if (!is_ws(cp)) break;
        e = p;
    }
    int64_t o = 0; int32_t units = 0; int truncated = 0;
    for (int64_t i = s; i < e;) {
        uint32_t cp; int k = utf8_decode((const unsigned char*)content + i, e - i, &cp);
        int u = cp >= 0x10000 ? 2 : 1;
        if (units + u > max_chars) { truncated = 1; break; }
        if (o + k > cap) return -1;
        if (cp == '\n' || cp == '\r') out[o++] = ' '; else { memcpy(out + o, content + i, (size_t)k); o += k; }
        units += u; i += k;
    }
    if (truncated) { if (o + 3 > cap) return -1; memcpy(out + o, "...", 3); o += 3; }
    return o;
}

int32_t oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int32_t)n;
}
