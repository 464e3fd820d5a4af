/*
 * shim_driver.c — the C# shim's side of the boundary, replayed from plain C (no .NET SDK in this image).
 *
 *   shim_driver --layout          prints sizeof / offsetof of orr_config, orr_hit, orr_timing as the C compiler lays
 *                                 them out; tests/test_host.py compares them with the sequential layout of the structs
 *                                 DECLARED in dotnet/OrrNative.cs (parsed from the C# source) and with the ctypes binding
 *   shim_driver --replay <lib>    dlopen()s liborr.so and makes exactly the calls dotnet/GpuIngestionStore.cs and
 *                                 dotnet/GpuRecallSearchService.cs make, in their order, on the reference's own test
 *                                 fixture (tests/OmniRecall.Api.Tests/Services/RecallSearchServiceTests.cs:51-117:
 *                                 three chunks filed under "doc-1" by ONE UpsertChunksAsync call), and prints the hits
 *                                 as JSON lines; tests/test_gpu_parity.py checks them against the oracle.  Needs a GPU.
 *
 * Compile: gcc -O1 -Wall -I include tests/shim/shim_driver.c -ldl -o <out>
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orr.h"

#define OFF(T, f) printf("  \"%s.%s\": %zu,\n", #T, #f, offsetof(T, f))

static int layout(void) {
    printf("{\n");
    printf("  \"sizeof.orr_config\": %zu,\n", sizeof(orr_config));
    OFF(orr_config, abi_version); OFF(orr_config, device); OFF(orr_config, dim); OFF(orr_config, term_slots);
    OFF(orr_config, capacity_rows); OFF(orr_config, row_base); OFF(orr_config, w_cos); OFF(orr_config, w_kw);
    OFF(orr_config, w_rec); OFF(orr_config, recency_days);
    printf("  \"sizeof.orr_hit\": %zu,\n", sizeof(orr_hit));
    OFF(orr_hit, row); OFF(orr_hit, score); OFF(orr_hit, created_ticks);
    printf("  \"sizeof.orr_timing\": %zu,\n", sizeof(orr_timing));
    OFF(orr_timing, scan_ms); OFF(orr_timing, finalize_ms); OFF(orr_timing, total_device_ms); OFF(orr_timing, wall_ms);
    OFF(orr_timing, path); OFF(orr_timing, n_survivors); OFF(orr_timing, rows_scanned);
    printf("  \"ORR_ABI_VERSION\": %d\n}\n", ORR_ABI_VERSION);
    return 0;
}

/* the entry points the shim P/Invokes (dotnet/OrrNative.cs), resolved by name exactly as the .NET loader would */
typedef void (*fn_config_default)(orr_config*);
typedef int (*fn_store_create)(const orr_config*, orr_store**);
typedef void (*fn_store_destroy)(orr_store*);
typedef int (*fn_set_option)(orr_store*, const char*, double);
typedef int (*fn_upsert_texts)(orr_store*, uint64_t, int32_t, const float*, const uint8_t*, const int64_t*, const char*,
                               const uint64_t*, uint64_t*);
typedef int (*fn_delete)(orr_store*, uint64_t);
typedef int (*fn_search_query)(orr_store*, const char*, int32_t, const float*, int32_t, int64_t, int32_t, int32_t, int32_t,
                               orr_hit*, int32_t*);
typedef uint64_t (*fn_hash_term)(const char*, int32_t);
typedef int64_t (*fn_count)(const orr_store*);
typedef const char* (*fn_last_error)(void);

#define SYM(var, type, name)                                                        \
    type var = (type)dlsym(lib, name);                                              \
    if (!var) { fprintf(stderr, "missing symbol %s\n", name); return 3; }
#define CHECK(call)                                                                 \
    do { int _rc = (call); if (_rc != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, _rc, last_error()); return 4; } } while (0)

static int search_and_print(fn_search_query search_query, fn_last_error last_error, orr_store* s, const char* tag,
                            const char* query, const float* q, int q_dim, int64_t now, int top_k, int cap) {
    orr_hit hits[16];
    int32_t n = 0;
    const int rc = search_query(s, query, (int32_t)strlen(query), q, q_dim, now, top_k, cap, 0, hits, &n);
    if (rc != 0) { fprintf(stderr, "orr_search_query(%s) -> %d: %s\n", query, rc, last_error()); return 4; }
    printf("{\"case\": \"%s\", \"query\": \"%s\", \"hits\": [", tag, query);
    for (int i = 0; i < n; ++i)
        printf("%s{\"row\": %llu, \"score\": %.17g, \"ticks\": %lld}", i ? ", " : "", (unsigned long long)hits[i].row, hits[i].score,
               (long long)hits[i].created_ticks);
    printf("]}\n");
    return 0;
}

static int replay(const char* path) {
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(config_default, fn_config_default, "orr_config_default")
    SYM(store_create, fn_store_create, "orr_store_create")
    SYM(store_destroy, fn_store_destroy, "orr_store_destroy")
    SYM(set_option, fn_set_option, "orr_store_set_option")
    SYM(upsert_texts, fn_upsert_texts, "orr_store_upsert_document_texts")
    SYM(delete_document, fn_delete, "orr_store_delete_document")
    SYM(search_query, fn_search_query, "orr_search_query")
    SYM(hash_term, fn_hash_term, "orr_hash_term")
    SYM(store_count, fn_count, "orr_store_count")
    SYM(last_error, fn_last_error, "orr_last_error")

    /* GpuIngestionStore..ctor */
    orr_config cfg;
    config_default(&cfg);
    cfg.device = 0; cfg.dim = 4; cfg.term_slots = 128; cfg.capacity_rows = 1024;
    orr_store* s = NULL;
    CHECK(store_create(&cfg, &s));
    CHECK(set_option(s, "text_bytes_per_row", 2048.0));
    CHECK(set_option(s, "keep_text", 1.0));

    /* UpsertChunksAsync: the fixture's three chunks in one call => filed under chunks[0].DocumentId = "doc-1" */
    const int64_t now = 639963072000000000LL;
    const char* contents[3] = {"azure cosmos db vector search", "kubernetes deployment yaml and helm chart", "what is the and of for"};
    const float emb[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};           /* [1,0], [0,1], [0,0] padded to the store width */
    const uint8_t has[3] = {1, 1, 1};
    const int64_t ticks[3] = {now, now, now};
    char blob[256]; uint64_t off[4]; size_t o = 0;
    for (int i = 0; i < 3; ++i) { off[i] = o; memcpy(blob + o, contents[i], strlen(contents[i])); o += strlen(contents[i]); }
    off[3] = o;
    uint64_t rows[3];
    const char* key = "doc:doc-1";
    const uint64_t doc_key = hash_term(key, (int32_t)strlen(key));
    CHECK(upsert_texts(s, doc_key, 3, emb, has, ticks, blob, off, rows));
    printf("{\"case\": \"upsert\", \"rows\": [%llu, %llu, %llu], \"count\": %lld}\n", (unsigned long long)rows[0],
           (unsigned long long)rows[1], (unsigned long long)rows[2], (long long)store_count(s));

    /* SearchAsync x 3: RecallSearchServiceTests.cs:9-21, :24-35, :38-49 (candidate cap 300 = the reference) */
    const float q_azure[4] = {1, 0, 0, 0};
    int rc;
    if ((rc = search_and_print(search_query, last_error, s, "vector+keyword", "azure", q_azure, 4, now, 3, 300))) return rc;
    if ((rc = search_and_print(search_query, last_error, s, "keyword-only", "kubernetes", NULL, 0, now, 3, 300))) return rc;
    if ((rc = search_and_print(search_query, last_error, s, "stop-words", "what is the kubernetes", NULL, 0, now, 3, 300))) return rc;

    /* a blank query is the host's ArgumentException; the library refuses it too */
    orr_hit h1[1]; int32_t n1 = 0;
    const int blank = search_query(s, "   ", 3, NULL, 0, now, 3, 300, 0, h1, &n1);
    printf("{\"case\": \"blank\", \"rc\": %d}\n", blank);

    /* DeleteDocumentAsync, then the empty-store search (ChatEndpointTests.cs:26-58: no citations, no error) */
    CHECK(delete_document(s, doc_key));
    if ((rc = search_and_print(search_query, last_error, s, "empty-store", "azure", q_azure, 4, now, 3, 300))) return rc;
    store_destroy(s);                                                     /* Dispose() */
    dlclose(lib);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "--layout")) return layout();
    if (argc >= 3 && !strcmp(argv[1], "--replay")) return replay(argv[2]);
    fprintf(stderr, "usage: shim_driver --layout | --replay <liborr.so>\n");
    return 1;
}
