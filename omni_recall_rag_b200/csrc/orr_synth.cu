// orr_synth.cu — synthetic corpus: device fill kernel + the host generator, both built from
// the single definition in orr_synth.h (bench/test utility).
#include <cstdio>
#include <cstring>

#include "orr_internal.h"
#include "orr_synth.h"

// hash of the fixed-width token text "t%07d" — must equal orr_hash_term() of that text
__host__ __device__ static inline uint64_t synth_term_hash(uint32_t id) {
    char t[8];
    t[0] = 't';
    uint32_t v = id;
    for (int i = 7; i >= 1; --i) { t[i] = (char)('0' + (v % 10u)); v /= 10u; }
    // FNV-1a 64 + murmur finaliser: the same function as orr_hash_bytes (orr_text.cpp)
    uint64_t h = 0xCBF29CE484222325ULL;
    for (int i = 0; i < 8; ++i) { h ^= (uint8_t)t[i]; h *= 0x100000001B3ULL; }
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDULL; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ULL; h ^= h >> 33;
    return h ? h : 1ULL;
}

namespace {

struct FillArgs {
    float* emb; int64_t* ticks; uint32_t* terms32; uint64_t* terms64;
    int32_t dim, slots;
    orr_synth_spec spec;
    uint64_t first_row;       // global synthetic row id of the first generated row
    int64_t  local_first;     // local row index it lands in
    int64_t  n;
    // optional chunk text (text mode): the tokens "t%07d" joined by single spaces, 9*tpc-1 bytes per row
    uint8_t* text; uint64_t* text_off; uint32_t* text_len; uint64_t text_base;
};

// one warp per row: the lanes split the columns; lane 0 draws the terms and the timestamp
__global__ void __launch_bounds__(256) orr_synth_fill_kernel(const FillArgs a) {
    __shared__ uint32_t s_ids[8][128];
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gw; i < a.n; i += W) {
        const uint64_t row = a.first_row + (uint64_t)i;
        const int64_t lrow = a.local_first + i;
        const uint64_t crow = orr_synth_content_row(a.spec.seed, row, a.spec.dup_row_ppm);
        const bool zero = orr_synth_is_zero_row(a.spec.seed, crow, a.spec.zero_row_ppm);
        // sum of squares over gen_dim columns, exact in int64, lane-strided then reduced
        long long ss = 0;
        if (!zero) {
            for (int c = lane; c < a.spec.gen_dim; c += 32) {
                const long long v = orr_synth_component(a.spec.seed, crow, (uint32_t)c);
                ss += v * v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        const double scale = (zero || ss == 0) ? 0.0 : __ddiv_rn(1.0, __dsqrt_rn((double)ss));
        float* out = a.emb + lrow * (int64_t)a.dim;
        for (int c = lane; c < a.dim; c += 32) {
            const long long v = zero ? 0 : orr_synth_component(a.spec.seed, crow, (uint32_t)c);
            out[c] = orr_synth_scaled(v, scale);
        }
        const int tpc = a.spec.terms_per_chunk;
        uint32_t* ids = s_ids[threadIdx.x >> 5];
        __syncwarp();
        if (lane == 0) {
            a.ticks[lrow] = orr_synth_row_ticks(a.spec.seed, row, a.spec.now_ticks, a.spec.dup_row_ppm);
            orr_synth_chunk_terms(a.spec.seed, crow, tpc, ids);
        }
        __syncwarp();
        for (int s = lane; s < a.slots; s += 32) {
            uint64_t h = 0;
            if (s < tpc) h = synth_term_hash(ids[s]);
            a.terms64[lrow * (int64_t)a.slots + s] = h;
            a.terms32[lrow * (int64_t)a.slots + s] = h ? orr_hash_low(h) : 0u;
        }
        if (a.text) {
            const uint32_t len = tpc > 0 ? (uint32_t)(9 * tpc - 1) : 0u;
            const uint64_t off = a.text_base + (uint64_t)i * len;
            if (lane == 0) { a.text_off[lrow] = off; a.text_len[lrow] = len; }
            for (uint32_t p = lane; p < len; p += 32) {
                const uint32_t tok = p / 9u, c = p - tok * 9u;          // c: 0 't', 1..7 digits, 8 space
                uint8_t ch = ' ';
                if (c == 0) ch = 't';
                else if (c < 8) {
                    uint32_t v = ids[tok];
                    for (uint32_t d = 7; d > c; --d) v /= 10u;
                    ch = (uint8_t)('0' + v % 10u);
                }
                a.text[off + p] = ch;
            }
        }
    }
}

}  // namespace

int orr_launch_synth_fill(float* emb, int64_t* ticks, uint32_t* terms32, uint64_t* terms64,
                          int dim, int slots, const orr_synth_spec& spec, uint64_t first_row,
                          int64_t local_first, int64_t n, uint8_t* text, uint64_t* text_off, uint32_t* text_len,
                          uint64_t text_base, cudaStream_t st) {
    if (n <= 0) return ORR_OK;
    FillArgs a{emb, ticks, terms32, terms64, dim, slots, spec, first_row, local_first, n, text, text_off, text_len, text_base};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    orr_synth_fill_kernel<<<sms * 8, 256, 0, st>>>(a);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}

// ---- host side of the generator ----------------------------------------------------------
extern "C" {

void orr_synth_spec_default(orr_synth_spec* spec, int32_t dim) {
    memset(spec, 0, sizeof *spec);
    spec->seed = 20261018ULL;
    spec->dim = dim;
    spec->gen_dim = dim < 3072 ? 3072 : dim;
    spec->terms_per_chunk = 64;
    spec->vocab = 1 << ORR_SYNTH_VOCAB_LOG2;
    spec->now_ticks = 639963072000000000LL;      // 2026-10-18T00:00:00Z in .NET ticks
    spec->zero_row_ppm = 10000;                  // 1 % embedding failures
    spec->dup_row_ppm = 0;
}

static int synth_spec_ok(const orr_synth_spec* s) {
    if (!s || s->dim <= 0 || s->gen_dim < s->dim || s->terms_per_chunk < 0 || s->terms_per_chunk > 128 ||
        s->vocab != (1 << ORR_SYNTH_VOCAB_LOG2)) {
        orr_set_error("synth: bad spec");
        return 0;
    }
    return 1;
}

int orr_synth_rows_host(const orr_synth_spec* spec, uint64_t first_row, int64_t n, float* emb,
                        int64_t* ticks, uint32_t* term_ids, uint64_t* doc_first_row) {
    if (!synth_spec_ok(spec) || n < 0) return ORR_E_INVALID;
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t row = first_row + (uint64_t)i;
        const uint64_t crow = orr_synth_content_row(spec->seed, row, spec->dup_row_ppm);
        if (emb) {
            const int zero = orr_synth_is_zero_row(spec->seed, crow, spec->zero_row_ppm);
            const double scale = zero ? 0.0 : orr_synth_row_scale(spec->seed, crow, spec->gen_dim);
            float* out = emb + i * (int64_t)spec->dim;
            for (int c = 0; c < spec->dim; ++c)
                out[c] = orr_synth_scaled(zero ? 0 : orr_synth_component(spec->seed, crow, (uint32_t)c), scale);
        }
        if (ticks) ticks[i] = orr_synth_row_ticks(spec->seed, row, spec->now_ticks, spec->dup_row_ppm);
        if (term_ids) orr_synth_chunk_terms(spec->seed, crow, spec->terms_per_chunk,
                                            term_ids + i * (int64_t)spec->terms_per_chunk);
        if (doc_first_row) doc_first_row[i] = orr_synth_doc_first_row(spec->seed, row);
    }
    return ORR_OK;
}

int orr_synth_query_host(const orr_synth_spec* spec, uint64_t qi, uint64_t corpus_rows, int32_t n_terms,
                         int32_t frequent_terms, float* q, uint32_t* term_ids) {
    if (!synth_spec_ok(spec) || n_terms < 0 || frequent_terms > n_terms) return ORR_E_INVALID;
    if (q) {
        uint64_t src = 0;
        const int has_src = orr_synth_query_source(spec->seed, qi, corpus_rows, &src);
        const uint64_t src_c = has_src ? orr_synth_content_row(spec->seed, src, spec->dup_row_ppm) : 0;
        // a zero source row contributes nothing: the query is then pure noise
        const int use_src = has_src && !orr_synth_is_zero_row(spec->seed, src_c, spec->zero_row_ppm);
        int64_t ss_hi = 0; double ssd = 0.0;
        // components reach 21*131070 ~ 2.75e6, squares ~7.6e12, 3072 of them fit int64
        for (int c = 0; c < spec->gen_dim; ++c) {
            const int64_t v = orr_synth_query_component(spec->seed, qi, (uint32_t)c, use_src, src_c);
            ss_hi += v * v;
        }
        ssd = (double)ss_hi;
        const double scale = ss_hi == 0 ? 0.0 : 1.0 / __builtin_sqrt(ssd);
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
    return ORR_OK;
}

// Content of synthetic row `row`: its tokens joined by single spaces (what the reference's chunker would have stored,
// SlidingWindowTextChunker.cs:29); returns the byte length (9 * terms_per_chunk - 1), or a negative code.
int64_t orr_synth_row_text(const orr_synth_spec* spec, uint64_t row, char* out, int64_t cap) {
    if (!synth_spec_ok(spec) || !out) return ORR_E_INVALID;
    const int tpc = spec->terms_per_chunk;
    const int64_t len = tpc > 0 ? 9 * (int64_t)tpc - 1 : 0;
    if (cap < len) { orr_set_error("orr_synth_row_text: buffer too small"); return ORR_E_INVALID; }
    uint32_t ids[128];
    orr_synth_chunk_terms(spec->seed, orr_synth_content_row(spec->seed, row, spec->dup_row_ppm), tpc, ids);
    int64_t o = 0;
    for (int s = 0; s < tpc; ++s) {
        if (s) out[o++] = ' ';
        uint32_t v = ids[s];
        out[o] = 't';
        for (int d = 7; d >= 1; --d) { out[o + d] = (char)('0' + v % 10u); v /= 10u; }
        o += 8;
    }
    return len;
}

// CreatedAtUtc ticks and the first row of the document (run of rows sharing one timestamp) of synthetic row `row`
int orr_synth_row_info(const orr_synth_spec* spec, uint64_t row, int64_t* ticks, uint64_t* doc_first_row) {
    if (!synth_spec_ok(spec)) return ORR_E_INVALID;
    if (ticks) *ticks = orr_synth_row_ticks(spec->seed, row, spec->now_ticks, spec->dup_row_ppm);
    if (doc_first_row) *doc_first_row = orr_synth_doc_first_row(spec->seed, row);
    return ORR_OK;
}

int orr_synth_term_text(uint32_t term_id, char* out9) {
    if (!out9 || term_id >= (1u << ORR_SYNTH_VOCAB_LOG2)) return ORR_E_INVALID;
    snprintf(out9, 9, "t%07u", term_id);
    return ORR_OK;
}

}  // extern "C"
