// orr_synth.h — counter-based synthetic corpus generator, shared by the
// host generator (orr_synth_rows_host) and the device fill kernel.  Every value is a pure
// function of (seed, row, col) built from 64-bit integer mixing plus IEEE operations that
// are correctly rounded on both CPU and GPU (int->double, double sqrt, double divide,
// double multiply, double->float), so the two sides are bit-identical by construction.
//
// Shape of the data (mirrors what the reference's ingest produces):
//   - embeddings: near-Gaussian components (sum of four 16-bit uniforms), each row scaled to
//     unit L2 norm over gen_dim columns, of which the first `dim` are stored (dim < gen_dim
//     == "truncated embeddings", norms != 1);
//   - documents: runs of 1..64 consecutive rows sharing one CreatedAtUtc tick value
//     (DocumentIngestionService.cs:102 stamps one timestamp per document);
//   - terms: `terms_per_chunk` distinct tokens per chunk, Zipf-like (octave-uniform, P(rank)
//     ~ 1/rank) over a vocabulary of 2^20 fixed-width tokens "t%07d" (equal width => the
//     reference's substring Contains (RecallSearchService.cs:111) == set membership);
//   - zero rows (embedding failed -> cosine 0, :84-85) and duplicate rows (tie stress).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORR_HD __host__ __device__ __forceinline__
#else
#define ORR_HD static inline
#endif

#define ORR_SYNTH_VOCAB_LOG2 20
#define ORR_SYNTH_YEAR_TICKS (365LL * 864000000000LL)

enum { ORR_STREAM_EMB = 1, ORR_STREAM_DOC = 2, ORR_STREAM_TS = 3, ORR_STREAM_TERM = 4,
       ORR_STREAM_ZERO = 5, ORR_STREAM_DUP = 6, ORR_STREAM_QKIND = 7, ORR_STREAM_QTERM = 8 };

ORR_HD uint64_t orr_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

ORR_HD uint64_t orr_rng(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    uint64_t x = orr_mix64(seed + 0x9E3779B97F4A7C15ULL * (stream + 1));
    x = orr_mix64(x ^ (a * 0xD1342543DE82EF95ULL + 0x632BE59BD9B4E019ULL));
    x = orr_mix64(x + b * 0x9E3779B97F4A7C15ULL);
    return x;
}

// integer component in [-131070, 131070]; variance 4*(65536^2-1)/12
ORR_HD int32_t orr_synth_component(uint64_t seed, uint64_t row, uint32_t col) {
    uint64_t h = orr_rng(seed, ORR_STREAM_EMB, row, col);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) +
                (int32_t)((h >> 32) & 0xFFFF) + (int32_t)(h >> 48);
    return s - 131070;
}

// row whose CONTENT (embedding + terms) a row carries: itself, or an earlier row if it is
// a planted duplicate.  One level only: the source's own duplicate flag is ignored.
ORR_HD uint64_t orr_synth_content_row(uint64_t seed, uint64_t row, int32_t dup_ppm) {
    if (dup_ppm <= 0 || row == 0) return row;
    uint64_t h = orr_rng(seed, ORR_STREAM_DUP, row, 0);
    if ((int64_t)(h % 1000000ULL) >= dup_ppm) return row;
    return orr_rng(seed, ORR_STREAM_DUP, row, 1) % row;
}

ORR_HD int orr_synth_is_zero_row(uint64_t seed, uint64_t content_row, int32_t zero_ppm) {
    if (zero_ppm <= 0) return 0;
    return (int64_t)(orr_rng(seed, ORR_STREAM_ZERO, content_row, 0) % 1000000ULL) < zero_ppm;
}

// first row of the document `row` belongs to: blocks of 64 rows are cut where a sparse
// 64-bit mask (density 1/8, bit 0 forced) has a set bit.
ORR_HD uint64_t orr_synth_doc_first_row(uint64_t seed, uint64_t row) {
    uint64_t b = row >> 6; uint32_t j = (uint32_t)(row & 63);
    uint64_t m = orr_rng(seed, ORR_STREAM_DOC, b, 0) & orr_rng(seed, ORR_STREAM_DOC, b, 1) &
                 orr_rng(seed, ORR_STREAM_DOC, b, 2);
    m |= 1ULL;
    m &= (~0ULL) >> (63 - j);
    uint32_t hi = 63;
    while (!((m >> hi) & 1ULL)) --hi;          // bit 0 is set, terminates
    return (b << 6) + hi;
}

ORR_HD int64_t orr_synth_doc_ticks(uint64_t seed, uint64_t doc_first_row, int64_t now_ticks) {
    uint64_t age = orr_rng(seed, ORR_STREAM_TS, doc_first_row, 0) % (uint64_t)ORR_SYNTH_YEAR_TICKS;
    return now_ticks - (int64_t)age;
}

// CreatedAtUtc ticks of a row.  Duplicates: even duplicate rows take their source row's
// timestamp (exercises the stable row-order fallback), odd ones keep their own document's
// (exercises ThenByDescending(CreatedAtUtc), RecallSearchService.cs:35).
ORR_HD int64_t orr_synth_row_ticks(uint64_t seed, uint64_t row, int64_t now_ticks, int32_t dup_ppm) {
    uint64_t c = orr_synth_content_row(seed, row, dup_ppm);
    uint64_t r = (c != row && (row & 1ULL) == 0) ? c : row;
    return orr_synth_doc_ticks(seed, orr_synth_doc_first_row(seed, r), now_ticks);
}

// Zipf-like token id in [0, 2^20 - 1): octave o uniform in [0,20), rank uniform in
// [2^o, 2^(o+1)), id = rank - 1.  P(rank) ~ 1/rank; ranks < 1024 carry half the mass.
ORR_HD uint32_t orr_synth_zipf_token(uint64_t h) {
    uint32_t o = (uint32_t)((h >> 40) % ORR_SYNTH_VOCAB_LOG2);
    uint32_t rank = (1u << o) + (uint32_t)(h & ((1ULL << o) - 1ULL));
    return rank - 1u;
}
// uniform over the 1023 most frequent tokens (ranks 1..1023)
ORR_HD uint32_t orr_synth_frequent_token(uint64_t h) { return (uint32_t)(h % 1023ULL); }

// the `tpc` distinct tokens of a chunk, in draw order (rejection on repeats)
ORR_HD void orr_synth_chunk_terms(uint64_t seed, uint64_t content_row, int32_t tpc, uint32_t* out) {
    for (int32_t s = 0; s < tpc; ++s) {
        for (uint32_t attempt = 0;; ++attempt) {
            uint32_t t = orr_synth_zipf_token(
                orr_rng(seed, ORR_STREAM_TERM, content_row, ((uint64_t)s << 32) | attempt));
            int dup = 0;
            for (int32_t k = 0; k < s; ++k) dup |= (out[k] == t);
            if (!dup) { out[s] = t; break; }
        }
    }
}

// scale that takes the integer components of a row to a unit vector over gen_dim columns
ORR_HD double orr_synth_row_scale(uint64_t seed, uint64_t content_row, int32_t gen_dim) {
    int64_t ss = 0;
    for (int32_t c = 0; c < gen_dim; ++c) {
        int64_t v = orr_synth_component(seed, content_row, (uint32_t)c);
        ss += v * v;
    }
    if (ss == 0) return 0.0;
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(1.0, __dsqrt_rn((double)ss));
#else
    return 1.0 / __builtin_sqrt((double)ss);
#endif
}

ORR_HD float orr_synth_scaled(int64_t v, double scale) {
#if defined(__CUDA_ARCH__)
    return __double2float_rn(__dmul_rn((double)v, scale));
#else
    return (float)((double)v * scale);
#endif
}

// ---- queries --------------------------------------------------------------------------
// query qi is an independent draw (seed+1), or — one in ten — a corpus row plus 1/20 noise
ORR_HD int orr_synth_query_source(uint64_t seed, uint64_t qi, uint64_t corpus_rows, uint64_t* src) {
    uint64_t h = orr_rng(seed + 1, ORR_STREAM_QKIND, qi, 0);
    if (corpus_rows == 0 || (h % 10ULL) != 0) return 0;
    *src = orr_rng(seed + 1, ORR_STREAM_QKIND, qi, 1) % corpus_rows;
    return 1;
}
ORR_HD int64_t orr_synth_query_component(uint64_t seed, uint64_t qi, uint32_t col, int has_src,
                                         uint64_t src_content_row) {
    int64_t v = orr_synth_component(seed + 1, qi, col);
    if (has_src) v += 20LL * (int64_t)orr_synth_component(seed, src_content_row, col);
    return v;
}
