"""The exact path (orr_exact.cu: fp64 score keys + MSB-first radix select under the reference's tie chain) against
the oracle, through the C ABI.  Needs a B200 (-m gpu).

This is the path of the reference's DEFAULT configuration — no embeddings (NoOpEmbeddingClient.cs:5-8): every cosine
is 0 and the ranking is keyword + recency only, i.e. massive exact score ties resolved by CreatedAtUtc and then by
store order (RecallSearchService.cs:34-37).  Scores within SCORE_RTOL = 1e-12 of the oracle
(contract 1e-5), ids and order exact except inside near-tie groups (tests/util.py)."""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking, oracle_search_synth

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS


def _custom_store(rng, n, dim, slots, vocab, words_per_chunk, n_docs, *, with_emb=True):
    """Rows with explicit token lists over a small vocabulary (so every query term hits many rows) and few distinct
    timestamps (so ties are everywhere).  Returns (shard, emb, ticks, contents)."""
    emb = rng.standard_normal((n, dim)).astype(np.float32) if with_emb else None
    doc_ticks = NOW - rng.integers(0, 40, size=n_docs).astype(np.int64) * DAY
    ticks = doc_ticks[rng.integers(0, n_docs, size=n)]
    contents = [" ".join(rng.choice(vocab, size=rng.integers(1, words_per_chunk + 1), replace=False)) for _ in range(n)]
    sh = orr.RecallShard(dim, n, term_slots=slots)
    hashes = [orr.tokenize_content(c) for c in contents]
    sh.upsert_document_chunks(1, emb, ticks, hashes)
    return sh, emb, ticks, contents


def _oracle(emb, dim, ticks, contents, query, qvec, k, live=None):
    blob, off = oracle_c.pack_contents(contents)
    return oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query=query,
                           qvec=np.zeros(0, np.float32) if qvec is None else qvec, now_ticks=NOW, top_k=k, live=live, threads=4)


@pytest.mark.parametrize("slots,n_query_terms", [(32, 3), (64, 5), (128, 8), (64, 40), (128, 64)])
def test_keyword_and_recency_only_mode_matches_oracle(slots, n_query_terms):
    """No query embedding: orr_noemb_scores_kernel<slots/32> (terms64 + ticks only) + the radix select.  Query terms up
    to ORR_MAX_QUERY_TERMS (two mask words), few distinct timestamps: nearly every score is shared by many rows."""
    rng = np.random.default_rng(slots * 100 + n_query_terms)
    vocab = np.array([f"w{i:03d}" for i in range(90)])
    n, dim = 7_013, 64
    sh, emb, ticks, contents = _custom_store(rng, n, dim, slots, vocab, min(slots, 24), 9)
    with sh:
        for trial in range(4):
            words = rng.choice(vocab, size=n_query_terms, replace=False).tolist()
            query = " ".join(words)
            terms = orr.QueryTerms(n_query_terms, orr.tokenize_query(query), None)
            assert len(terms.probe_hash) == n_query_terms
            for k in (1, 10, 300, 4096):
                got = sh.search(None, terms, NOW, k)
                t = sh.last_timing()
                assert t["path"] == N.PATH_EXACT, t
                er, es, et = _oracle(emb, dim, ticks, contents, query, None, k)
                # scores to 1e-12 (CUDA's exp vs glibc's differ by an ulp now and then); rows exact outside near-tie groups
                assert_same_ranking(got.rows, got.scores, er, es, what=f"slots={slots} terms={n_query_terms} k={k} trial={trial}")
                assert sorted(got.ticks.tolist()) == sorted(et.tolist())


def test_every_row_ties_the_walk_reaches_the_row_digits():
    """No terms, no embedding, one timestamp: all rows have the same score and CreatedAtUtc, so the answer is the first k
    rows in store order (the stable-sort fallback) and the radix walk has to go through all score and ticks digits
    into the row digits.  Then with three timestamps, and with tombstones in front."""
    n, dim = 20_000, 8
    for n_ticks in (1, 3):
        ticks = (NOW - DAY * (np.arange(n) % n_ticks)).astype(np.int64)
        with orr.RecallShard(dim, n) as sh:
            for d in range(4):                                   # four documents of 5000 rows
                sh.upsert_document_chunks(d + 1, None, ticks[d * 5000:(d + 1) * 5000])
            for k in (1, 10, 1000, 4096, 4097, 12_345):
                got = sh.search(None, orr.QueryTerms.none(), NOW, k)
                assert sh.last_timing()["path"] == N.PATH_EXACT
                er, es, et = _oracle(None, dim, ticks, [""] * n, "x", None, k)
                assert got.rows.tolist() == er.tolist(), (n_ticks, k)
                assert_same_ranking(got.rows, got.scores, er, es, what=f"all-ties k={k}")
                assert got.ticks.tolist() == et.tolist()
            sh.delete_document(1)                                # rows 0..4999 become tombstones
            live = np.ones(n, dtype=np.uint8)
            live[:5000] = 0
            for k in (10, 15_000, 20_000):
                got = sh.search(None, orr.QueryTerms.none(), NOW, k)
                er, es, _ = _oracle(None, dim, ticks, [""] * n, "x", None, k, live=live)
                assert len(got) == min(k, 15_000)
                assert got.rows.tolist() == er.tolist(), (n_ticks, k)


def test_large_top_k_with_embeddings_uses_the_global_sort():
    """top_k beyond the in-CTA sorter (4096): exactly k rows are selected by the radix walk and ordered in global
    memory.  With embeddings (general scoring kernel), planted duplicates and zero rows."""
    spec = synth.make_spec(256, gen_dim=256, dup_row_ppm=30000)
    n = 9_000
    rows = synth.rows_host(spec, 0, n)
    with orr.RecallShard(256, n) as sh:
        sh.fill_synthetic(spec, 0, n)
        for qi, k in ((0, 5_000), (1, 8_999), (2, 9_000), (3, 20_000), (4, 4_097)):
            q = synth.query_host(spec, qi, n, n_terms=4)
            got = sh.search(q.q, q.terms, NOW, k)
            assert sh.last_timing()["path"] == N.PATH_EXACT
            er, es, et = oracle_search_synth(rows, q, NOW, k, threads=4)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"big k={k}")
            assert got.ticks.tolist() == et.tolist()


def test_nan_and_negative_scores_keep_the_reference_order():
    """NaN scores sort last (Comparer<double>, :34), negative cosines sort below zero rows; checked on the exact path."""
    rng = np.random.default_rng(5)
    n, dim = 3_000, 32
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    emb[10] = np.nan
    emb[999, 3] = np.inf
    emb[5:8] = 0.0
    ticks = (NOW - DAY * rng.integers(0, 5, size=n)).astype(np.int64)
    q = rng.standard_normal(dim).astype(np.float32)
    with orr.RecallShard(dim, n) as sh:
        sh.upsert_document_chunks(1, emb, ticks)
        for k in (n, 2_990, 300):
            got = sh.search(q, orr.QueryTerms.none(), NOW, k)
            assert sh.last_timing()["path"] == N.PATH_EXACT
            er, es, _ = _oracle(emb, dim, ticks, [""] * n, "x", q, k)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"nan k={k}")
        got = sh.search(q, orr.QueryTerms.none(), NOW, n)
        assert np.isnan(got.scores[-2:]).all() and not np.isnan(got.scores[:-2]).any()
        assert sorted(got.rows[-2:].tolist()) == [10, 999]


@pytest.mark.parametrize("scale", [1e-18, 1e-22, 3e19, 1e25])
def test_queries_whose_norm_leaves_fp32_are_ranked_by_the_exact_path(scale):
    """the fp32 scan's ||q||^2 under/overflows for |q_i| ~ 1e-20 / 1e+19 while the reference accumulates it in
    fp64 and returns a real cosine.  Such queries must escalate (single) or re-run singly (batch), never return cosine 0."""
    spec = synth.make_spec(768)
    n = 6_000
    rows = synth.rows_host(spec, 0, n)
    with orr.RecallShard(768, n) as sh:
        sh.fill_synthetic(spec, 0, n)
        qs = [synth.query_host(spec, qi, n, n_terms=4) for qi in range(12)]
        for q in qs:
            q.q = (q.q.astype(np.float64) * scale).astype(np.float32)
        for q in qs[:3]:
            got = sh.search(q.q, q.terms, NOW, 10)
            assert sh.last_timing()["path"] == N.PATH_EXACT | N.PATH_ESCALATED
            er, es, _ = oracle_search_synth(rows, q, NOW, 10)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"scale {scale}")
        hits = sh.search_batch(np.stack([q.q for q in qs]), [q.terms for q in qs], NOW, 10)
        assert (sh.last_timing()["n_survivors"] & 0xffff) == len(qs)      # every query re-ran singly
        for b, q in enumerate(qs):
            er, es, _ = oracle_search_synth(rows, q, NOW, 10)
            assert_same_ranking(hits[b].rows, hits[b].scores, er, es, what=f"batch scale {scale} b={b}")


def test_a_32_bit_hash_collision_cannot_change_the_result():
    """The no-embedding kernel screens on the 32-bit term table; a stored hash that agrees with a probe in its low word but
    not in its high word is a false match there.  The gather recounts the candidates on the 64-bit table and, when the
    collision leaves fewer than k proven candidates, the query re-runs on the 64-bit kernel: the result is the oracle's."""
    n, dim, k = 4_000, 8, 13
    rng = np.random.default_rng(11)
    word = "needle"
    h = orr.hash_term(word)
    fake = np.uint64(h ^ (1 << 40))                                   # same low 32 bits, different token
    ticks = (NOW - DAY * (1 + rng.integers(0, 300, size=n))).astype(np.int64)
    true_rows = rng.choice(np.arange(100, n), size=12, replace=False)
    contents = ["filler"] * n
    hashes = [np.array([orr.hash_term("filler")], dtype=np.uint64) for _ in range(n)]
    for r in true_rows:
        contents[r] = "filler needle"
        hashes[r] = np.array([orr.hash_term("filler"), h], dtype=np.uint64)
    ticks[7] = NOW                                                    # the colliding row is the newest: first on the screen
    hashes[7] = np.array([orr.hash_term("filler"), fake], dtype=np.uint64)
    with orr.RecallShard(dim, n, term_slots=32) as sh:
        sh.upsert_document_chunks(1, None, ticks, hashes)
        terms = orr.QueryTerms(1, np.array([h], dtype=np.uint64), None)
        for kk in (k, 12, 5, 40):
            got = sh.search(None, terms, NOW, kk)
            assert sh.last_timing()["path"] == N.PATH_EXACT
            er, es, _ = _oracle(None, dim, ticks, contents, word, None, kk)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"collision k={kk}")
        assert 7 not in sh.search(None, terms, NOW, 12).rows.tolist()
