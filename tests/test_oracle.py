"""The oracle itself: pinned against the reference's own fixtures, and the two independent
restatements (C and numpy) triangulated against each other.  CPU only."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle_c, oracle_np

NOW = 639_963_072_000_000_000
DAY = 864_000_000_000


def _run_c(case):
    chunks = case["chunks"]
    n = len(chunks)
    blob, off = oracle_c.pack_contents([c["content"] for c in chunks])
    lens = [len(c["embedding"]) for c in chunks]
    emb_off = np.zeros(n + 1, dtype=np.int64)
    emb_off[1:] = np.cumsum(lens)
    emb = np.array([v for c in chunks for v in c["embedding"]] or [0.0], dtype=np.float32)
    ticks = np.full(n, case["now_ticks"], dtype=np.int64)
    return oracle_c.search(emb=emb, dim=0, emb_off=emb_off, ticks=ticks, content_blob=blob, content_off=off,
                           query=case["query"], qvec=np.array(case["query_embedding"], dtype=np.float32),
                           now_ticks=case["now_ticks"], top_k=case["top_k"], candidate_cap=300)


def test_golden_fixtures_pin_both_oracles(golden):
    """Every fixture the reference's tests hold for this path: the asserted top-1 (pinned by
    the reference) and the derived order/scores."""
    assert len(golden["cases"]) == 6
    for case in golden["cases"]:
        rows, scores, _ = _run_c(case)
        exp = case["derived_hits"]
        assert [int(r) for r in rows] == [h["row"] for h in exp], case["name"]
        for s, h in zip(scores, exp):
            assert s == float.fromhex(h["score_hex"]), (case["name"], s, h)   # bit-exact
            assert oracle_c.round4(s) == h["rounded"]
        if case["asserted_first_row"] is not None:
            assert int(rows[0]) == case["asserted_first_row"]
        else:
            assert len(rows) == 0
        # numpy restatement agrees too
        recs = [oracle_np.Chunk(c["content"], c["embedding"], case["now_ticks"], i) for i, c in enumerate(case["chunks"])]
        hits = oracle_np.search(recs, case["query"], case["query_embedding"], case["now_ticks"], case["top_k"], 300)
        assert [h[0] for h in hits] == [h["row"] for h in exp]
        assert [h[1] for h in hits] == [float.fromhex(h["score_hex"]) for h in exp]


def test_appendix_b_known_answers(golden):
    by = {c["name"]: c for c in golden["cases"]}
    assert by["with_embeddings_most_similar_first"]["derived_hits"][0]["score"] == 0.9999999999999999
    assert by["no_query_embedding_falls_back_to_keyword"]["derived_hits"][0]["score"] == 0.30000000000000004
    assert by["chat_after_upload_citation_passes_guard"]["derived_hits"][0]["score"] == 0.8999999999999998
    assert by["chat_after_upload_citation_passes_guard"]["derived_hits"][0]["rounded"] == 0.9  # >= 0.25 guard
    assert by["endpoint_after_upload"]["derived_hits"][0]["rounded"] == 0.3


def test_stop_words_and_terms():
    assert oracle_c.query_terms("what is the kubernetes") == ["kubernetes"]
    assert oracle_c.query_terms("What backend did we choose?") == ["backend", "did", "we", "choose?"]
    assert oracle_c.query_terms("the of and") == ["the", "of", "and"]        # all stop words -> raw terms
    assert oracle_c.query_terms("Azure AZURE  azure\tCosmos") == ["azure", "cosmos"]
    assert oracle_c.query_terms("   \t\n") == []
    assert oracle_c.query_terms("a b c") == ["b", "c"]             # NBSP / EM SPACE split; 'a' is a stop word
    for q in ["what is the kubernetes", "What backend did we choose?", "the of and", "ÀÉÎ Straße ΣΟΦΙΑ Привет"]:
        assert oracle_c.query_terms(q) == oracle_np.query_terms(q)


def test_keyword_is_substring_not_token():
    # "we" is found inside "answer"; "did" is not inside "decided" (RecallSearchService.cs:111)
    assert oracle_c.keyword("we", "the answer") == 1.0
    assert oracle_c.keyword("did", "we decided") == 0.0
    assert oracle_c.keyword("azure functions", "Azure Functions rock") == 1.0
    assert oracle_c.keyword("azure k8s", "Azure Functions rock") == 0.5
    assert oracle_c.keyword("azure", "   ") == 0.0
    assert oracle_c.keyword("", "azure") == 0.0


def test_cosine_edge_cases():
    assert oracle_c.cosine([], [1, 2]) == 0.0
    assert oracle_c.cosine([1, 2], None) == 0.0
    assert oracle_c.cosine([1, 2], []) == 0.0
    assert oracle_c.cosine([1, 2], [1, 2, 3]) == 0.0          # length mismatch (:71-72)
    assert oracle_c.cosine([0, 0], [1, 2]) == 0.0             # zero norm (:84-85)
    assert oracle_c.cosine([1, 0], [1, 0]) == 1.0
    assert math.isnan(oracle_c.cosine([1, float("nan")], [1, 1]))   # NaN propagates
    assert oracle_c.cosine([0.2, 0.8, 0.4], [0.2, 0.8, 0.4]) == 0.9999999999999999


def test_recency():
    assert oracle_c.recency(NOW, NOW) == 1.0
    assert oracle_c.recency(NOW, NOW + DAY) == 1.0            # future-dated clamps to age 0
    assert oracle_c.recency(NOW, NOW - 30 * DAY) == math.exp(-1.0)
    assert oracle_c.recency(NOW, NOW - 45 * DAY) == oracle_np.recency(NOW, NOW - 45 * DAY)


def test_round4_is_bankers_on_scaled_double():
    assert oracle_c.round4(0.30000000000000004) == 0.3
    assert oracle_c.round4(0.00005) == oracle_np.round4(0.00005)
    assert oracle_c.round4(0.12345) == oracle_np.round4(0.12345)
    assert oracle_c.round4(2.5e-4) == 0.0002                   # exactly representable tie -> even


def test_snippet():
    assert oracle_c.snippet("  a\nb\r\n c  ") == "a b   c" == oracle_np.build_snippet("  a\nb\r\n c  ")
    long = "x" * 200
    assert oracle_c.snippet(long) == "x" * 180 + "..." == oracle_np.build_snippet(long)


words = st.sampled_from(["azure", "Azure", "cosmos", "k8s", "the", "what", "is", "helm", "we", "answer", "did",
                         "decided", "choose?", "ÉCOLE", "école", "Привет", "a", "of", "x"])


@settings(max_examples=150, deadline=None)
@given(st.lists(words, min_size=0, max_size=8), st.lists(words, min_size=0, max_size=14), st.sampled_from([" ", "  ", "\t", "\n", " "]))
def test_keyword_c_equals_numpy(qw, cw, sep):
    q, c = sep.join(qw), " ".join(cw)
    assert oracle_c.keyword(q, c) == oracle_np.keyword_score(q, c)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(1, 40), st.sampled_from([2, 3, 16, 129]), st.integers(1, 12),
       st.sampled_from([0, 1, 5, 300]))
def test_search_c_equals_numpy_with_ties(seed, n, dim, top_k, cap):
    """Random small corpora with planted duplicates, zero rows, missing/mismatched embeddings and
    shared timestamps: the two restatements must agree on rows, order and bit-exact scores."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    vocab = ["t%03d" % i for i in range(12)]
    contents, embs, ticks = [], [], []
    for i in range(n):
        kind = rng.integers(0, 10)
        if i > 0 and kind == 0:                       # exact duplicate of an earlier row
            j = int(rng.integers(0, i))
            contents.append(contents[j]); embs.append(embs[j]); ticks.append(ticks[j] if rng.integers(0, 2) else NOW - int(rng.integers(0, 5)) * DAY)
            continue
        contents.append(" ".join(rng.choice(vocab, size=int(rng.integers(0, 6)))))
        if kind == 1:
            embs.append([])                           # embedding failed
        elif kind == 2:
            embs.append([0.0] * dim)                  # zero vector
        elif kind == 3:
            embs.append(list(base[i][: max(1, dim - 1)]))   # another width
        else:
            embs.append(list(base[i]))
        ticks.append(NOW - int(rng.integers(0, 4)) * DAY)
    qvec = list(rng.standard_normal(dim).astype(np.float32)) if rng.integers(0, 4) else []
    query = " ".join(rng.choice(vocab + ["the", "what"], size=int(rng.integers(1, 5))))
    recs = [oracle_np.Chunk(c, e, t, i) for i, (c, e, t) in enumerate(zip(contents, embs, ticks))]
    exp = oracle_np.search(recs, query, qvec, NOW, top_k, cap)
    blob, off = oracle_c.pack_contents(contents)
    emb_off = np.zeros(n + 1, dtype=np.int64)
    emb_off[1:] = np.cumsum([len(e) for e in embs])
    flat = np.array([v for e in embs for v in e] or [0.0], dtype=np.float32)
    rows, scores, tk = oracle_c.search(emb=flat, dim=dim, emb_off=emb_off, ticks=np.array(ticks, dtype=np.int64),
                                       content_blob=blob, content_off=off, query=query,
                                       qvec=np.array(qvec, dtype=np.float32), now_ticks=NOW, top_k=top_k,
                                       candidate_cap=cap)
    assert [int(r) for r in rows] == [e[0] for e in exp]
    assert list(scores) == [e[1] for e in exp]
    assert list(tk) == [e[2] for e in exp]


def test_threads_do_not_change_results():
    rng = np.random.default_rng(7)
    n, dim = 5000, 64
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    ticks = (NOW - rng.integers(0, 365, n) * DAY).astype(np.int64)
    contents = [" ".join("t%03d" % t for t in rng.integers(0, 50, 8)) for _ in range(n)]
    blob, off = oracle_c.pack_contents(contents)
    q = rng.standard_normal(dim).astype(np.float32)
    a = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query="t001 t002 the",
                        qvec=q, now_ticks=NOW, top_k=20, threads=1)
    b = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query="t001 t002 the",
                        qvec=q, now_ticks=NOW, top_k=20, threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_nan_scores_sort_last():
    emb = np.array([[1, 0], [float("nan"), 1], [0, 1]], dtype=np.float32)
    ticks = np.full(3, NOW, dtype=np.int64)
    blob, off = oracle_c.pack_contents(["a", "b", "c"])
    rows, scores, _ = oracle_c.search(emb=emb, dim=2, ticks=ticks, content_blob=blob, content_off=off, query="zzz",
                                      qvec=np.array([1, 0], dtype=np.float32), now_ticks=NOW, top_k=3)
    assert [int(r) for r in rows] == [0, 2, 1] and math.isnan(scores[2])


# ---- the block-streamed form used at BASELINE sizes (oracle/orr_oracle_stream.c) ------------------------------
def test_oracle_generator_equals_the_library_host_generator():
    """oracle_synth_rows / oracle_synth_query (built from csrc/orr_synth.h inside the oracle library, so that the
    reference arm never maps liborr.so) produce the rows and queries of liborr's own host generator bit for bit."""
    from omni_recall_rag_b200 import synth
    for dim, dup in ((768, 5000), (3072, 0), (100, 20000)):
        spec = synth.make_spec(dim, gen_dim=max(dim, 256), dup_row_ppm=dup)
        rows = synth.rows_host(spec, 12345, 300)
        emb, ticks, tids = oracle_c.synth_rows(spec, 12345, 300, threads=3)
        assert np.array_equal(emb, rows.emb) and np.array_equal(ticks, rows.ticks) and np.array_equal(tids, rows.term_ids)
        blob, off = oracle_c.synth_contents(tids)
        blob2, off2 = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
        assert np.array_equal(off, off2) and blob[: off[-1]].tobytes() == blob2[: off2[-1]].tobytes()
        for qi in range(12):
            for nt, fr in ((4, 0), (16, 8), (0, 0)):
                q = synth.query_host(spec, qi, 100_000, n_terms=nt, frequent_terms=fr)
                q2, t2, text2 = oracle_c.synth_query(spec, qi, 100_000, nt, fr)
                assert np.array_equal(q2, q.q) and text2 == q.text
        s2 = oracle_c.synth_spec(dim, gen_dim=max(dim, 256), dup_row_ppm=dup)
        assert all(getattr(s2, f) == getattr(spec, f) for f, _ in oracle_c.SynthSpec._fields_)


@pytest.mark.parametrize("block", [97, 1000, 4096, 100_000])
def test_streamed_oracle_equals_the_oracle_on_the_whole_corpus(block):
    """Block by block with a running top-k == one oracle_search over all rows, ties across block borders included
    (planted duplicate rows share scores and, half of them, timestamps)."""
    from omni_recall_rag_b200 import synth
    from tests.util import oracle_search_synth
    spec = synth.make_spec(256, gen_dim=256, dup_row_ppm=30000)
    n, first = 6000, 1000
    rows = synth.rows_host(spec, first, n)
    qs = [synth.query_host(spec, qi, n, n_terms=nt, frequent_terms=fr) for qi, (nt, fr) in enumerate([(4, 0), (16, 8), (0, 0), (3, 3)])]
    for k in (1, 10, 257):
        res = oracle_c.search_streamed(spec, n, [q.text for q in qs], np.stack([q.q for q in qs]), NOW, k, first_row=first,
                                       block_rows=block, threads=2)
        noemb = oracle_c.search_streamed(spec, n, [q.text for q in qs], None, NOW, k, first_row=first, block_rows=block,
                                         with_emb=False, threads=2)
        for q, (r, s, t), (r0, s0, t0) in zip(qs, res, noemb):
            er, es, et = oracle_search_synth(rows, q, NOW, k)
            assert (r - first).tolist() == er.tolist() and s.tolist() == es.tolist() and t.tolist() == et.tolist()
            q0 = synth.HostQuery(np.zeros(0, np.float32), q.term_ids, q.text, q.terms)
            er, es, et = oracle_search_synth(rows, q0, NOW, k)
            assert (r0 - first).tolist() == er.tolist() and s0.tolist() == es.tolist() and t0.tolist() == et.tolist()
