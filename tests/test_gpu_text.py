"""Text mode (orr_search_text): the keyword predicate evaluated as an ordinal
substring search on the chunk text kept in HBM, checked against the oracle's KeywordScore restatement
(RecallSearchService.cs:90-113) on natural-looking text.  Needs a B200 (-m gpu).

Tolerance: as everywhere, scores within 1e-12 relative of the oracle, ids and order exact outside ties."""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import recall as R
from omni_recall_rag_b200 import store as S
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking, same_score

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS
SYLL = ["ai", "go", "ra", "ne", "ml", "to", "ka", "zu", "Re", "mi", "lo", "XY", "qu", "en", "st"]


def _corpus(rng, n, words_per_chunk=40, long_every=0, vocab_size=600):
    vocab = ["".join(rng.choice(SYLL, size=rng.integers(1, 5))) for _ in range(vocab_size)]
    contents = []
    for i in range(n):
        if long_every and i % long_every == 0:      # > 4 KB of text: several staging windows (<= 100 distinct words)
            contents.append(" ".join(rng.choice(vocab[:100], size=700)))
        else:
            contents.append(" ".join(rng.choice(vocab, size=words_per_chunk)))
    return vocab, contents


class _Emb:
    def __init__(self, table):
        self.table = table

    def embed(self, text):
        v = self.table.get(text)
        return R.EmbeddingResult(v if v is not None else [], "Success" if v is not None else "Empty")


def _service(contents, emb, ticks, dim, *, mode, cap, qtable):
    st = S.GpuIngestionStore(dim, len(contents) + 8, term_slots=128, text_bytes_per_row=6000)
    per_doc = 7
    for d0 in range(0, len(contents), per_doc):
        doc_id = f"doc{d0 // per_doc}"
        st.upsert_document(S.CosmosDocumentRecord(id=doc_id, file_name=f"{doc_id}.md", created_at_utc=int(ticks[d0])))
        st.upsert_chunks([S.CosmosChunkRecord(id=f"{doc_id}:{j:04d}", document_id=doc_id, chunk_index=j, content=contents[d0 + j],
                                              embedding=None if emb is None else emb[d0 + j].tolist(), created_at_utc=int(ticks[d0 + j]))
                          for j in range(min(per_doc, len(contents) - d0))])
    return st, R.GpuRecallSearchService(st, _Emb(qtable), candidate_cap=cap, clock=lambda: NOW, keyword_mode=mode)


@pytest.mark.parametrize("cap", [300, 0])
@pytest.mark.parametrize("with_emb", [False, True])
def test_text_mode_matches_oracle_on_natural_text(cap, with_emb):
    rng = np.random.default_rng(23)
    n, dim = 1500, 64
    vocab, contents = _corpus(rng, n, long_every=97)
    ticks = (NOW - rng.integers(0, 400, size=n) // 7 * 7 * DAY).astype(np.int64)      # shared timestamps: ties
    emb = rng.standard_normal((n, dim)).astype(np.float32) if with_emb else None
    queries = ["ai", "go ml", "What is the RAne", "x", "zz-never-there", "a", "ra ra RA", "mi lo XY qu en st ka zu",
               vocab[3] + " " + vocab[10], "the of and"]
    qtable = {q: (rng.standard_normal(dim).astype(np.float32).tolist() if with_emb else None) for q in queries}
    st, svc = _service(contents, emb, ticks, dim, mode="text", cap=cap, qtable=qtable)
    try:
        blob, off = oracle_c.pack_contents(contents)
        for q in queries:
            for k in (1, 10, 50):
                got = st.shard.search_text(np.asarray(qtable[q] or [], dtype=np.float32), svc.filtered_terms(q), NOW, k,
                                           candidate_cap=cap)
                assert st.shard.last_timing()["path"] == N.PATH_TEXT
                er, es, et = oracle_c.search(emb=emb, dim=dim if with_emb else 0, ticks=ticks, content_blob=blob, content_off=off,
                                             query=q, qvec=np.asarray(qtable[q] or [], dtype=np.float32), now_ticks=NOW, top_k=k,
                                             candidate_cap=cap)
                assert_same_ranking(got.rows, got.scores, er, es, what=f"text cap={cap} q={q!r} k={k}")
        # through the service: rounded citation scores equal the oracle's
        resp = svc.search("go ml", 5)
        er, es, _ = oracle_c.search(emb=emb, dim=dim if with_emb else 0, ticks=ticks, content_blob=blob, content_off=off,
                                    query="go ml", qvec=np.asarray(qtable["go ml"] or [], dtype=np.float32), now_ticks=NOW,
                                    top_k=5, candidate_cap=cap)
        assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es]
    finally:
        st.close()


def test_auto_mode_uses_the_fused_scan_when_the_expansion_fits_and_text_when_it_does_not():
    rng = np.random.default_rng(29)
    n, dim = 800, 32
    vocab, contents = _corpus(rng, n, vocab_size=4000)
    ticks = np.full(n, NOW - 2 * DAY, dtype=np.int64)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    long_word = max(vocab, key=len)
    queries = {long_word: "hashed", "ai": "text"}                 # "ai" is inside hundreds of vocabulary words
    qtable = {q: rng.standard_normal(dim).astype(np.float32).tolist() for q in queries}
    st, svc = _service(contents, emb, ticks, dim, mode="auto", cap=0, qtable=qtable)
    try:
        blob, off = oracle_c.pack_contents(contents)
        for q, route in queries.items():
            if route == "text":
                with pytest.raises(R.UnsupportedQueryError):
                    svc.query_terms(q)
            resp = svc.search(q, 8)
            path = st.shard.last_timing()["path"]
            assert (path == N.PATH_TEXT) == (route == "text"), (q, path)
            er, es, _ = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query=q,
                                        qvec=np.asarray(qtable[q], dtype=np.float32), now_ticks=NOW, top_k=8)
            assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es], q
        # both modes agree where both apply
        hashed = st.shard.search(np.asarray(qtable[long_word], dtype=np.float32), svc.query_terms(long_word), NOW, 20)
        text = st.shard.search_text(np.asarray(qtable[long_word], dtype=np.float32), svc.filtered_terms(long_word), NOW, 20)
        assert hashed.rows.tolist() == text.rows.tolist() and hashed.scores.tolist() == text.scores.tolist()
    finally:
        st.close()


def test_matches_across_window_boundaries_and_multibyte_text():
    """A term straddling the 1024-byte staging window, at the very end of the text, and UTF-8 multi-byte
    content: byte-level matching of valid UTF-8 is code-point matching (ordinal Contains)."""
    dim = 4
    filler = "ab " * 400                                            # 1200 bytes
    texts = [
        filler[:1019] + "needle" + filler[:50],                    # bytes 1019..1024 straddle the first window
        filler[:2045] + "needle",                                   # at the very end, third window
        "grüße aus münchen — naïve café",                          # multi-byte
        "needl needl eneedl",                                       # near misses only
        "",                                                         # empty content: keyword 0 (:92)
        filler[:1023] + "é" + "needle",                             # a 2-byte char split by the window edge
    ]
    n = len(texts)
    ticks = np.full(n, NOW - DAY, dtype=np.int64)
    with orr.RecallShard(dim, 16) as sh:
        sh.set_option("text_bytes_per_row", 4096)
        sh.upsert_document_chunks(1, None, ticks, None, None, texts_lower=texts)
        for terms, want in [(["needle"], {0, 1, 5}), (["ünch"], {2}), (["é"], {2, 5}), (["naïve", "needle"], {0, 1, 2, 5}),
                            (["eneedl"], {3}), (["zzz"], set())]:
            got = sh.search_text(None, terms, NOW, n)
            rec = oracle_c.recency(NOW, NOW - DAY)
            hit_rows = {int(r) for r, s in zip(got.rows, got.scores) if s > oracle_c.fuse(0.0, 0.0, rec) + 1e-12}
            assert hit_rows == want, (terms, hit_rows)
            for r, s in zip(got.rows, got.scores):
                kw = oracle_c.keyword(" ".join(terms), texts[int(r)])
                assert same_score(s, oracle_c.fuse(0.0, kw, rec)), (terms, int(r))


def test_text_mode_needs_text_for_every_row():
    dim = 4
    with orr.RecallShard(dim, 16) as sh:
        sh.upsert_document_chunks(1, None, np.array([NOW - DAY]), None, None)          # no text
        with pytest.raises(N.OrrError):
            sh.search_text(None, ["x"], NOW, 3)
        with pytest.raises(N.OrrError):                                                  # mixing is refused
            sh.upsert_document_chunks(2, None, np.array([NOW - DAY]), None, None, texts_lower=["abc"])
        assert len(sh.search_text(None, [], NOW, 3)) == 1                                # no terms: nothing to match


def test_synthetic_text_fill_and_prefix_terms_match_the_oracle():
    """The device-side synthetic fill can also write the chunk text (option synth_text): text mode over it
    equals the oracle, including prefix terms that are substrings of thousands of vocabulary tokens, and
    equals the fused hashed path for whole-token terms."""
    dim, n = 256, 20_000
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=24, dup_row_ppm=5000)
    rows = synth.rows_host(spec, 0, n)
    contents = synth.contents_of(rows.term_ids)
    blob, off = oracle_c.pack_contents(contents)
    with orr.RecallShard(dim, n) as sh:
        sh.set_option("synth_text", 1)
        sh.set_option("text_bytes_per_row", 256)
        sh.fill_synthetic(spec, 0, n)
        for qi, terms in enumerate([["t00000"], ["t0000012", "00"], ["7", "t1", "zz"]]):
            q = synth.query_host(spec, qi, n, n_terms=0)
            got = sh.search_text(q.q, terms, NOW, 25)
            er, es, _ = oracle_c.search(emb=rows.emb, dim=dim, ticks=rows.ticks, content_blob=blob, content_off=off,
                                        query=" ".join(terms), qvec=q.q, now_ticks=NOW, top_k=25)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"synthetic text terms={terms}")
        for qi in range(4):
            q = synth.query_host(spec, qi, n, n_terms=4, frequent_terms=2)
            a = sh.search(q.q, q.terms, NOW, 10)
            b = sh.search_text(q.q, q.text.split(), NOW, 10)
            assert a.rows.tolist() == b.rows.tolist() and a.scores.tolist() == b.scores.tolist()
