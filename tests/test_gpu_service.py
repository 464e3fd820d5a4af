"""The service-level path — query STRING in, citations out — through the reference-shaped host classes and the
text-level C ABI (orr_store_upsert_document_texts, orr_search_query): tokenising, stop words, the live vocabulary in
HBM and its GPU substring expansion, then the fused scan.  Needs a B200 (-m gpu).

Reference: RecallSearchService.SearchAsync :20-57 with KeywordScore :90-113 (Contains is a SUBSTRING test) and
InMemoryIngestionStore's replace / delete semantics (:17-25, :50-55).  Scores within 1e-12 relative of the oracle,
ids and order exact outside near-tie groups (tests/util.py); citation scores are Math.Round(score, 4) and must be equal."""
import os

import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import recall as R
from omni_recall_rag_b200 import store as S
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS


class _Emb:
    def __init__(self, table):
        self.table = table

    def embed(self, text):
        v = self.table.get(text)
        return R.EmbeddingResult(v if v is not None else [], "Success" if v is not None else "Empty")


def _ingest(st, contents, emb, ticks, per_doc=5):
    for d0 in range(0, len(contents), per_doc):
        doc_id = f"doc{d0 // per_doc}"
        st.upsert_document(S.CosmosDocumentRecord(id=doc_id, file_name=f"{doc_id}.md", created_at_utc=int(ticks[d0])))
        st.upsert_chunks([S.CosmosChunkRecord(id=f"{doc_id}:{j:04d}", document_id=doc_id, chunk_index=j, content=contents[d0 + j],
                                              embedding=None if emb is None else emb[d0 + j].tolist(), created_at_utc=int(ticks[d0 + j]))
                          for j in range(min(per_doc, len(contents) - d0))])


def test_vocabulary_tracks_replace_and_delete():
    """A word matches only while a live chunk holds it: replace-by-document and delete release the old words."""
    st = S.GpuIngestionStore(4, 64, term_slots=32)
    try:
        t = NOW - DAY
        mk = lambda d, texts: st.upsert_chunks([S.CosmosChunkRecord(id=f"{d}:{i:04d}", document_id=d, chunk_index=i, content=x,
                                                                    embedding=None, created_at_utc=t) for i, x in enumerate(texts)])
        mk("a", ["alpha beta gamma", "beta delta"])
        mk("b", ["Gamma epsilon", "zeta"])
        assert st.vocabulary_size == 6                                  # alpha beta gamma delta epsilon zeta
        svc = R.GpuRecallSearchService(st, R.NoOpEmbeddingClient(), clock=lambda: NOW)
        kw = lambda q: sorted(c.chunk_id for c in svc.search(q, 10).citations if c.score > 0.1)
        assert kw("gam") == ["a:0000", "b:0000"]                        # substring of "gamma", case folded
        assert kw("ta") == ["a:0000", "a:0001", "b:0001"]               # beta, delta, zeta
        mk("a", ["omega"])                                              # replace: alpha beta gamma delta leave with the old rows
        assert st.vocabulary_size == 4                                  # gamma epsilon zeta omega
        assert kw("beta") == [] and kw("gam") == ["b:0000"] and kw("meg") == ["a:0000"]
        st.delete_document("b")
        assert st.vocabulary_size == 1 and kw("gam") == [] and kw("zeta") == []
        qt, n_probes = st.shard.expand_query("what is the OMEGA gamma")
        assert qt.n_terms == 2 and n_probes == 1                        # stop words dropped; "gamma" has no live word
    finally:
        st.close()


@pytest.mark.parametrize("keep_text", [True, False])
def test_service_matches_oracle_on_natural_text_with_substring_terms(keep_text):
    rng = np.random.default_rng(101)
    syll = ["ai", "go", "ra", "ne", "ml", "to", "ka", "zu", "Re", "mi", "lo", "XY", "qu", "en", "st", "Çe", "ß"]
    vocab = ["".join(rng.choice(syll, size=rng.integers(2, 6))) for _ in range(3000)]
    n, dim = 1200, 32
    contents = [" ".join(rng.choice(vocab, size=30)) for _ in range(n)]
    ticks = (NOW - rng.integers(0, 60, size=n) * DAY).astype(np.int64)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    queries = [vocab[5], vocab[7].upper() + " " + vocab[9], "What is the " + vocab[11][:5], "zz-never", "the of and",
               vocab[100][1:] + "  " + vocab[101][:-1] + "\t" + vocab[102], "çe" + vocab[3][:2]]
    qtable = {q: rng.standard_normal(dim).astype(np.float32).tolist() for q in queries}
    st = S.GpuIngestionStore(dim, n + 8, term_slots=32, keep_text=keep_text)
    try:
        _ingest(st, contents, emb, ticks)
        blob, off = oracle_c.pack_contents(contents)
        for cap in (300, 0):
            svc = R.GpuRecallSearchService(st, _Emb(qtable), candidate_cap=cap, clock=lambda: NOW, keyword_mode="auto")
            for q in queries:
                er, es, _ = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query=q,
                                            qvec=np.asarray(qtable[q], dtype=np.float32), now_ticks=NOW, top_k=10, candidate_cap=cap)
                try:
                    resp = svc.search(q, 10)
                except R.UnsupportedQueryError:
                    assert not keep_text                                # only a store without text may refuse (too many probes)
                    continue
                assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es], (q, cap)
                hits = st.shard.search_query(q, np.asarray(qtable[q], dtype=np.float32), NOW, 10, candidate_cap=cap)
                assert_same_ranking(hits.rows, hits.scores, er, es, what=f"service q={q!r} cap={cap}")
    finally:
        st.close()


def test_chunks_with_more_tokens_than_slots_go_through_text_mode_or_are_refused_cleanly():
    rng = np.random.default_rng(7)
    words = [f"tok{i}" for i in range(400)]
    long_chunk = " ".join(words[:100])                                 # 100 distinct tokens, the store has 32 slots
    normal = [" ".join(rng.choice(words, size=10)) for _ in range(40)]
    ticks = np.full(41, NOW - DAY, dtype=np.int64)
    contents = normal + [long_chunk]
    blob, off = oracle_c.pack_contents(contents)
    st = S.GpuIngestionStore(4, 64, term_slots=32, keep_text=True)
    try:
        _ingest(st, contents, None, ticks, per_doc=41)
        svc = R.GpuRecallSearchService(st, R.NoOpEmbeddingClient(), candidate_cap=0, clock=lambda: NOW)
        for q in ("tok99", "tok5 tok77 tok399", "ok9"):                # tok99 sits beyond the 32nd slot of the long chunk
            resp = svc.search(q, 41)
            assert st.shard.last_timing()["path"] & 0xff == N.PATH_TEXT
            er, es, _ = oracle_c.search(emb=None, dim=0, ticks=ticks, content_blob=blob, content_off=off, query=q,
                                        qvec=np.zeros(0, np.float32), now_ticks=NOW, top_k=41)
            assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es], q
        st.upsert_chunks([S.CosmosChunkRecord(id="doc0:0000", document_id="doc0", chunk_index=0, content="tok1 tok2",
                                              embedding=None, created_at_utc=NOW - DAY)])   # the over-long chunk is replaced
        svc.search("tok1", 3)
        assert st.shard.last_timing()["path"] & 0xff != N.PATH_TEXT   # hashed probes again
    finally:
        st.close()
    st = S.GpuIngestionStore(4, 64, term_slots=32, keep_text=False)
    try:
        with pytest.raises(N.OrrError) as e:
            _ingest(st, [long_chunk], None, ticks[:1])
        assert e.value.code == N.ORR_E_UNSUPPORTED
        assert st.shard.count == 0 and st.vocabulary_size == 0          # nothing was changed
    finally:
        st.close()


def test_service_over_the_synthetic_bench_corpus_with_a_1m_word_vocabulary():
    """What bench.py's e2e_service leg runs: the store filled on the device, the 2^20-token vocabulary registered, query
    strings expanded on the GPU, citations rebuilt from the generator — against the streamed oracle."""
    n, dim, k = 150_000, 768, 10
    spec = synth.make_spec(dim)
    st = S.GpuIngestionStore(dim, n, term_slots=64, keep_text=False)
    try:
        st.fill_synthetic(spec, n)
        assert st.vocabulary_size == 1 << 20
        qs = [synth.query_host(spec, qi, n, n_terms=4) for qi in range(6)]
        exp = oracle_c.search_streamed(spec, n, [q.text for q in qs], np.stack([q.q for q in qs]), NOW, k, block_rows=50_000)
        svc = R.GpuRecallSearchService(st, _Emb({q.text: q.q.tolist() for q in qs}), candidate_cap=0, clock=lambda: NOW)
        for q, (er, es, et) in zip(qs, exp):
            resp = svc.search(q.text, k)
            assert st.shard.last_timing()["path"] == N.PATH_FUSED
            assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es]
            assert [c.created_at_utc for c in resp.citations] == et.tolist()
            for c, r in zip(resp.citations, er):
                first = int(c.document_id.split("-")[1])
                assert first + c.chunk_index == int(r)
                one = synth.rows_host(spec, int(r), 1, want_emb=False)
                assert c.snippet.startswith(synth.contents_of(one.term_ids)[0][:180])
            qt, n_probes = st.shard.expand_query(q.text)
            assert qt.n_terms == 4 and n_probes == 4                    # fixed-width tokens: the expansion is the identity
            assert sorted(qt.probe_hash.tolist()) == sorted(q.terms.probe_hash.tolist())
    finally:
        st.close()


def test_snapshot_carries_the_vocabulary_and_corrupt_snapshots_are_refused(tmp_path):
    rng = np.random.default_rng(3)
    words = [f"w{i}x" for i in range(200)]
    contents = [" ".join(rng.choice(words, size=12)) for _ in range(60)]
    ticks = (NOW - rng.integers(0, 9, size=60) * DAY).astype(np.int64)
    emb = rng.standard_normal((60, 8)).astype(np.float32)
    st = S.GpuIngestionStore(8, 128, term_slots=32)
    d = str(tmp_path / "snap")
    try:
        _ingest(st, contents, emb, ticks)
        st.delete_document("doc3")
        svc = R.GpuRecallSearchService(st, R.NoOpEmbeddingClient(), candidate_cap=0, clock=lambda: NOW)
        before = [(c.chunk_id, c.score) for c in svc.search("w17 w4x 9x", 20).citations]
        vocab_before = st.vocabulary_size
        st.save(d)
    finally:
        st.close()
    assert not os.path.exists(os.path.join(d, "host.pkl"))             # plain data only
    st2 = S.GpuIngestionStore(8, 128, term_slots=32)
    try:
        st2.load(d)
        assert st2.vocabulary_size == vocab_before
        svc2 = R.GpuRecallSearchService(st2, R.NoOpEmbeddingClient(), candidate_cap=0, clock=lambda: NOW)
        assert [(c.chunk_id, c.score) for c in svc2.search("w17 w4x 9x", 20).citations] == before
        st2.delete_document("doc1")                                     # the document -> words table came back too
        assert st2.vocabulary_size <= vocab_before
    finally:
        st2.close()
    # corrupt files: truncated, garbage header counts, wrong row_base — an error code, never a crash or a half-loaded store
    raw = open(os.path.join(d, "shard.orrsnap"), "rb").read()
    cases = {"truncated": raw[: len(raw) // 2], "short": raw[:40], "docs": raw[:64] + b"\xff" * 8 + raw[72:],
             "rows": raw[:24] + (10 ** 12).to_bytes(8, "little") + raw[32:]}
    for name, data in cases.items():
        p = str(tmp_path / f"bad_{name}.orrsnap")
        open(p, "wb").write(data)
        with orr.RecallShard(8, 128, term_slots=32) as sh:
            sh.set_option("keep_text", 1)
            with pytest.raises(N.OrrError) as e:
                sh.load(p)
            assert e.value.code in (N.ORR_E_INVALID, N.ORR_E_OOM), name
            assert sh.count == 0 and sh.rows_used == 0
    with orr.RecallShard(8, 128, term_slots=32, row_base=1 << 40) as sh:
        with pytest.raises(N.OrrError):
            sh.load(os.path.join(d, "shard.orrsnap"))


def test_bulk_ingest_equals_document_by_document_ingest_and_searches_run_meanwhile():
    """orr_store_upsert_documents_texts (warm load / hydration) against the per-document path: same rows,
    same vocabulary, same search results — including replacement of documents that already exist — while another thread
    keeps searching (it must only ever see published rows: a well-formed, ordered list every time)."""
    import threading

    rng = np.random.default_rng(77)
    words = [f"w{i:04d}" for i in range(3000)]
    dim, n_docs = 64, 400
    docs = {}
    for d in range(n_docs):
        nch = int(rng.integers(1, 9))
        t = NOW - int(rng.integers(0, 200)) * DAY
        docs[f"doc{d}"] = [S.CosmosChunkRecord(id=f"doc{d}:{j:04d}", document_id=f"doc{d}", chunk_index=j,
                                              content=" ".join(rng.choice(words, size=int(rng.integers(5, 40)))),
                                              embedding=None if rng.random() < 0.1 else rng.standard_normal(dim).astype(np.float32).tolist(),
                                              created_at_utc=t) for j in range(nch)]
    second = {k: [S.CosmosChunkRecord(id=c.id, document_id=c.document_id, chunk_index=c.chunk_index, content=c.content + " fresh",
                                      embedding=c.embedding, created_at_utc=c.created_at_utc + DAY) for c in v[: max(1, len(v) - 1)]]
              for k, v in list(docs.items())[:150]}                          # 150 documents are replaced (with fewer chunks)
    a = S.GpuIngestionStore(dim, 8192, term_slots=64)
    b = S.GpuIngestionStore(dim, 8192, term_slots=64)
    try:
        for v in docs.values():
            a.upsert_chunks(v)
        for v in second.values():
            a.upsert_chunks(v)
        stop, errors = threading.Event(), []

        def searcher():
            qv = rng.standard_normal(dim).astype(np.float32)
            try:
                while not stop.is_set():
                    h = b.shard.search(qv, orr.QueryTerms.none(), NOW, 10)
                    assert np.all(np.diff(h.scores) <= 0) and len(set(h.rows.tolist())) == len(h)
            except Exception as e:                                     # noqa: BLE001
                errors.append(e)

        b.upsert_chunks_bulk(list(docs.values())[:50])                     # something to search from the start
        th = threading.Thread(target=searcher)
        th.start()
        try:
            b.upsert_chunks_bulk(list(docs.values())[50:])
            b.upsert_chunks_bulk(list(second.values()))
        finally:
            stop.set()
            th.join(timeout=60)
        assert not errors, errors[:1]
        assert a.shard.count == b.shard.count and a.vocabulary_size == b.vocabulary_size
        table = {}
        svc_a = R.GpuRecallSearchService(a, _Emb(table), candidate_cap=0, clock=lambda: NOW)
        svc_b = R.GpuRecallSearchService(b, _Emb(table), candidate_cap=0, clock=lambda: NOW)
        for q in ("w0001 w0002 fresh", "w12", "what is the w0999", "nothing-here"):
            table[q] = rng.standard_normal(dim).astype(np.float32).tolist()
            ca, cb = svc_a.search(q, 25).citations, svc_b.search(q, 25).citations
            assert [(c.chunk_id, c.score) for c in ca] == [(c.chunk_id, c.score) for c in cb], q
        with pytest.raises(N.OrrError):                                    # a document twice in one call is refused whole
            b.shard.upsert_documents_texts([5, 5], [1, 1], None, np.array([NOW, NOW]), ["x", "y"])
    finally:
        a.close()
        b.close()
