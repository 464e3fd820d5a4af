"""Parity of the CUDA path with the oracle, through the C ABI.  Needs a B200 (-m gpu).

Tolerance: scores within SCORE_RTOL = 1e-12 relative of the oracle's fp64 scores (contract:
1e-5); row ids and order exact except inside near-tie groups (tests/util.py)."""
import math

import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import recall as R
from omni_recall_rag_b200 import store as S
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking, oracle_search_synth, same_score

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS


def _filled_shard(spec, n, **kw):
    sh = orr.RecallShard(spec.dim, max(n, 1), term_slots=kw.pop("term_slots", 64), **kw)
    sh.fill_synthetic(spec, 0, n)
    return sh


# ---- the reference's own fixtures, through the reference-shaped host API -----------------------
class _StubEmbedding:
    def __init__(self, v):
        self.v = v

    def embed(self, text):
        return R.EmbeddingResult(self.v, "Success" if len(self.v) else "Empty", "stub")


def test_reference_fixtures_through_the_drop_in_services(golden):
    """RecallSearchServiceTests.cs / RecallEndpointTests.cs / ChatEndpointTests.cs fixtures,
    run through GpuIngestionStore + GpuRecallSearchService (candidate_cap=300 as the reference)."""
    for case in golden["cases"]:
        qv = case["query_embedding"]
        dim = max(len(qv), max([len(c["embedding"]) for c in case["chunks"]] + [0]), 4)
        dim = (dim + 3) // 4 * 4
        pad = lambda v: list(v) + [0.0] * (dim - len(v)) if len(v) else []
        st = S.GpuIngestionStore(dim, 64)
        try:
            by_doc = {}
            for i, c in enumerate(case["chunks"]):
                by_doc.setdefault(c["document_id"], i)
                st.upsert_document(S.CosmosDocumentRecord(id=c["document_id"], file_name=case["asserted_file_name"] or "f",
                                                          created_at_utc=case["now_ticks"]))
            # the reference fixture inserts all three chunks in ONE UpsertChunksAsync call
            st.upsert_chunks([S.CosmosChunkRecord(id=f'{c["document_id"]}:{c["chunk_index"]:04d}', document_id=c["document_id"],
                                                  chunk_index=c["chunk_index"], content=c["content"],
                                                  embedding=pad(c["embedding"]), created_at_utc=case["now_ticks"])
                              for c in case["chunks"]])
            svc = R.GpuRecallSearchService(st, _StubEmbedding(pad(qv)), clock=lambda: case["now_ticks"])
            resp = svc.search(case["query"], case["top_k"])
            exp = case["derived_hits"]
            assert len(resp.citations) == len(exp), case["name"]
            if case["asserted_first_row"] is not None:   # what the reference test itself asserts
                assert resp.citations[0].document_id == case["chunks"][case["asserted_first_row"]]["document_id"]
            for cit, h in zip(resp.citations, exp):
                assert cit.document_id == case["chunks"][h["row"]]["document_id"], case["name"]
                assert cit.score == h["rounded"], (case["name"], cit.score, h)
            # raw fp64 scores
            hits = st.shard.search(np.array(pad(qv), dtype=np.float32), svc.query_terms(case["query"]),
                                   case["now_ticks"], case["top_k"], candidate_cap=300)
            for s, h in zip(hits.scores, exp):
                assert same_score(s, h["score"]), (case["name"], s, h["score"])
        finally:
            st.close()


def test_blank_query_raises_like_the_reference():
    st = S.GpuIngestionStore(4, 8)
    try:
        svc = R.GpuRecallSearchService(st, R.NoOpEmbeddingClient())
        with pytest.raises(ValueError):
            svc.search("   ", 3)
        assert svc.search("anything", 5).citations == []          # empty store: 200 + no citations
    finally:
        st.close()


# ---- synthetic corpora vs the oracle -----------------------------------------------------------
@pytest.mark.parametrize("dim,gen_dim,n,top_k,n_terms", [
    (3072, 3072, 10_000, 10, 4),      # C1: the reference's CPU-runnable shape
    (768, 3072, 20_000, 10, 4),       # truncated embeddings, norms != 1
    (1536, 1536, 6_000, 5, 2),
    (64, 64, 5_000, 32, 3),           # generic-dim kernel
    (3072, 3072, 333, 10, 0),         # no query terms, ragged last tile
    (100, 100, 1_000, 1, 1),          # dim not a multiple of 128
    (4096, 4096, 3_000, 10, 2),       # wide rows: fewer warps per CTA fit the smem ring
])
def test_fused_path_matches_oracle(dim, gen_dim, n, top_k, n_terms):
    spec = synth.make_spec(dim, gen_dim=gen_dim, dup_row_ppm=2000)
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        for qi in range(6):
            q = synth.query_host(spec, qi, n, n_terms=n_terms)
            got = sh.search(q.q, q.terms, NOW, top_k)
            t = sh.last_timing()
            assert t["path"] & 0xff in (N.PATH_FUSED, N.PATH_EXACT)
            er, es, et = oracle_search_synth(rows, q, NOW, top_k)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"dim={dim} q={qi}")
            assert got.ticks.tolist() == et.tolist()


def test_device_fill_equals_host_generator():
    """The corpus the bench scans is the corpus the oracle sees: device fill == host rows."""
    spec = synth.make_spec(768, dup_row_ppm=5000)
    n = 4096
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        # score every row exactly against a basis-like query and compare with the oracle per row
        q = synth.query_host(spec, 1, n, n_terms=4)
        got = sh.search(q.q, q.terms, NOW, n)                    # k = n -> exact path, all rows
        assert len(got) == n
        blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
        sc, _, _, _ = oracle_c.score_rows(emb=rows.emb, dim=768, ticks=rows.ticks, content_blob=blob,
                                          content_off=off, query=q.text, qvec=q.q, now_ticks=NOW)
        order = np.argsort(got.rows)
        assert got.rows[order].tolist() == list(range(n))
        np.testing.assert_allclose(got.scores[order], sc, rtol=1e-12, atol=1e-15)
        assert got.ticks[order].tolist() == rows.ticks.tolist()


def test_no_embedding_mode_and_tie_chain():
    """The reference's DEFAULT configuration (NoOp embeddings): cosine 0 everywhere, ranking by
    keyword + recency with massive score ties inside documents."""
    spec = synth.make_spec(256, gen_dim=256, terms_per_chunk=16, dup_row_ppm=50000)
    n = 30_000
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        for qi, k in [(0, 10), (1, 50), (2, 300)]:
            q = synth.query_host(spec, qi, n, n_terms=4, frequent_terms=2)
            got = sh.search(None, q.terms, NOW, k)
            assert sh.last_timing()["path"] == N.PATH_EXACT
            q_noemb = synth.HostQuery(np.zeros(0, np.float32), q.term_ids, q.text, q.terms)
            er, es, et = oracle_search_synth(rows, q_noemb, NOW, k)
            assert got.rows.tolist() == er.tolist()               # exact, ties included
            assert all(same_score(a, b) for a, b in zip(got.scores, es))
        # a query of another width scores cosine 0 everywhere (RecallSearchService.cs:71-72)
        q = synth.query_host(spec, 3, n, n_terms=4)
        got = sh.search(q.q[:128], q.terms, NOW, 10)
        q_noemb = synth.HostQuery(np.zeros(0, np.float32), q.term_ids, q.text, q.terms)
        er, es, _ = oracle_search_synth(rows, q_noemb, NOW, 10)
        assert got.rows.tolist() == er.tolist()


def test_duplicate_rows_order_by_ticks_then_row():
    """Planted exact duplicates straddling the top-k: ThenByDescending(CreatedAtUtc) then the
    stable row order decide (RecallSearchService.cs:34-35)."""
    dim, n = 128, 2000
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal(dim).astype(np.float32)
    ticks = np.full(n, NOW - 3 * DAY, dtype=np.int64)
    best = 17
    emb[best] = q * 0.5
    for j, r in enumerate([400, 900, 1500, 1999, 3]):
        emb[r] = emb[best]                                      # same cosine (scale-invariant)
    ticks[900] = NOW - 1 * DAY                                   # newer duplicate wins
    ticks[1500] = NOW - 9 * DAY                                  # older duplicate loses
    with orr.RecallShard(dim, n) as sh:
        sh.upsert_document_chunks(1, emb, ticks)
        got = sh.search(q, orr.QueryTerms.none(), NOW, 8)
        blob, off = oracle_c.pack_contents([""] * n)
        er, es, _ = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query="x",
                                    qvec=q, now_ticks=NOW, top_k=8)
        assert got.rows.tolist() == er.tolist()
        assert got.rows[:6].tolist() == [900, 3, 17, 400, 1999, 1500]


def test_candidate_cap_300_is_the_reference_preselection():
    spec = synth.make_spec(128, gen_dim=128, terms_per_chunk=8)
    n = 5000
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        for qi in range(3):
            q = synth.query_host(spec, qi, n, n_terms=3)
            for cap in (300, 1, 5000, 4096):
                if cap > 4096:
                    continue
                got = sh.search(q.q, q.terms, NOW, 10, candidate_cap=cap)
                assert sh.last_timing()["path"] == N.PATH_SUBSET
                er, es, _ = oracle_search_synth(rows, q, NOW, 10, candidate_cap=cap)
                assert_same_ranking(got.rows, got.scores, er, es, what=f"cap={cap}")


def test_edge_shapes():
    with orr.RecallShard(8, 16) as sh:
        q = np.ones(8, dtype=np.float32)
        assert len(sh.search(q, orr.QueryTerms.none(), NOW, 5)) == 0         # N = 0 -> no hits, no error
        emb = np.eye(8, dtype=np.float32)[:3]
        sh.upsert_document_chunks(7, emb, np.array([NOW, NOW - DAY, NOW - 2 * DAY]))
        got = sh.search(q, orr.QueryTerms.none(), NOW, 10)                   # k > N
        assert got.rows.tolist() == [0, 1, 2]
        assert len(sh.search(q, orr.QueryTerms.none(), NOW, 0)) == 1         # Math.Max(1, topK)
        assert len(sh.search(q, orr.QueryTerms.none(), NOW, -3)) == 1
        with pytest.raises(N.OrrError):
            sh.search(q, orr.QueryTerms(65, np.arange(1, 66, dtype=np.uint64), None), NOW, 1)


def test_nan_rows_rank_last_and_zero_rows_score_zero_cosine():
    dim = 32
    emb = np.zeros((5, dim), dtype=np.float32)
    emb[0, 0] = 1.0
    emb[1, 1] = float("nan")
    emb[2, 0] = 0.5
    ticks = np.full(5, NOW, dtype=np.int64)
    q = np.zeros(dim, dtype=np.float32)
    q[0] = 1.0
    with orr.RecallShard(dim, 8) as sh:
        sh.upsert_document_chunks(1, emb, ticks)
        got = sh.search(q, orr.QueryTerms.none(), NOW, 5)
        assert got.rows.tolist() == [0, 2, 3, 4, 1]
        assert math.isnan(got.scores[4]) and same_score(got.scores[2], 0.1)


def test_mutations_replace_and_delete_by_document():
    """UpsertChunksAsync replaces a document's rows; DeleteDocumentAsync removes them
    (InMemoryIngestionStore.cs:17-25,50-55)."""
    dim = 64
    rng = np.random.default_rng(11)
    q = rng.standard_normal(dim).astype(np.float32)
    with orr.RecallShard(dim, 256) as sh:
        a = rng.standard_normal((10, dim)).astype(np.float32)
        b = rng.standard_normal((10, dim)).astype(np.float32)
        b[4] = q                                                   # the best row lives in doc 2
        sh.upsert_document_chunks(1, a, np.full(10, NOW - DAY))
        rows_b = sh.upsert_document_chunks(2, b, np.full(10, NOW - DAY))
        assert sh.count == 20
        assert sh.search(q, orr.QueryTerms.none(), NOW, 1).rows[0] == rows_b[4]
        sh.delete_document(2)
        assert sh.count == 10
        got = sh.search(q, orr.QueryTerms.none(), NOW, 20)
        assert len(got) == 10 and set(got.rows.tolist()) == set(range(10))
        c = rng.standard_normal((3, dim)).astype(np.float32)
        c[1] = q
        rows_c = sh.upsert_document_chunks(1, c, np.full(3, NOW - DAY))     # replace doc 1
        assert sh.count == 3
        got = sh.search(q, orr.QueryTerms.none(), NOW, 20)
        assert got.rows[0] == rows_c[1] and set(got.rows.tolist()) == set(rows_c.tolist())
        # exact path sees the same live set
        got2 = sh.search(None, orr.QueryTerms.none(), NOW, 20)
        assert set(got2.rows.tolist()) == set(rows_c.tolist())


def test_keyword_substring_expansion_through_the_service():
    """Natural text: 'we' must match 'answer' (substring Contains, :111) — the host expands the
    query term over the live vocabulary into several probes."""
    st = S.GpuIngestionStore(4, 64)
    try:
        t = NOW - DAY
        docs = ["the answer is forty two", "we decided to use azure functions", "nothing relevant here"]
        for i, text in enumerate(docs):
            st.upsert_document(S.CosmosDocumentRecord(id=f"d{i}", file_name=f"f{i}.md", created_at_utc=t))
            st.upsert_chunks([S.CosmosChunkRecord(id=f"d{i}:0000", document_id=f"d{i}", chunk_index=0, content=text,
                                                  embedding=None, created_at_utc=t)])
        svc = R.GpuRecallSearchService(st, R.NoOpEmbeddingClient(), clock=lambda: NOW)
        for query in ["we", "did we answer", "AZURE functions?", "what is the"]:
            resp = svc.search(query, 3)
            for cit in resp.citations:
                content = docs[int(cit.document_id[1:])]
                exp = oracle_c.fuse(0.0, oracle_c.keyword(query, content), oracle_c.recency(NOW, t))
                assert cit.score == oracle_c.round4(exp), (query, content)
    finally:
        st.close()


def test_full_size_properties_1m_x_3072():
    """BASELINE.json configs[1] at full size, where the oracle cannot run in seconds: properties.
    (1) a query planted next to a known row returns that row first with cosine ~ 1;
    (2) the fused path and the exact fp64 path agree on the whole top-10;
    (3) the oracle, run on just the returned rows (regenerated on the host), reproduces the scores;
    (4) results are idempotent."""
    n, dim = 1_000_000, 3072
    spec = synth.make_spec(dim)
    with _filled_shard(spec, n) as sh:
        checked = 0
        for qi in range(40):
            q = synth.query_host(spec, qi, n, n_terms=4)
            got = sh.search(q.q, q.terms, NOW, 10)
            assert sh.last_timing()["path"] == N.PATH_FUSED
            again = sh.search(q.q, q.terms, NOW, 10)
            assert got.rows.tolist() == again.rows.tolist() and got.scores.tolist() == again.scores.tolist()
            assert np.all(np.diff(got.scores) <= 0)
            # oracle on the returned rows only
            for r, s in zip(got.rows[:3], got.scores[:3]):
                one = synth.rows_host(spec, int(r), 1)
                er, es, _ = oracle_search_synth(one, q, NOW, 1)
                assert same_score(s, es[0])
            if qi % 8 == 0:
                exact = sh.search(q.q, q.terms, NOW, 300)          # k > 224 -> exact fp64 path
                assert sh.last_timing()["path"] == N.PATH_EXACT
                assert exact.rows[:10].tolist() == got.rows.tolist()
                assert exact.scores[:10].tolist() == got.scores.tolist()
                checked += 1
        assert checked >= 5


def test_error_codes_and_limits():
    """The C ABI reports bad input as ORR_E_* codes with a message, never by aborting."""
    import ctypes as C

    L = N.lib()
    cfg = N.OrrConfig()
    L.orr_config_default(C.byref(cfg))
    h = C.c_void_p()
    for field, bad, code in [("dim", 6, N.ORR_E_UNSUPPORTED), ("term_slots", 48, N.ORR_E_UNSUPPORTED),
                             ("capacity_rows", 0, N.ORR_E_INVALID), ("abi_version", 99, N.ORR_E_INVALID),
                             ("device", 99, N.ORR_E_CUDA), ("recency_days", 0.0, N.ORR_E_INVALID)]:
        c2 = N.OrrConfig()
        L.orr_config_default(C.byref(c2))
        setattr(c2, field, bad)
        assert L.orr_store_create(C.byref(c2), C.byref(h)) == code, field
        assert L.orr_last_error()
    with orr.RecallShard(8, 4, term_slots=32) as sh:
        emb = np.eye(8, dtype=np.float32)[:3]
        sh.upsert_document_chunks(1, emb, np.array([NOW, NOW, NOW]))
        with pytest.raises(N.OrrError) as e:                         # store full
            sh.upsert_document_chunks(2, emb, np.array([NOW, NOW, NOW]))
        assert e.value.code == N.ORR_E_OOM
        with pytest.raises(N.OrrError) as e:                         # more distinct terms than slots
            sh.upsert_document_chunks(3, emb[:1], np.array([NOW]), [np.arange(1, 40, dtype=np.uint64)])
        assert e.value.code == N.ORR_E_UNSUPPORTED
        with pytest.raises(N.OrrError) as e:                         # reserved ticks value
            sh.upsert_document_chunks(4, emb[:1], np.array([np.iinfo(np.int64).min]))
        assert e.value.code == N.ORR_E_INVALID
        q = np.ones(8, dtype=np.float32)
        with pytest.raises(N.OrrError) as e:
            sh.search(q, orr.QueryTerms(2, np.array([1, 2, 3], dtype=np.uint64), np.array([0, 1, 5], dtype=np.int32)), NOW, 3)
        assert e.value.code == N.ORR_E_INVALID                       # probe_term out of range
        with pytest.raises(N.OrrError):
            sh.search(q, orr.QueryTerms.none(), NOW, 3, candidate_cap=5000)
        with pytest.raises(N.OrrError):
            sh.set_option("no_such_option", 1)
        with pytest.raises(N.OrrError):
            sh.search_text(q, ["x" * 300], NOW, 3)                   # term longer than the substring kernel takes
        assert sh.count == 3 and len(sh.search(q, orr.QueryTerms.none(), NOW, 3)) == 3   # the store survived all of it


def test_shim_call_sequence_replayed_from_plain_c(tmp_path, golden):
    """tests/shim/shim_driver.c dlopen()s liborr.so and makes the calls dotnet/GpuIngestionStore.cs and
    dotnet/GpuRecallSearchService.cs make, in their order, on the reference's own fixture
    (RecallSearchServiceTests.cs:51-117); its output must equal the golden cases derived from that fixture."""
    import json
    import subprocess

    from tests.test_host import _build_shim_driver

    exe = _build_shim_driver(tmp_path)
    out = subprocess.run([exe, "--replay", N.library_path()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = {j["case"]: j for j in map(json.loads, out.stdout.strip().splitlines())}
    assert lines["upsert"]["rows"] == [0, 1, 2] and lines["upsert"]["count"] == 3
    by_query = {c["query"]: c for c in golden["cases"] if len(c["chunks"]) == 3}
    for case in ("vector+keyword", "keyword-only", "stop-words"):
        got = lines[case]
        exp = by_query[got["query"]]["derived_hits"]
        assert [h["row"] for h in got["hits"]] == [h["row"] for h in exp], case
        for g, e in zip(got["hits"], exp):
            assert same_score(g["score"], e["score"]), (case, g, e)
    assert lines["blank"]["rc"] == N.ORR_E_INVALID
    assert lines["empty-store"]["hits"] == []
