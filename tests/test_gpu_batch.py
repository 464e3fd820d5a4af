"""Parity of the batched path (tcgen05 contraction + exact re-rank, orr_search_batch) with the
oracle, through the C ABI.  Needs a B200 (-m gpu).

The batched path only SELECTS with tensor-core arithmetic (bf16x3 split precision, or one bf16
pass with a deeper list); every returned score is the exact fp64 re-score, so the tolerance is the
same as the single-query path: scores within SCORE_RTOL = 1e-12 relative of the oracle (contract
1e-5), ids and order exact except inside near-tie groups (tests/util.py)."""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import synth
from tests.util import assert_same_ranking, oracle_search_synth, same_score

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS


def _filled_shard(spec, n, cap=None):
    sh = orr.RecallShard(spec.dim, max(cap or n, 1))
    sh.fill_synthetic(spec, 0, n)
    return sh


def _check_batch(sh, rows, qs, k, what):
    Q = np.stack([q.q for q in qs])
    got = sh.search_batch(Q, [q.terms for q in qs], NOW, k)
    t = sh.last_timing()
    assert t["path"] & 0xff == N.PATH_BATCH, t
    assert len(got) == len(qs)
    for b, q in enumerate(qs):
        er, es, et = oracle_search_synth(rows, q, NOW, k)
        assert_same_ranking(got[b].rows, got[b].scores, er, es, what=f"{what} b={b}")
        assert got[b].ticks.tolist() == et.tolist()
    return t


@pytest.mark.parametrize("passes", [0, 3, 1])
@pytest.mark.parametrize("dim,gen_dim,n,batch,top_k,n_terms,freq", [
    (768, 3072, 30_000, 64, 100, 4, 0),     # C3 in small: truncated embeddings, top-100
    (768, 3072, 20_000, 40, 50, 16, 8),     # C5 in small: 16-term queries, frequent terms, planted duplicates
    (3072, 3072, 5_000, 16, 10, 0, 0),      # full-width rows, no terms
    (128, 128, 2_049, 9, 7, 2, 1),          # one row past a 256-row tile; batch padded 9 -> 256
    (256, 256, 700, 300, 128, 3, 0),        # two query blocks, k at the path's limit, k > live rows / 6
])
def test_batched_path_matches_oracle(passes, dim, gen_dim, n, batch, top_k, n_terms, freq):
    spec = synth.make_spec(dim, gen_dim=gen_dim, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        sh.set_option("batch_passes", passes)
        qs = [synth.query_host(spec, qi, n, n_terms=n_terms, frequent_terms=freq) for qi in range(batch)]
        _check_batch(sh, rows, qs, top_k, f"passes={passes} dim={dim}")


def test_batch_equals_single_query_path_bitwise():
    """Both paths end in the same exact re-score: identical hits, bit for bit."""
    spec = synth.make_spec(768, dup_row_ppm=5000)
    n = 50_000
    with _filled_shard(spec, n) as sh:
        qs = [synth.query_host(spec, qi, n, n_terms=4) for qi in range(128)]
        Q = np.stack([q.q for q in qs])
        for k in (1, 10, 100):
            got = sh.search_batch(Q, [q.terms for q in qs], NOW, k)
            assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
            for b in range(0, 128, 5):
                one = sh.search(qs[b].q, qs[b].terms, NOW, k)
                assert one.rows.tolist() == got[b].rows.tolist()
                assert one.scores.tolist() == got[b].scores.tolist()
                assert one.ticks.tolist() == got[b].ticks.tolist()


def test_mixed_queries_zero_vectors_and_small_batches():
    spec = synth.make_spec(256, gen_dim=256, terms_per_chunk=16, dup_row_ppm=10000)
    n = 6_000
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        qs = [synth.query_host(spec, qi, n, n_terms=(qi % 4), frequent_terms=(qi % 2)) for qi in range(24)]
        qs[3].q[:] = 0.0                                         # zero query: cosine 0 (:84-85)
        qs[7].q[:] = 0.0
        Q = np.stack([q.q for q in qs])
        got = sh.search_batch(Q, [q.terms for q in qs], NOW, 10)
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        for b, q in enumerate(qs):
            er, es, _ = oracle_search_synth(rows, q, NOW, 10)
            assert_same_ranking(got[b].rows, got[b].scores, er, es, what=f"mixed b={b}")
        # fewer than 8 queries run query by query through the fused scan: same answers
        small = sh.search_batch(Q[:3], [q.terms for q in qs[:3]], NOW, 10)
        assert sh.last_timing()["path"] & 0xff != N.PATH_BATCH
        for b in range(3):
            assert small[b].rows.tolist() == got[b].rows.tolist() and small[b].scores.tolist() == got[b].scores.tolist()
        assert sh.search_batch(Q[:0], [], NOW, 10) == []


@pytest.mark.parametrize("passes", [0, 3])
def test_term_counts_from_0_to_16_inside_one_warp(passes):
    """Queries with 0..16 terms side by side (one warp of the epilogue holds 32 of them): the <= 4-term ripple
    counter and the 16-input adder tree must agree with the oracle's match counts, frequent terms included
    (rows matching several terms at once fill the higher count planes)."""
    spec = synth.make_spec(256, gen_dim=256, terms_per_chunk=64)
    n = 9_000
    rows = synth.rows_host(spec, 0, n)
    with _filled_shard(spec, n) as sh:
        sh.set_option("batch_passes", passes)
        qs = [synth.query_host(spec, qi, n, n_terms=(qi % 17), frequent_terms=(qi % 17) // 2) for qi in range(70)]
        _check_batch(sh, rows, qs, 20, f"term counts 0..16 passes={passes}")
        # all queries with exactly 16 frequent terms: rows match several terms at once
        qs = [synth.query_host(spec, 100 + qi, n, n_terms=16, frequent_terms=16) for qi in range(33)]
        _check_batch(sh, rows, qs, 20, f"16 frequent terms passes={passes}")


def test_rows_matching_up_to_all_16_terms_fill_every_count_plane():
    """Planted rows hold 0..16 of a query's 16 terms, so the per-row match count runs through every bit of the
    5 count planes (1, 2, 4, 8, 16) — the carry chain of the adder tree against the oracle's plain count."""
    dim, n = 128, 4096
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=64)
    rows = synth.rows_host(spec, 0, n)
    qs = [synth.query_host(spec, qi, n, n_terms=16, frequent_terms=4) for qi in range(40)]
    tids = rows.term_ids.copy()
    for i in range(0, n, 3):
        q = qs[(i // 3) % len(qs)]
        j = (i // 3) % 17
        tids[i, :j] = q.term_ids[:j]
    rows = synth.HostRows(rows.emb, rows.ticks, tids, rows.doc_first_row)
    contents = synth.contents_of(tids)
    with orr.RecallShard(dim, n) as sh:
        for d in range(0, n, 64):
            got_rows = sh.upsert_document_chunks(d // 64 + 1, rows.emb[d:d + 64], rows.ticks[d:d + 64],
                                                 [orr.tokenize_content(c) for c in contents[d:d + 64]])
            assert got_rows.tolist() == list(range(d, d + 64))
        for passes in (0, 3):
            sh.set_option("batch_passes", passes)
            _check_batch(sh, rows, qs, 30, f"planted 0..16 matches passes={passes}")


def test_batch_sees_mutations_and_rebuilds_its_term_bitmaps():
    """The batched path keeps bf16 planes and per-term row bitmaps beside the store; appends,
    replace-by-document and deletes must show up in the next batch."""
    dim, n0 = 128, 3_000
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16)
    rows = synth.rows_host(spec, 0, n0)
    with _filled_shard(spec, n0, cap=n0 + 64) as sh:
        qs = [synth.query_host(spec, qi, n0, n_terms=2) for qi in range(16)]
        Q = np.stack([q.q for q in qs])
        terms = [q.terms for q in qs]
        before = sh.search_batch(Q, terms, NOW, 5)
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        # a new document whose chunk b is query b itself and holds both of its terms: cosine 1, keyword 1
        th = [np.asarray(q.terms.probe_hash, dtype=np.uint64) for q in qs]
        new_rows = sh.upsert_document_chunks(991, Q.copy(), np.full(16, NOW - DAY, dtype=np.int64), th)
        after = sh.search_batch(Q, terms, NOW, 5)
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        for b in range(16):
            assert after[b].rows[0] == new_rows[b]
            exp = 0.7 * 1.0 + 0.2 * 1.0 + 0.1 * np.exp(-1.0 / 30.0)
            assert abs(after[b].scores[0] - exp) < 1e-6
            one = sh.search(qs[b].q, qs[b].terms, NOW, 5)
            assert one.rows.tolist() == after[b].rows.tolist() and one.scores.tolist() == after[b].scores.tolist()
        sh.delete_document(991)
        again = sh.search_batch(Q, terms, NOW, 5)
        for b in range(16):
            assert again[b].rows.tolist() == before[b].rows.tolist()
            assert again[b].scores.tolist() == before[b].scores.tolist()
            er, es, _ = oracle_search_synth(rows, qs[b], NOW, 5)
            assert_same_ranking(again[b].rows, again[b].scores, er, es, what=f"after delete b={b}")


def test_tie_heavy_batch_orders_by_ticks_then_row():
    """Planted exact duplicates of each query's best row: same score bit for bit, so
    ThenByDescending(CreatedAtUtc) and the stable row order decide (RecallSearchService.cs:34-35)."""
    dim, n, B = 128, 4_000, 32
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    Q = rng.standard_normal((B, dim)).astype(np.float32)
    ticks = np.full(n, NOW - 3 * DAY, dtype=np.int64)
    for b in range(B):
        base = 100 * b + 7
        # scaled by powers of two the products scale exactly: the cosine is bit-identical
        for j in range(6):
            emb[base + 10 * j] = Q[b] * np.float32(2.0 ** (j - 2))
        ticks[base + 20] = NOW - 1 * DAY                           # newer duplicate wins
        ticks[base + 40] = NOW - 9 * DAY                           # older duplicate loses
    from oracle import oracle_c
    blob, off = oracle_c.pack_contents([""] * n)
    with orr.RecallShard(dim, n) as sh:
        sh.upsert_document_chunks(1, emb, ticks)
        got = sh.search_batch(Q, None, NOW, 8)
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        for b in range(B):
            er, es, _ = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query="x",
                                        qvec=Q[b], now_ticks=NOW, top_k=8)
            assert got[b].rows.tolist() == er.tolist(), b
            base = 100 * b + 7
            assert got[b].rows[:6].tolist() == [base + 20, base, base + 10, base + 30, base + 50, base + 40]
            assert all(same_score(x, y) for x, y in zip(got[b].scores, es))


def _properties_at_full_size(n, dim, batch, top_k, n_terms, freq, dup_ppm, sample):
    """Full-size configs, where the oracle cannot run in seconds: (1) the batched hits equal the
    single-query path's (itself pinned to the oracle at small sizes) bit for bit on a sample of the
    batch; (2) the oracle, run on just the returned rows regenerated on the host, reproduces the
    scores; (3) lists are complete, sorted by (score desc, ticks desc, row asc) and idempotent."""
    spec = synth.make_spec(dim, dup_row_ppm=dup_ppm)
    with _filled_shard(spec, n) as sh:
        qs = [synth.query_host(spec, qi, n, n_terms=n_terms, frequent_terms=freq) for qi in range(batch)]
        Q = np.stack([q.q for q in qs])
        terms = [q.terms for q in qs]
        got = sh.search_batch(Q, terms, NOW, top_k)
        t = sh.last_timing()
        assert t["path"] & 0xff == N.PATH_BATCH
        again = sh.search_batch(Q, terms, NOW, top_k)
        for b in range(batch):
            h = got[b]
            assert len(h) == top_k
            assert h.rows.tolist() == again[b].rows.tolist() and h.scores.tolist() == again[b].scores.tolist()
            key = list(zip((-h.scores).tolist(), (-h.ticks).tolist(), h.rows.tolist()))
            assert key == sorted(key), f"b={b}: hits out of reference order"
            assert len(set(h.rows.tolist())) == top_k
        for b in range(0, batch, max(1, batch // sample)):
            one = sh.search(qs[b].q, qs[b].terms, NOW, top_k)
            assert one.rows.tolist() == got[b].rows.tolist(), b
            assert one.scores.tolist() == got[b].scores.tolist(), b
            for r, s in list(zip(got[b].rows, got[b].scores))[:: max(1, top_k // 4)]:
                row = synth.rows_host(spec, int(r), 1)
                er, es, _ = oracle_search_synth(row, qs[b], NOW, 1)
                assert same_score(s, es[0]), (b, int(r), s, es[0])


def test_full_size_properties_c3_5m_x_768_batch_1024_top_100():
    """BASELINE.json configs[2]."""
    _properties_at_full_size(5_000_000, 768, 1024, 100, 4, 0, 0, sample=24)


def test_full_size_properties_c5_keyword_heavy_batch_256_top_50():
    """BASELINE.json configs[4]: 16-term queries, half from the 1000 most frequent tokens, planted
    duplicates (same and different timestamps) for the tie chain."""
    _properties_at_full_size(5_000_000, 768, 256, 50, 16, 8, 1000, sample=24)


def test_auto_mode_cascades_to_split_precision_when_the_screen_cannot_prove_a_query():
    """2000 rows packed within ~2e-3 of each other in score: the bf16 screen's error bound (5.5e-3) cannot
    separate rank k from the survivor boundary, bf16x3 (2e-4) can.  Auto mode must cascade, still return
    the oracle's hits, and then skip the screen for the following batches."""
    dim, n, B, k = 128, 3_000, 32, 10
    rng = np.random.default_rng(17)
    base = rng.standard_normal(dim).astype(np.float32)
    base /= np.linalg.norm(base)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    emb[:2000] = base + 0.1 * rng.standard_normal((2000, dim)).astype(np.float32) / np.sqrt(dim)
    Q = (base + 0.02 * rng.standard_normal((B, dim)) / np.sqrt(dim)).astype(np.float32)
    ticks = np.full(n, NOW - 5 * DAY, dtype=np.int64)
    from oracle import oracle_c
    blob, off = oracle_c.pack_contents([""] * n)
    with orr.RecallShard(dim, n) as sh:
        sh.upsert_document_chunks(1, emb, ticks)
        got = sh.search_batch(Q, None, NOW, k)                       # default = auto
        t = sh.last_timing()
        assert t["path"] == N.PATH_BATCH | N.PATH_ESCALATED, t       # bf16 screen, then bf16x3 for the unproven queries
        assert (t["n_survivors"] & 0xffff) <= B // 4                  # few, if any, needed the single-query path
        for b in range(B):
            er, es, _ = oracle_c.search(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query="x",
                                        qvec=Q[b], now_ticks=NOW, top_k=k)
            assert_same_ranking(got[b].rows, got[b].scores, er, es, what=f"cascade b={b}")
        again = sh.search_batch(Q, None, NOW, k)                      # the screen is on hold: bf16x3 directly
        assert sh.last_timing()["path"] == N.PATH_BATCH
        assert all(again[b].rows.tolist() == got[b].rows.tolist() for b in range(B))
        sh.set_option("batch_passes", 1)                              # screen only: unproven queries run singly, same hits
        single = sh.search_batch(Q, None, NOW, k)
        assert all(single[b].rows.tolist() == got[b].rows.tolist() and single[b].scores.tolist() == got[b].scores.tolist()
                   for b in range(B))


def test_device_resident_answers_equal_the_host_form_and_unsupported_shapes_are_refused():
    """orr_search_batch_device leaves the hits in HBM (row-sharded batches feed them to the all-gather): same bytes as
    orr_search_batch; shapes outside the single-launch tcgen05 path return False instead of half an answer."""
    import torch

    dim, n, B, k = 128, 9_000, 40, 12
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16, dup_row_ppm=20000)
    qs = [synth.query_host(spec, 300 + i, n, n_terms=i % 5) for i in range(B)]
    Q = np.stack([q.q for q in qs])
    terms = [q.terms for q in qs]
    dev = torch.device("cuda", 0)
    with orr.RecallShard(dim, n, term_slots=32) as sh:
        sh.fill_synthetic(spec, 0, n)
        host = sh.search_batch(Q, terms, NOW, k)
        out = torch.zeros(B * k * 24, dtype=torch.uint8, device=dev)
        out_n = torch.zeros(B, dtype=torch.int32, device=dev)
        assert sh.search_batch_device(Q, terms, NOW, k, out.data_ptr(), out_n.data_ptr())
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        raw = np.frombuffer(out.cpu().numpy().tobytes(), dtype=host.raw.dtype).reshape(B, k)
        assert out_n.cpu().tolist() == host.n_out.tolist()
        for b in range(B):
            nb = int(host.n_out[b])
            assert raw[b][:nb].tobytes() == host.raw[b][:nb].tobytes(), b
        assert not sh.search_batch_device(Q[:4], terms[:4], NOW, k, out.data_ptr(), out_n.data_ptr())      # batch < 8
        assert not sh.search_batch_device(Q, terms, NOW, 200, out.data_ptr(), out_n.data_ptr())            # k > 128
