"""Randomised differential test of the whole drop-in path against the C oracle: random documents (chunks without an
embedding, with an embedding of another width, zero vectors, duplicated vectors, equal timestamps, Unicode and mixed-case
words), a random history of replace / delete / bulk ingest / compaction, then random query strings (substring terms, stop
words, blanks between terms, no embedding) at random top_k through GpuRecallSearchService — the reference's own boundary,
RecallSearchService.SearchAsync :20-57 over InMemoryIngestionStore's semantics (:17-25, :50-65).

The oracle sees the LIVE chunks in the store's row order (the insertion order the reference's stable sort falls back to);
scores must agree to 1e-12 relative, ids and order exactly outside near-tie groups, citation scores are Math.Round(.., 4)."""
import numpy as np
import pytest

from omni_recall_rag_b200 import recall as R
from omni_recall_rag_b200 import store as S
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS
SYLL = ["ai", "go", "ra", "ne", "ml", "to", "ka", "zu", "Re", "mi", "lo", "XY", "qu", "en", "st", "Çe", "ß", "Ω", "я", "Ж", "İ", "ǅ"]
STOP = ["the", "what", "is", "of", "and", "How", "WHERE"]


class _Emb:
    def __init__(self):
        self.table = {}

    def embed(self, text):
        v = self.table.get(text)
        return R.EmbeddingResult(v if v is not None else [], "Success" if v is not None else "Empty")


def _oracle_inputs(st, dim):
    """The live chunks in row order -> ragged embeddings, ticks, contents, chunk ids."""
    items = sorted(st._chunk_by_row.items())
    embs, off, ticks, contents, ids = [], [0], [], [], []
    for _, c in items:
        e = c.embedding
        e = np.asarray(e, dtype=np.float32) if e is not None else np.zeros(0, np.float32)
        embs.append(e)
        off.append(off[-1] + len(e))
        ticks.append(c.created_at_utc)
        contents.append(c.content or "")
        ids.append(c.id)
    flat = np.concatenate(embs) if embs and off[-1] else np.zeros(1, np.float32)
    return flat, np.asarray(off, np.int64), np.asarray(ticks, np.int64), contents, ids, [r for r, _ in items]


@pytest.mark.parametrize("seed", range(40))
def test_random_store_history_and_queries_match_the_oracle(seed):
    rng = np.random.default_rng(4200 + seed)
    dim = int(rng.choice([4, 32, 64, 128, 260]))
    slots = int(rng.choice([32, 64, 128]))
    keep_text = bool(seed & 1)
    vocab = ["".join(rng.choice(SYLL, size=rng.integers(1, 5))) for _ in range(int(rng.integers(50, 600)))]
    n_docs = int(rng.integers(20, 90))
    pool = rng.standard_normal((64, dim)).astype(np.float32)          # vectors chunks may share (exact duplicates -> ties)

    def make_doc(d, version):
        t_doc = NOW - int(rng.integers(0, 90)) * DAY
        chunks = []
        for j in range(int(rng.integers(1, 8))):
            r = rng.random()
            if r < 0.10:
                e = None                                               # no embedding
            elif r < 0.15:
                e = rng.standard_normal(dim + 4).astype(np.float32).tolist()   # another width -> cosine 0 (:71-72)
            elif r < 0.20:
                e = np.zeros(dim, np.float32).tolist()                 # zero norm -> cosine 0 (:84-85)
            elif r < 0.45:
                e = pool[int(rng.integers(0, len(pool)))].tolist()     # shared vector
            else:
                e = rng.standard_normal(dim).astype(np.float32).tolist()
            words = rng.choice(vocab, size=int(rng.integers(1, min(25, slots - 1))))
            sep = [" ", "  ", "\t", "\n"][int(rng.integers(0, 4))]
            t = t_doc if rng.random() < 0.7 else NOW - int(rng.integers(0, 90)) * DAY - int(rng.integers(0, 1000))
            chunks.append(S.CosmosChunkRecord(id=f"d{d}v{version}:{j:04d}", document_id=f"d{d}", chunk_index=j,
                                              content=sep.join(words), embedding=e, created_at_utc=int(t)))
        return chunks

    st = S.GpuIngestionStore(dim, 4096, term_slots=slots, keep_text=keep_text)
    emb_client = _Emb()
    try:
        first = [make_doc(d, 0) for d in range(n_docs)]
        half = n_docs // 2
        for b in first[:half]:
            st.upsert_chunks(b)
        st.upsert_chunks_bulk(first[half:])                            # bulk ingest = the same store
        for step in range(int(rng.integers(5, 25))):                   # a random history
            d = int(rng.integers(0, n_docs))
            op = rng.random()
            if op < 0.45:
                st.upsert_chunks(make_doc(d, step + 1))                # replace-by-document
            elif op < 0.75:
                st.delete_document(f"d{d}")
            elif op < 0.90:
                st.upsert_chunks_bulk([make_doc(x, 100 + step) for x in set(int(y) for y in rng.integers(0, n_docs, 3))])
            else:
                st.compact()
        flat, off, ticks, contents, ids, rows = _oracle_inputs(st, dim)
        if not ids:
            pytest.skip("the random history deleted everything")
        blob, coff = oracle_c.pack_contents(contents)
        row_to_pos = {r: i for i, r in enumerate(rows)}
        for qi in range(14):
            nt = int(rng.integers(0, 5))
            words = []
            for _ in range(nt):
                w = str(rng.choice(vocab))
                r = rng.random()
                words.append(w[: max(1, len(w) // 2)] if r < 0.3 else w.upper() if r < 0.5 else w[1:] if r < 0.6 and len(w) > 1 else w)
            words += [str(x) for x in rng.choice(STOP, size=int(rng.integers(0, 3)))]
            rng.shuffle(words)
            q = ("  " if rng.random() < 0.3 else " ").join(words) or "zz-nothing"
            with_vec = rng.random() < 0.8
            qv = rng.standard_normal(dim).astype(np.float32) if with_vec else np.zeros(0, np.float32)
            if with_vec and rng.random() < 0.3:
                qv = pool[int(rng.integers(0, len(pool)))].copy()      # exactly a stored vector: cosine 1 ties
            emb_client.table = {q: qv.tolist()} if with_vec else {}
            k = int(rng.choice([1, 3, 10, 40, 300]))
            cap = int(rng.choice([0, 300, 7]))
            er, es, _ = oracle_c.search(emb=flat, emb_off=off, dim=dim, ticks=ticks, content_blob=blob, content_off=coff, query=q,
                                        qvec=qv, now_ticks=NOW, top_k=k, candidate_cap=cap)
            svc = R.GpuRecallSearchService(st, emb_client, candidate_cap=cap, clock=lambda: NOW, keyword_mode="auto")
            try:
                resp = svc.search(q, k)
            except R.UnsupportedQueryError:
                assert not keep_text                                    # only a store without text may refuse (too many probes)
                continue
            what = f"seed={seed} q={q!r} k={k} cap={cap} vec={with_vec} dim={dim} slots={slots}"
            assert [c.score for c in resp.citations] == [oracle_c.round4(x) for x in es], what
            hits = st.shard.search_query(q, qv, NOW, k, candidate_cap=cap)
            assert_same_ranking([row_to_pos[int(r)] for r in hits.rows], hits.scores, er, es, what=what)
            got_ids = [c.chunk_id for c in resp.citations]
            if [row_to_pos[int(r)] for r in hits.rows] == [int(r) for r in er]:
                assert got_ids == [ids[int(r)] for r in er], what
    finally:
        st.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_batches_match_the_oracle(seed):
    """The batched tcgen05 path on random shapes: width, rows (not a multiple of the 256-row tile), batch size (one to three
    query blocks, padded), top_k up to the path's 128, 0..16 terms MIXED inside one batch, frequent terms, planted duplicate
    rows, every screen setting (auto / bf16 / bf16x3), zero and vanishing-norm query vectors — every query against the oracle."""
    import omni_recall_rag_b200 as orr
    from tests.util import oracle_search_synth

    rng = np.random.default_rng(9100 + seed)
    dim = int(rng.choice([64, 128, 192, 768]))
    n = int(rng.integers(300, 12_000))
    B = int(rng.choice([8, 9, 31, 64, 257, 300, 520]))
    k = int(rng.choice([1, 10, 50, 128]))
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=int(rng.choice([8, 16, 32])), dup_row_ppm=int(rng.choice([0, 20000, 200000])))
    rows = synth.rows_host(spec, 0, n)
    qs = []
    for b in range(B):
        nt = int(rng.integers(0, 17))
        q = synth.query_host(spec, 7000 * seed + b, n, n_terms=nt, frequent_terms=int(rng.integers(0, nt + 1)))
        r = rng.random()
        if r < 0.05:
            q = synth.HostQuery(np.zeros(dim, np.float32), q.term_ids, q.text, q.terms)                 # zero query: cosine 0
        elif r < 0.10:
            q = synth.HostQuery((q.q * np.float32(1e-22)).astype(np.float32), q.term_ids, q.text, q.terms)   # norm leaves fp32
        qs.append(q)
    with orr.RecallShard(dim, n + 16, term_slots=32) as sh:
        sh.fill_synthetic(spec, 0, n)
        sh.set_option("batch_passes", int(rng.choice([0, 1, 3])))
        got = sh.search_batch(np.stack([q.q for q in qs]), [q.terms for q in qs], NOW, k)
        assert len(got) == B
        for b in rng.choice(B, size=min(B, 24), replace=False):
            er, es, et = oracle_search_synth(rows, qs[int(b)], NOW, k)
            assert_same_ranking(got[int(b)].rows, got[int(b)].scores, er, es, what=f"seed={seed} b={b} dim={dim} n={n} B={B} k={k}")
