#!/usr/bin/env python
"""Bulk ingest (orr_store_upsert_documents_texts): 1 M chunks loaded in calls of 5 000 documents x 10 chunks
while another thread keeps searching; prints the load time and the search latency (idle vs during the load).
python tools/ingest_check.py [chunks] [dim]"""
import os, statistics, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
per_doc, docs_per_call = 10, 5_000
spec = synth.make_spec(dim)
NOW = spec.now_ticks
rng = np.random.default_rng(5)
chunks_per_call = per_doc * docs_per_call
# one call's worth of host data, reused (new document keys each call): embeddings, ticks, Content of 64 words
emb = rng.standard_normal((chunks_per_call, dim)).astype(np.float32)
ticks = (NOW - rng.integers(0, 365, chunks_per_call) * 864_000_000_000).astype(np.int64)
vocab = [synth.term_text(int(t)) for t in rng.integers(0, 200_000, 4096)]
contents = [" ".join(vocab[int(j)] for j in rng.integers(0, 4096, 64)) for _ in range(chunks_per_call)]
counts = [per_doc] * docs_per_call
with orr.RecallShard(dim, total + chunks_per_call, term_slots=64) as sh:
    sh.upsert_documents_texts(np.arange(docs_per_call, dtype=np.uint64) + 1, counts, emb, ticks, contents)   # something to search
    q = rng.standard_normal(dim).astype(np.float32)
    terms = orr.QueryTerms(2, np.array([orr.hash_term(vocab[0]), orr.hash_term(vocab[1])], dtype=np.uint64), None)

    def lat(n):
        out = []
        for _ in range(n):
            t0 = time.perf_counter(); sh.search(q, terms, NOW, 10); out.append((time.perf_counter() - t0) * 1000)
        return out
    lat(20)
    idle = lat(200)
    stop, during = threading.Event(), []

    def searcher():
        while not stop.is_set():
            t0 = time.perf_counter(); sh.search(q, terms, NOW, 10); during.append((time.perf_counter() - t0) * 1000)
    th = threading.Thread(target=searcher); th.start()
    t0 = time.perf_counter()
    done, key = chunks_per_call, docs_per_call + 1
    while done < total:
        sh.upsert_documents_texts(np.arange(docs_per_call, dtype=np.uint64) + key, counts, emb, ticks, contents)
        key += docs_per_call; done += chunks_per_call
    dt = time.perf_counter() - t0
    stop.set(); th.join()
    p = lambda xs, f: sorted(xs)[min(len(xs) - 1, int(f * len(xs)))]
    print(f"bulk ingest: {done - chunks_per_call} chunks x {dim} (64 words each) in {dt:.2f} s = {(done - chunks_per_call) / dt / 1e3:.0f} k chunks/s "
          f"({sh.count} live rows, vocabulary {sh.vocab_size}); search idle median {statistics.median(idle):.3f} / p99 {p(idle, .99):.3f} ms at "
          f"{chunks_per_call} rows; during the load ({len(during)} searches, store growing to {done} rows) median "
          f"{statistics.median(during):.3f} / p99 {p(during, .99):.3f} ms; final-size idle median {statistics.median(lat(100)):.3f} ms")
