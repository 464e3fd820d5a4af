#!/usr/bin/env python
"""GPU check of the tcgen05 batched core: raw GEMM scores vs numpy, then orr_search_batch vs the
per-query path.  python tools/batch_check.py [rows] [batch] [dim]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 768
n_terms = int(sys.argv[4]) if len(sys.argv) > 4 else 0
freq = int(sys.argv[5]) if len(sys.argv) > 5 else 0
spec = synth.make_spec(dim)
NOW = spec.now_ticks
sh = orr.RecallShard(dim, rows)
sh.fill_synthetic(spec, 0, rows)
qs = [synth.query_host(spec, i, rows, n_terms=0) for i in range(batch)]
Q = np.stack([q.q for q in qs])
t0 = time.time()
got = sh.debug_batch_scores(Q, NOW, 1)
print("debug gemm call: %.3f s" % (time.time() - t0), got.shape)
hr = synth.rows_host(spec, 0, min(rows, 20000))
E = hr.emb.astype(np.float64)
n = E.shape[0]
en = np.linalg.norm(E, axis=1); qn = np.linalg.norm(Q.astype(np.float64), axis=1)
cos = (Q.astype(np.float64) @ E.T) / np.maximum(qn[:, None] * en[None, :], 1e-300)
cos[:, en == 0] = 0
rec = np.exp(-np.maximum(0, (NOW - hr.ticks) / 864e9) / 30.0)
ref = 0.7 * cos + 0.1 * rec[None, :]
err = np.abs(got[:, :n] - ref)
print("max abs err vs fp64 numpy: %.3e  (mean %.3e)" % (err.max(), err.mean()))
passes = int(os.environ.get("ORR_BATCH_PASSES", "3"))
assert err.max() < (2e-4 if passes == 3 else 6e-3), "GEMM core wrong"
print("pad rows -inf:", np.all(np.isneginf(got[:, rows:])) if got.shape[1] > rows else True)
# full batched search vs single-query path
if n_terms:
    qs = [synth.query_host(spec, i, rows, n_terms=n_terms, frequent_terms=freq) for i in range(batch)]
terms = [q.terms for q in qs] if n_terms else None
for k in (10, 100):
    sh.search_batch(Q, terms, NOW, k)
    t0 = time.time(); hb = sh.search_batch(Q, terms, NOW, k); tb = time.time() - t0
    tm = sh.last_timing()
    bad = 0
    for b in range(min(batch, 64)):
        h1 = sh.search(Q[b], qs[b].terms if n_terms else orr.QueryTerms.none(), NOW, k)
        if h1.rows.tolist() != hb[b].rows.tolist() or h1.scores.tolist() != hb[b].scores.tolist():
            bad += 1
    qps = batch / tb
    print(f"k={k}: {qps:.0f} QPS; batch call {tb*1e3:.1f} ms, gemm main {tm['scan_ms']:.3f} ms, prep+sample+re-rank {tm['finalize_ms']:.3f} ms, "
          f"redo={tm['n_survivors'] & 0xffff}, mismatches vs single-query path: {bad}")
    assert bad == 0
print("batch check ok")
