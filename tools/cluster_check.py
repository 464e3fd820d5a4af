#!/usr/bin/env python
"""One process, all visible GPUs (orr_cluster_*): parity with the oracle and queries/s at 1M rows per GPU.
python tools/cluster_check.py [rows_per_gpu] [n_queries]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth
from tests.util import assert_same_ranking, oracle_search_synth

per = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ng = torch.cuda.device_count()
dim, k = 3072, 10
spec = synth.make_spec(dim)
NOW = spec.now_ticks
# parity on a small corpus first
small = 5_000
with orr.RecallCluster(dim, small, list(range(ng)), max_top_k=32) as cl:
    cl.fill_synthetic(spec, 0, small)
    rows = synth.rows_host(spec, 0, small * ng)
    for qi in range(6):
        q = synth.query_host(spec, qi, small * ng, n_terms=4)
        got = cl.search(q.q, q.terms, NOW, k)
        er, es, _ = oracle_search_synth(rows, q, NOW, k)
        assert_same_ranking(got.rows, got.scores, [((int(r) // small) << 40) | (int(r) % small) for r in er], es, what=f"q={qi}")
print(f"cluster parity ok on {ng} GPU(s)")
with orr.RecallCluster(dim, per, list(range(ng)), max_top_k=32) as cl:
    cl.fill_synthetic(spec, 0, per)
    qs = [synth.query_host(spec, qi, per * ng, n_terms=4) for qi in range(nq + 20)]
    for q in qs[:20]:
        cl.search(q.q, q.terms, NOW, k)
    t0 = time.perf_counter()
    for q in qs[20:]:
        cl.search(q.q, q.terms, NOW, k)
    dt = time.perf_counter() - t0
    print(f"single process, {ng} GPU(s) x {per} rows x {dim}: {nq / dt:.1f} queries/s over {per * ng} rows "
          f"({1000 * dt / nq:.3f} ms/query, host buffers in and out) = {nq / dt * per * ng / 1e6:.1f} x (1M-row scans/s)")
