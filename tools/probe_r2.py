#!/usr/bin/env python
"""Small GPU probes used while tuning round 2 (not part of the product or the tests).
    python tools/probe_r2.py noemb [rows]     keyword + recency-only searches: scoring kernel / selection / call times
    python tools/probe_r2.py c1               10k x 3072: where a call's time goes (cap = 300 and all rows)
"""
import os, statistics, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth

mode = sys.argv[1] if len(sys.argv) > 1 else "noemb"
if mode == "noemb":
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    dim = 64                                   # the no-embedding path never reads the embeddings
    spec = synth.make_spec(dim, gen_dim=dim)
    with orr.RecallShard(dim, rows, term_slots=64) as sh:
        sh.fill_synthetic(spec, 0, rows)
        qs = [synth.query_host(spec, qi, rows, n_terms=4) for qi in range(40)]
        sc, se, wall = [], [], []
        for i, q in enumerate(qs):
            sh.search(None, q.terms, spec.now_ticks, 10)
            t = sh.last_timing()
            if i >= 10:
                sc.append(t["scan_ms"]); se.append(t["finalize_ms"]); wall.append(t["wall_ms"])
        b = rows * 264
        print(f"noemb {rows} rows: score kernel {statistics.mean(sc)*1000:.1f} us ({b/statistics.mean(sc)/1e6:.0f} GB/s), "
              f"select {statistics.mean(se)*1000:.1f} us, call {statistics.median(wall)*1000:.1f} us")
elif mode == "c1":
    dim, rows = 3072, 10_000
    spec = synth.make_spec(dim)
    with orr.RecallShard(dim, rows, term_slots=64) as sh:
        sh.fill_synthetic(spec, 0, rows)
        qs = [synth.query_host(spec, qi, rows, n_terms=4) for qi in range(300)]
        for cap in (300, 0):
            sc, fi, wall, outer = [], [], [], []
            for i, q in enumerate(qs):
                t0 = time.perf_counter()
                sh.search(q.q, q.terms, spec.now_ticks, 10, candidate_cap=cap)
                t1 = time.perf_counter()
                t = sh.last_timing()
                if i >= 50:
                    sc.append(t["scan_ms"]); fi.append(t["finalize_ms"]); wall.append(t["wall_ms"]); outer.append((t1 - t0) * 1000)
            print(f"c1 cap={cap}: ev0->ev1 {statistics.median(sc)*1000:.1f} us, ev1->ev2 {statistics.median(fi)*1000:.1f} us, "
                  f"C call {statistics.median(wall)*1000:.1f} us (p99 {sorted(wall)[int(.99*len(wall))]*1000:.1f}), python call {statistics.median(outer)*1000:.1f} us")
