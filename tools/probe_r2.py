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
elif mode == "c5kw":
    # main-pass time of the batched path at 5M x 768, B = 256 (one query block), as a function of the keyword side:
    # 0 / 4 / 8 / 16 terms per query (frequent = half of them), fresh terms every batch vs the same batch repeated
    dim, rows = 768, int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
    spec = synth.make_spec(dim, dup_row_ppm=1000)
    with orr.RecallShard(dim, rows, term_slots=64) as sh:
        sh.fill_synthetic(spec, 0, rows)
        for B in (256, 512, 1024):
            for nt in (0, 4, 8, 16):
                main, step = [], []
                for it in range(7):
                    qs = [synth.query_host(spec, (it * 4096 + nt * 100000 + B * 7) + j, rows, n_terms=max(nt, 1), frequent_terms=nt // 2) for j in range(B)]
                    terms = [q.terms for q in qs] if nt else [synth.QueryTerms(0, np.zeros(0, np.uint64), None) for _ in qs]
                    for rep in range(2):                      # rep 1: every term bitmap is cached
                        sh.search_batch(np.stack([q.q for q in qs]), terms, spec.now_ticks, 50)
                        t = sh.last_timing()
                        if it >= 2 and rep == 1:
                            main.append(t["scan_ms"]); step.append(t["total_device_ms"])
                print(f"c5kw B={B} terms={nt}: main {statistics.mean(main):.3f} ms, step (warm terms) {statistics.mean(step):.3f} ms")
elif mode == "c5same":
    # does the keyword side cost SM time or memory-system time?  B = 256, 16 terms per query, but every query asks for the
    # SAME 16 terms (their bitmaps are 10 MB: L2-resident) vs 16 different terms per query (1.6 GB of bitmaps)
    dim, rows = 768, 5_000_000
    spec = synth.make_spec(dim, dup_row_ppm=1000)
    with orr.RecallShard(dim, rows, term_slots=64) as sh:
        sh.fill_synthetic(spec, 0, rows)
        B = 256
        for kind in ("same16", "diff16", "same4", "diff4", "none"):
            nt = 0 if kind == "none" else int(kind[4:])
            main = []
            for it in range(6):
                qs = [synth.query_host(spec, it * 4096 + 17 * nt + j, rows, n_terms=max(nt, 1), frequent_terms=nt // 2) for j in range(B)]
                if kind == "none":
                    terms = [synth.QueryTerms(0, np.zeros(0, np.uint64), None) for _ in qs]
                elif kind.startswith("same"):
                    terms = [qs[0].terms for _ in qs]
                else:
                    terms = [q.terms for q in qs]
                for rep in range(2):
                    sh.search_batch(np.stack([q.q for q in qs]), terms, spec.now_ticks, 50)
                    t = sh.last_timing()
                    if it >= 2 and rep == 1:
                        main.append(t["scan_ms"])
            print(f"c5same {kind}: main {statistics.mean(main):.3f} ms")
