#!/usr/bin/env python
"""Text mode at scale: 1M x 3072 synthetic chunks with their text in HBM; timing of orr_search_text against the
fused hashed path, and equality of the two for whole-token terms.  python tools/text_check.py [rows] [dim]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
spec = synth.make_spec(dim)
NOW = spec.now_ticks
sh = orr.RecallShard(dim, rows)
sh.set_option("synth_text", 1)
sh.set_option("text_bytes_per_row", 600)
sh.fill_synthetic(spec, 0, rows)
for qi in range(6):
    q = synth.query_host(spec, qi, rows, n_terms=4)
    a = sh.search(q.q, q.terms, NOW, 10); ta = sh.last_timing()
    b = sh.search_text(q.q, q.text.split(), NOW, 10); tb = sh.last_timing()
    assert a.rows.tolist() == b.rows.tolist() and a.scores.tolist() == b.scores.tolist(), qi
    print(f"q{qi}: fused {ta['total_device_ms']:.3f} ms | text: match {tb['scan_ms']:.3f} ms + exact score/select {tb['finalize_ms']:.3f} ms")
for terms in (["t00000"], ["1", "23", "t0004"], ["t000001" + str(d) for d in range(10)]):
    q = synth.query_host(spec, 9, rows, n_terms=0)
    b = sh.search_text(q.q, terms, NOW, 10); tb = sh.last_timing()
    print(f"{len(terms)} prefix terms: match {tb['scan_ms']:.3f} ms + exact {tb['finalize_ms']:.3f} ms; top score {b.scores[0]:.4f}")
b = sh.search_text(q.q, ["t00000"], NOW, 10, candidate_cap=300); tb = sh.last_timing()
print(f"candidate_cap=300: match {tb['scan_ms']:.3f} ms + rescore {tb['finalize_ms']:.3f} ms")
print("text check ok")
