"""Shared test helpers: synthetic corpora on the host, the oracle front-end, and the
tie-aware ranking comparator.

Tolerances (stated once, used everywhere):
  SCORE_RTOL = 1e-12   CUDA path vs oracle fp64 scores.  The contract in BASELINE.json is
                       1e-5 relative; the CUDA path re-scores survivors in fp64 with the
                       reference's arithmetic, so it is held to 1e-12 (the residual is the
                       order of fp64 additions and CUDA's exp vs glibc's, a few ulp).
  Row ids and their order must match exactly, except inside groups of hits whose oracle
  scores agree to within SCORE_RTOL (ties the contract leaves free).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCORE_RTOL = 1e-12
SCORE_ATOL = 1e-15


def oracle_search_synth(rows, query, now_ticks, top_k, candidate_cap=0, live=None, threads=1):
    """Runs the C oracle on host rows produced by omni_recall_rag_b200.synth (chunk Content is
    the space-joined token text, exactly what the reference would hold)."""
    from oracle import oracle_c
    from omni_recall_rag_b200 import synth

    blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
    return oracle_c.search(emb=rows.emb, dim=rows.emb.shape[1], ticks=rows.ticks, content_blob=blob,
                           content_off=off, query=query.text, qvec=query.q, now_ticks=now_ticks,
                           top_k=top_k, candidate_cap=candidate_cap, live=live, threads=threads)


def same_score(a: float, b: float) -> bool:
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return abs(a - b) <= SCORE_ATOL + SCORE_RTOL * max(abs(a), abs(b))


def assert_same_ranking(got_rows, got_scores, exp_rows, exp_scores, all_scores=None, what=""):
    """got == expected up to permutations inside near-tie groups.  `all_scores` (oracle score
    of every row, optional) lets a near-tie at the k-th boundary swap in a row from outside."""
    got_rows = [int(r) for r in got_rows]
    exp_rows = [int(r) for r in exp_rows]
    assert len(got_rows) == len(exp_rows), f"{what}: {len(got_rows)} hits, oracle has {len(exp_rows)}"
    for i, (g, e) in enumerate(zip(got_scores, exp_scores)):
        assert same_score(float(g), float(e)), f"{what}: score[{i}] {g!r} vs oracle {e!r}"
    if got_rows == exp_rows:
        return
    # group positions whose oracle scores are mutually within tolerance
    n = len(exp_rows)
    i = 0
    while i < n:
        j = i + 1
        while j < n and same_score(float(exp_scores[j]), float(exp_scores[i])):
            j += 1
        g, e = sorted(got_rows[i:j]), sorted(exp_rows[i:j])
        if g != e:
            assert j == n and all_scores is not None, f"{what}: rows differ at ranks {i}..{j - 1}: {got_rows[i:j]} vs {exp_rows[i:j]}"
            for r in got_rows[i:j]:
                assert same_score(float(all_scores[r]), float(exp_scores[i])), \
                    f"{what}: row {r} at the k-th boundary is not a near-tie"
        i = j


def random_unit_rows(rng, n, dim):
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return x.astype(np.float32)
