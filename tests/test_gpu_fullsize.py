"""Parity at the BASELINE.json sizes against the ORACLE (not against another CUDA path).  Needs a B200 (-m gpu).

The corpus is never materialised on the host: oracle_c.search_streamed regenerates it block by block on the
box's host cores (oracle/orr_oracle_stream.c, built from the same header-only generator as the device fill, which
tests/test_gpu_parity.py::test_device_fill_equals_host_generator pins bit for bit), scores every block with the
oracle and keeps a running top-k under the reference tie chain (RecallSearchService.cs:34-37).  The whole top-k of
the CUDA path must equal it: ids and order exact, scores within SCORE_RTOL = 1e-12 relative (contract 1e-5); only
groups of hits whose oracle scores agree within that tolerance may permute (tests/util.py).

Pinning: the oracle itself is reference-pinned on the six fixtures the reference's tests hold (top-1 identity;
tests/golden/reference_fixtures.json) — the C# scorer cannot run in this image.
"""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import synth
from oracle import oracle_c
from tests.util import assert_same_ranking

pytestmark = pytest.mark.gpu

NOW = synth.NOW_TICKS


def _streamed(spec, n, qs, k, with_emb=True):
    return oracle_c.search_streamed(spec, n, [q.text for q in qs], np.stack([q.q for q in qs]) if with_emb else None,
                                    NOW, k, block_rows=100_000, with_emb=with_emb)


def test_c2_1m_x_3072_top10_full_ranking_vs_streamed_oracle():
    """BASELINE.json configs[1]: 1M x 3072, single queries, top-10, fused scan + exact re-score.  Queries 0..5 plus the
    first planted ones (a corpus row + noise: a clear top hit), and the keyword+recency-only mode (no query embedding,
    the reference's default NoOp provider) on the same corpus through the exact path."""
    n, dim, k = 1_000_000, 3072, 10
    spec = synth.make_spec(dim)
    qis = list(range(6)) + [qi for qi in range(6, 60) if oracle_c.synth_query_source(spec, qi, n) is not None][:2]
    qs = [synth.query_host(spec, qi, n, n_terms=4) for qi in qis]
    exp = _streamed(spec, n, qs, k)
    exp_noemb = _streamed(spec, n, qs[:3], k, with_emb=False)
    with orr.RecallShard(dim, n) as sh:
        sh.fill_synthetic(spec, 0, n)
        for q, (er, es, et) in zip(qs, exp):
            got = sh.search(q.q, q.terms, NOW, k)
            assert sh.last_timing()["path"] == N.PATH_FUSED
            assert_same_ranking(got.rows, got.scores, er, es, what=f"c2 q={q.text}")
            assert got.ticks.tolist() == et.tolist()
        for q, (er, es, et) in zip(qs[:3], exp_noemb):
            got = sh.search(None, q.terms, NOW, k)
            assert sh.last_timing()["path"] == N.PATH_EXACT
            assert_same_ranking(got.rows, got.scores, er, es, what=f"c2 no-embedding q={q.text}")
            assert got.ticks.tolist() == et.tolist()


def _batch_vs_streamed(n, dim, batch, k, n_terms, freq, dup_ppm, checked):
    spec = synth.make_spec(dim, dup_row_ppm=dup_ppm)
    qs = [synth.query_host(spec, qi, n, n_terms=n_terms, frequent_terms=freq) for qi in range(batch)]
    pick = sorted(set(np.linspace(0, batch - 1, checked).astype(int).tolist()))
    exp = _streamed(spec, n, [qs[b] for b in pick], k)
    with orr.RecallShard(dim, n) as sh:
        sh.fill_synthetic(spec, 0, n)
        got = sh.search_batch(np.stack([q.q for q in qs]), [q.terms for q in qs], NOW, k)
        assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
        for b, (er, es, et) in zip(pick, exp):
            assert_same_ranking(got[b].rows, got[b].scores, er, es, what=f"batch b={b}")
            assert got[b].ticks.tolist() == et.tolist()


def test_c3_5m_x_768_batch_1024_top100_vs_streamed_oracle():
    """BASELINE.json configs[2]: the tcgen05 batched path; 8 queries of the 1024-batch checked on their whole top-100."""
    _batch_vs_streamed(5_000_000, 768, 1024, 100, 4, 0, 0, checked=8)


def test_c5_keyword_heavy_batch_256_top50_vs_streamed_oracle():
    """BASELINE.json configs[4]: 16-term queries (8 frequent), planted duplicate rows with equal and different
    timestamps (the ThenByDescending / stable-order fallbacks), batch 256, top-50; 8 queries checked in full."""
    _batch_vs_streamed(5_000_000, 768, 256, 50, 16, 8, 1000, checked=8)
