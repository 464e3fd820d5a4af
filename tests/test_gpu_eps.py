"""The error bounds the selection proofs ASSUME, measured.  Needs a B200 (-m gpu).

The fused path returns a row set chosen by an fp32 score and proves it complete with
    exact[k-th] - eps > tau   (tau = the best fp32 score of any discarded row),
which is sound only if |screen score - exact score| <= eps for every row the screen ranks.  The bounds are
    ORR_SELECT_EPS = 2e-5 * sum|w|                      fp32 scan (orr_scan.cu)
    ORR_BATCH_EPS  = 2e-4 * sum|w|                      bf16x3 split-precision tcgen05 screen (orr_batch.cu)
    ORR_BATCH_EPS * sum|w| + 0.0079 * |w_cos|           single bf16 tcgen05 screen
(csrc/orr_internal.h, orr_api.cu batch_eps).  Here the screens' own scores are read back (orr_debug_scan_scores,
orr_debug_batch_scores) on rows built to stress them — wide dynamic range inside a row, heavy cancellation, the widest
supported rows (dim 8192), tiny and huge norms — and compared with the ORACLE's per-row fp64 scores.  Rows the scan
itself declares unrankable (FLT_MAX: forced into the exact re-score) are exempt by construction."""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import synth
from oracle import oracle_c

pytestmark = pytest.mark.gpu

DAY = 864_000_000_000
NOW = synth.NOW_TICKS
SELECT_EPS = 2.0e-5          # ORR_SELECT_EPS x (0.7 + 0.2 + 0.1)
BATCH_EPS3 = 2.0e-4          # ORR_BATCH_EPS
BATCH_EPS1 = 2.0e-4 + 0.0079 * 0.7
FLT_MAX = np.finfo(np.float32).max


def _adversarial_rows(rng, n, dim):
    """Row families, n/8 each: unit Gaussian; element magnitudes log-uniform over 12 decades; two huge opposite
    elements + noise (cancellation in the dot product); near-duplicates of the query direction (cos ~ 1 - 1e-6);
    tiny norm (1e-9); huge norm (1e9); sparse one-hot-ish; constant sign (dot = sum of same-sign terms)."""
    q = rng.standard_normal(dim).astype(np.float32)
    fam = n // 8
    x = rng.standard_normal((n, dim))
    x[fam:2 * fam] *= 10.0 ** rng.uniform(-6, 6, size=(fam, dim))
    big = 10.0 ** rng.uniform(2, 5, size=fam)
    x[2 * fam:3 * fam, 0] = big / max(abs(float(q[0])), 1e-3)
    x[2 * fam:3 * fam, 1] = -big / max(abs(float(q[1])), 1e-3) * np.sign(q[0] * q[1])
    x[3 * fam:4 * fam] = q[None, :] + 1e-3 * x[3 * fam:4 * fam]
    x[4 * fam:5 * fam] *= 1e-9
    x[5 * fam:6 * fam] *= 1e9
    x[6 * fam:7 * fam] *= (rng.random((fam, dim)) < 4.0 / dim)
    x[7 * fam:8 * fam] = np.abs(x[7 * fam:8 * fam]) * np.sign(q)[None, :]
    return q, x.astype(np.float32)


def _oracle_scores(emb, dim, ticks, contents, query, q):
    blob, off = oracle_c.pack_contents(contents)
    return oracle_c.score_rows(emb=emb, dim=dim, ticks=ticks, content_blob=blob, content_off=off, query=query, qvec=q,
                               now_ticks=NOW)


@pytest.mark.parametrize("dim", [3072, 768, 8192, 100])
def test_fp32_scan_score_is_within_select_eps_of_the_exact_score(dim):
    rng = np.random.default_rng(dim)
    n = 2_048
    q, emb = _adversarial_rows(rng, n, dim)
    ticks = (NOW - (rng.random(n) * 400 * DAY).astype(np.int64)).astype(np.int64)
    vocab = np.array([f"k{i:02d}" for i in range(40)])
    contents = [" ".join(rng.choice(vocab, size=12, replace=False)) for _ in range(n)]
    query = " ".join(vocab[:5])
    terms = orr.QueryTerms(5, orr.tokenize_query(query), None)
    with orr.RecallShard(dim, n) as sh:
        sh.upsert_document_chunks(1, emb, ticks, [orr.tokenize_content(c) for c in contents])
        scan = sh.debug_scan_scores(q, terms, NOW).astype(np.float64)
    exact, cosv, kw, rec = _oracle_scores(emb, dim, ticks, contents, query, q)
    ranked = scan < FLT_MAX
    assert ranked.sum() >= n * 7 // 8 - 8                     # only the tiny-norm family may be handed to K3 unranked
    err = np.abs(scan[ranked] - exact[ranked])
    assert np.all(np.isfinite(err))
    worst = float(err.max())
    print(f"dim={dim}: max |fp32 scan - exact| = {worst:.3e} over {int(ranked.sum())} rows (eps {SELECT_EPS:.1e})")
    assert worst <= 0.5 * SELECT_EPS, worst                   # the bound holds with a factor 2 to spare


@pytest.mark.parametrize("passes,eps", [(3, BATCH_EPS3), (1, BATCH_EPS1)])
@pytest.mark.parametrize("dim", [768, 3072])
def test_tcgen05_screen_score_is_within_batch_eps_of_the_exact_score(dim, passes, eps):
    """The GEMM screen's raw score (w_cos*cos + w_rec*rec; the keyword term is added exactly, a multiple of w_kw/|terms|
    in fp32) against the oracle's cos*0.7 + rec*0.1.  Rows whose squared norm leaves fp32 screen as cosine 0 by design
     and are excluded like the scan's FLT_MAX rows: only norms in [1e-15, 1e15] are built here."""
    rng = np.random.default_rng(dim + passes)
    n, B = 2_048, 16
    q0, emb = _adversarial_rows(rng, n, dim)
    fam = n // 8
    emb[4 * fam:5 * fam] *= 1e3                               # 1e-6 .. : keep every row's norm inside the screen's range
    emb[5 * fam:6 * fam] *= 1e-3
    Q = np.stack([q0] + [rng.standard_normal(dim).astype(np.float32) * 10.0 ** rng.uniform(-3, 3) for _ in range(B - 1)])
    ticks = (NOW - (rng.random(n) * 400 * DAY).astype(np.int64)).astype(np.int64)
    with orr.RecallShard(dim, n) as sh:
        sh.set_option("batch_passes", passes)
        sh.upsert_document_chunks(1, emb, ticks)
        dense = sh.debug_batch_scores(Q, NOW).astype(np.float64)[:, :n]
    worst = 0.0
    for b in range(B):
        exact, cosv, kw, rec = _oracle_scores(emb, dim, ticks, [""] * n, "x", Q[b])
        err = np.abs(dense[b] - (cosv * 0.7 + rec * 0.1))
        assert np.all(np.isfinite(err)), b
        worst = max(worst, float(err.max()))
    print(f"dim={dim} passes={passes}: max |screen - exact| = {worst:.3e} (eps {eps:.3e})")
    assert worst <= 0.75 * eps, worst
