"""Synthetic corpora — Python face of the counter-based generator in
csrc/orr_synth.h.  Host rows (for the oracle) and the device fill are bit-identical."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .shard import QueryTerms, hash_term

NOW_TICKS = 639_963_072_000_000_000  # 2026-10-18T00:00:00Z


def make_spec(dim: int, *, seed: int = 20261018, gen_dim: int | None = None, terms_per_chunk: int = 64,
              zero_row_ppm: int = 10000, dup_row_ppm: int = 0, now_ticks: int = NOW_TICKS) -> N.OrrSynthSpec:
    spec = N.OrrSynthSpec()
    N.lib().orr_synth_spec_default(C.byref(spec), dim)
    spec.seed = seed
    spec.gen_dim = gen_dim if gen_dim is not None else max(dim, 3072)
    spec.terms_per_chunk = terms_per_chunk
    spec.zero_row_ppm = zero_row_ppm
    spec.dup_row_ppm = dup_row_ppm
    spec.now_ticks = now_ticks
    return spec


@dataclass
class HostRows:
    emb: np.ndarray        # float32 [n, dim]
    ticks: np.ndarray      # int64 [n]
    term_ids: np.ndarray   # uint32 [n, terms_per_chunk]
    doc_first_row: np.ndarray  # uint64 [n]


def rows_host(spec: N.OrrSynthSpec, first_row: int, n: int, *, want_emb: bool = True) -> HostRows:
    emb = np.zeros((n, spec.dim), dtype=np.float32) if want_emb else None
    ticks = np.zeros(n, dtype=np.int64)
    tids = np.zeros((n, max(spec.terms_per_chunk, 1)), dtype=np.uint32)
    docs = np.zeros(n, dtype=np.uint64)
    N.check(N.lib().orr_synth_rows_host(
        C.byref(spec), first_row, n, None if emb is None else emb.ctypes.data_as(C.c_void_p),
        ticks.ctypes.data_as(C.c_void_p), tids.ctypes.data_as(C.c_void_p) if spec.terms_per_chunk else None,
        docs.ctypes.data_as(C.c_void_p)))
    return HostRows(emb, ticks, tids[:, : spec.terms_per_chunk], docs)


def term_text(term_id: int) -> str:
    return "t%07d" % int(term_id)


def contents_of(term_ids: np.ndarray) -> list[str]:
    """Chunk Content as the reference's chunker would produce it: words joined by one space
    (SlidingWindowTextChunker.cs:29)."""
    return [" ".join(term_text(t) for t in row) for row in term_ids]


@dataclass
class HostQuery:
    q: np.ndarray            # float32 [dim]
    term_ids: np.ndarray     # uint32 [n_terms]
    text: str                # the query string the reference would receive
    terms: QueryTerms        # hashed form for the C ABI


def query_host(spec: N.OrrSynthSpec, qi: int, corpus_rows: int, n_terms: int = 4, frequent_terms: int = 0) -> HostQuery:
    q = np.zeros(spec.dim, dtype=np.float32)
    tids = np.zeros(max(n_terms, 1), dtype=np.uint32)
    N.check(N.lib().orr_synth_query_host(C.byref(spec), qi, corpus_rows, n_terms, frequent_terms,
                                         q.ctypes.data_as(C.c_void_p), tids.ctypes.data_as(C.c_void_p)))
    tids = tids[:n_terms]
    text = " ".join(term_text(t) for t in tids)
    hashes = np.array([hash_term(term_text(t)) for t in tids], dtype=np.uint64)
    return HostQuery(q, tids, text, QueryTerms(n_terms, hashes, None))
