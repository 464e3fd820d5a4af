"""ctypes front-end of oracle/liborr_oracle.so (the C restatement) — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs; never from omni_recall_rag_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liborr_oracle.so")
_lib = None


class OracleHit(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_double), ("created_ticks", C.c_int64)]


class SynthSpec(C.Structure):
    """orr_synth_spec (include/orr.h), declared here too so the reference arm never touches liborr."""
    _fields_ = [
        ("seed", C.c_uint64), ("dim", C.c_int32), ("gen_dim", C.c_int32), ("terms_per_chunk", C.c_int32),
        ("vocab", C.c_int32), ("now_ticks", C.c_int64), ("zero_row_ppm", C.c_int32), ("dup_row_ppm", C.c_int32),
    ]


_SOURCES = ["orr_oracle.c", "orr_oracle_stream.c", "../omni_recall_rag_b200/csrc/orr_synth.h", "../include/orr.h", "Makefile"]


def build(force: bool = False) -> str:
    newest = max(os.path.getmtime(os.path.join(_DIR, s)) for s in _SOURCES)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", _DIR, "-B", "liborr_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_cosine.restype = C.c_double
        L.oracle_cosine.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.oracle_recency.restype = C.c_double
        L.oracle_recency.argtypes = [C.c_int64, C.c_int64]
        L.oracle_fuse.restype = C.c_double
        L.oracle_fuse.argtypes = [C.c_double, C.c_double, C.c_double]
        L.oracle_keyword.restype = C.c_double
        L.oracle_keyword.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
        L.oracle_query_terms.restype = C.c_int32
        L.oracle_query_terms.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
        L.oracle_round4.restype = C.c_double
        L.oracle_round4.argtypes = [C.c_double]
        L.oracle_snippet.restype = C.c_int64
        L.oracle_snippet.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_char_p, C.c_int64]
        L.oracle_max_threads.restype = C.c_int32
        L.oracle_search.restype = C.c_int32
        L.oracle_search.argtypes = [
            C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
            C.c_int32, C.c_int32, C.c_void_p]
        L.oracle_score_rows.restype = C.c_int32
        L.oracle_score_rows.argtypes = [
            C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_char_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_synth_spec_default.restype = None
        L.oracle_synth_spec_default.argtypes = [C.c_void_p, C.c_int32]
        L.oracle_synth_rows.restype = C.c_int32
        L.oracle_synth_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.oracle_synth_query.restype = C.c_int32
        L.oracle_synth_query.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.oracle_synth_query_source.restype = C.c_int32
        L.oracle_synth_query_source.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_synth_contents.restype = C.c_int64
        L.oracle_synth_contents.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.oracle_search_streamed.restype = C.c_int32
        L.oracle_search_streamed.argtypes = [
            C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int32,
            C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_contents(contents: Sequence[str]):
    """-> (uint8 blob, int64 offsets[n+1]) of UTF-8 chunk texts."""
    enc = [c.encode("utf-8") for c in contents]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc])
    blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
    return blob, off


def cosine(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if b is None:
        return float(lib().oracle_cosine(_ptr(a), a.size, None, 0))
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().oracle_cosine(_ptr(a), a.size, _ptr(b), b.size))


def recency(now_ticks: int, ticks: int) -> float:
    return float(lib().oracle_recency(now_ticks, ticks))


def fuse(c: float, k: float, r: float) -> float:
    return float(lib().oracle_fuse(c, k, r))


def keyword(query: str, content: str) -> float:
    q = query.encode("utf-8")
    c = content.encode("utf-8")
    return float(lib().oracle_keyword(q, len(q), c, len(c)))


def query_terms(query: str) -> list[str]:
    q = query.encode("utf-8")
    buf = C.create_string_buffer(4 * len(q) + 64)
    n = lib().oracle_query_terms(q, len(q), buf, len(buf))
    assert n >= 0
    parts = buf.raw.split(b"\0")[:n]
    return [p.decode("utf-8") for p in parts]


def round4(x: float) -> float:
    return float(lib().oracle_round4(x))


def snippet(content: str, max_chars: int = 180) -> str:
    c = content.encode("utf-8")
    buf = C.create_string_buffer(len(c) + 8)
    n = lib().oracle_snippet(c, len(c), max_chars, buf, len(buf))
    assert n >= 0
    return buf.raw[:n].decode("utf-8")


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def search(*, emb: Optional[np.ndarray], dim: int, ticks: np.ndarray,
           content_blob: Optional[np.ndarray], content_off: Optional[np.ndarray],
           query: str, qvec: np.ndarray, now_ticks: int, top_k: int, candidate_cap: int = 0,
           emb_off: Optional[np.ndarray] = None, live: Optional[np.ndarray] = None,
           threads: int = 1):
    """Runs oracle_search; returns (rows uint64[k], scores float64[k], ticks int64[k])."""
    n = int(ticks.shape[0])
    ticks = np.ascontiguousarray(ticks, dtype=np.int64)
    if emb is not None:
        emb = np.ascontiguousarray(emb, dtype=np.float32)
    if emb_off is not None:
        emb_off = np.ascontiguousarray(emb_off, dtype=np.int64)
    if live is not None:
        live = np.ascontiguousarray(live, dtype=np.uint8)
    qvec = np.ascontiguousarray(qvec, dtype=np.float32)
    k = max(1, int(top_k))
    out = (OracleHit * k)()
    q = query.encode("utf-8")
    got = lib().oracle_search(n, dim, _ptr(emb), _ptr(emb_off), _ptr(content_blob), _ptr(content_off),
                              _ptr(ticks), _ptr(live), q, len(q), _ptr(qvec), qvec.size,
                              now_ticks, top_k, candidate_cap, threads, C.cast(out, C.c_void_p))
    assert got >= 0
    rows = np.array([out[i].row for i in range(got)], dtype=np.uint64)
    scores = np.array([out[i].score for i in range(got)], dtype=np.float64)
    tk = np.array([out[i].created_ticks for i in range(got)], dtype=np.int64)
    return rows, scores, tk


def score_rows(*, emb: Optional[np.ndarray], dim: int, ticks: np.ndarray,
               content_blob: Optional[np.ndarray], content_off: Optional[np.ndarray],
               query: str, qvec: np.ndarray, now_ticks: int, emb_off: Optional[np.ndarray] = None):
    n = int(ticks.shape[0])
    ticks = np.ascontiguousarray(ticks, dtype=np.int64)
    if emb is not None:
        emb = np.ascontiguousarray(emb, dtype=np.float32)
    qvec = np.ascontiguousarray(qvec, dtype=np.float32)
    sc, cs, kw, rc = (np.zeros(n, dtype=np.float64) for _ in range(4))
    q = query.encode("utf-8")
    lib().oracle_score_rows(n, dim, _ptr(emb), _ptr(emb_off), _ptr(content_blob), _ptr(content_off),
                            _ptr(ticks), q, len(q), _ptr(qvec), qvec.size, now_ticks,
                            _ptr(sc), _ptr(cs), _ptr(kw), _ptr(rc))
    return sc, cs, kw, rc


# ---- synthetic corpora without liborr (orr_oracle_stream.c) -----------------------------------------------
def synth_spec(dim: int, *, seed: int = 20261018, gen_dim: Optional[int] = None, terms_per_chunk: int = 64,
               zero_row_ppm: int = 10000, dup_row_ppm: int = 0, now_ticks: Optional[int] = None) -> SynthSpec:
    """Same defaults as omni_recall_rag_b200.synth.make_spec; accepts a liborr OrrSynthSpec via copy_spec()."""
    sp = SynthSpec()
    lib().oracle_synth_spec_default(C.byref(sp), dim)
    sp.seed = seed
    sp.gen_dim = gen_dim if gen_dim is not None else max(dim, 3072)
    sp.terms_per_chunk = terms_per_chunk
    sp.zero_row_ppm = zero_row_ppm
    sp.dup_row_ppm = dup_row_ppm
    if now_ticks is not None:
        sp.now_ticks = now_ticks
    return sp


def copy_spec(spec) -> SynthSpec:
    """Field-by-field copy of any struct with orr_synth_spec's fields (e.g. _native.OrrSynthSpec)."""
    sp = SynthSpec()
    for f, _ in SynthSpec._fields_:
        setattr(sp, f, getattr(spec, f))
    return sp


def term_text(term_id: int) -> str:
    return "t%07d" % int(term_id)


def synth_rows(spec, first_row: int, n: int, *, want_emb: bool = True, threads: int = 0):
    """-> (emb float32[n, dim] | None, ticks int64[n], term_ids uint32[n, tpc]) on `threads` host threads."""
    sp = copy_spec(spec)
    emb = np.zeros((n, sp.dim), dtype=np.float32) if want_emb else None
    ticks = np.zeros(n, dtype=np.int64)
    tids = np.zeros((n, max(sp.terms_per_chunk, 1)), dtype=np.uint32)
    rc = lib().oracle_synth_rows(C.byref(sp), first_row, n, _ptr(emb), _ptr(ticks), _ptr(tids) if sp.terms_per_chunk else None,
                                 threads or max_threads())
    assert rc == 0
    return emb, ticks, tids[:, : sp.terms_per_chunk]


def synth_contents(term_ids: np.ndarray):
    """(uint8 blob, int64 offsets[n+1]) of the rows' Content: tokens joined by single spaces."""
    term_ids = np.ascontiguousarray(term_ids, dtype=np.uint32)
    n, tpc = term_ids.shape
    blob = np.zeros(max(1, n * max(0, 9 * tpc - 1)), dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    lib().oracle_synth_contents(_ptr(term_ids), n, tpc, _ptr(blob), _ptr(off))
    return blob, off


def synth_query(spec, qi: int, corpus_rows: int, n_terms: int = 4, frequent_terms: int = 0):
    """-> (q float32[dim], term_ids uint32[n_terms], query text)."""
    sp = copy_spec(spec)
    q = np.zeros(sp.dim, dtype=np.float32)
    tids = np.zeros(max(n_terms, 1), dtype=np.uint32)
    rc = lib().oracle_synth_query(C.byref(sp), qi, corpus_rows, n_terms, frequent_terms, _ptr(q), _ptr(tids))
    assert rc == 0
    tids = tids[:n_terms]
    return q, tids, " ".join(term_text(t) for t in tids)


def synth_query_source(spec, qi: int, corpus_rows: int) -> Optional[int]:
    """The corpus row query qi was planted next to (a clear top hit), or None for an independent draw."""
    sp = copy_spec(spec)
    src = C.c_uint64(0)
    return int(src.value) if lib().oracle_synth_query_source(C.byref(sp), qi, corpus_rows, C.byref(src)) else None


def search_streamed(spec, n_rows: int, queries: Sequence[str], qvecs: Optional[np.ndarray], now_ticks: int, top_k: int, *,
                    first_row: int = 0, block_rows: int = 100_000, with_emb: bool = True, threads: int = 0):
    """The oracle over synthetic rows [first_row, first_row + n_rows), generated block by block on the host cores
    (never materialised as a whole), for several queries at once.  Returns a list of (rows, scores, ticks) per query
    with GLOBAL row ids, in reference order."""
    sp = copy_spec(spec)
    nq = len(queries)
    enc = [q.encode("utf-8") for q in queries]
    qoff = np.zeros(nq + 1, dtype=np.int64)
    qoff[1:] = np.cumsum([len(e) for e in enc])
    qblob = b"".join(enc) or b"\0"
    q_len = 0
    if qvecs is not None:
        qvecs = np.ascontiguousarray(qvecs, dtype=np.float32)
        assert qvecs.shape[0] == nq
        q_len = int(qvecs.shape[1])
    k = max(1, int(top_k))
    out = (OracleHit * (nq * k))()
    n_out = np.zeros(max(nq, 1), dtype=np.int32)
    rc = lib().oracle_search_streamed(C.byref(sp), first_row, n_rows, block_rows, nq, qblob, _ptr(qoff),
                                      _ptr(qvecs) if q_len else None, q_len, 1 if with_emb else 0, now_ticks, top_k,
                                      threads or max_threads(), C.cast(out, C.c_void_p), _ptr(n_out))
    assert rc == 0
    a = np.frombuffer(out, dtype=np.dtype([("row", "<u8"), ("score", "<f8"), ("ticks", "<i8")])).reshape(nq, k)
    return [(a[q, : n_out[q]]["row"].copy(), a[q, : n_out[q]]["score"].copy(), a[q, : n_out[q]]["ticks"].copy()) for q in range(nq)]
