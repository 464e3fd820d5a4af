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


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "orr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
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
