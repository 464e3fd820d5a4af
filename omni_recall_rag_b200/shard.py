"""RecallShard — one GPU's slice of the chunk store, a thin typed wrapper over the C ABI.

Everything here goes through liborr.so (include/orr.h); there is no Python-side scoring.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native as N


@dataclass
class Hits:
    rows: np.ndarray      # uint64 global row ids, reference order
    scores: np.ndarray    # float64
    ticks: np.ndarray     # int64 CreatedAtUtc ticks

    def __len__(self) -> int:
        return int(self.rows.shape[0])


@dataclass
class QueryTerms:
    """A query's keyword side: |terms| and the (hash, term) probes that satisfy them."""
    n_terms: int
    probe_hash: np.ndarray            # uint64[n_probes]
    probe_term: Optional[np.ndarray]  # int32[n_probes] or None (identity)

    @staticmethod
    def none() -> "QueryTerms":
        return QueryTerms(0, np.zeros(0, dtype=np.uint64), None)


@dataclass
class BatchTerms:
    """The keyword side of a query batch, packed once as the CSR orr_search_batch takes."""
    n_terms: np.ndarray                 # int32[B]
    probe_hash: np.ndarray              # uint64[total probes] (>= 1 element)
    probe_term: Optional[np.ndarray]    # int32[total probes] or None (identity)
    probe_offsets: np.ndarray           # uint32[B + 1]

    @staticmethod
    def pack(terms: Sequence[QueryTerms]) -> "BatchTerms":
        B = len(terms)
        nt = np.array([t.n_terms for t in terms], dtype=np.int32)
        po = np.zeros(B + 1, dtype=np.uint32)
        po[1:] = np.cumsum([len(t.probe_hash) for t in terms])
        ph = np.ascontiguousarray(np.concatenate([np.asarray(t.probe_hash, dtype=np.uint64) for t in terms])
                                  if B and po[-1] else np.zeros(1, dtype=np.uint64))
        pt = None
        if any(t.probe_term is not None for t in terms):
            pt = np.ascontiguousarray(np.concatenate(
                [np.asarray(t.probe_term if t.probe_term is not None else np.arange(len(t.probe_hash)), dtype=np.int32)
                 for t in terms]) if po[-1] else np.zeros(1, dtype=np.int32))
        return BatchTerms(nt, ph, pt, po)


_HIT_DTYPE = np.dtype([("row", "<u8"), ("score", "<f8"), ("ticks", "<i8")])


class BatchHits(Sequence):
    """Result of orr_search_batch: hits[b] is query b's Hits; .raw is the [B, k] record array."""

    def __init__(self, raw: np.ndarray, n_out: np.ndarray):
        self.raw, self.n_out = raw, n_out

    def __len__(self) -> int:
        return int(self.n_out.shape[0])

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        s = self.raw[b, : int(self.n_out[b])]
        return Hits(s["row"].copy(), s["score"].copy(), s["ticks"].copy())

    def __eq__(self, other):
        return list(self) == list(other)


def hash_term(term_lower: str) -> int:
    b = term_lower.encode("utf-8")
    return int(N.lib().orr_hash_term(b, len(b)))


def tokenize_query(query: str) -> np.ndarray:
    """KeywordScore's query side (RecallSearchService.cs:95-108) -> uint64 term hashes."""
    b = query.encode("utf-8")
    cap = max(16, len(b))
    out = np.zeros(cap, dtype=np.uint64)
    n = C.c_int32(0)
    N.check(N.lib().orr_tokenize_query(b, len(b), out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return out[: n.value].copy()


def tokenize_content(content: str) -> np.ndarray:
    """Distinct lower-cased white-space tokens of a chunk's Content -> uint64 hashes."""
    b = content.encode("utf-8")
    cap = max(16, len(b))
    out = np.zeros(cap, dtype=np.uint64)
    n = C.c_int32(0)
    N.check(N.lib().orr_tokenize_content(b, len(b), out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return out[: n.value].copy()


def merge_hits(lists: Sequence[Hits], top_k: int) -> Hits:
    """Global top-k of per-shard hit lists with the reference tie chain (orr_merge_hits)."""
    stride = max([len(h) for h in lists] + [1])
    buf = (N.OrrHit * (stride * len(lists)))()
    lens = np.zeros(len(lists), dtype=np.int32)
    for l, h in enumerate(lists):
        lens[l] = len(h)
        for i in range(len(h)):
            e = buf[l * stride + i]
            e.row, e.score, e.created_ticks = int(h.rows[i]), float(h.scores[i]), int(h.ticks[i])
    k = max(1, int(top_k))
    out = (N.OrrHit * k)()
    n = C.c_int32(0)
    N.check(N.lib().orr_merge_hits(C.cast(buf, C.c_void_p), lens.ctypes.data_as(C.c_void_p), len(lists), stride,
                                   top_k, C.cast(out, C.c_void_p), C.byref(n)))
    return _hits_from(out, n.value)


def _hits_from(arr, n: int) -> Hits:
    a = np.frombuffer(arr, dtype=np.dtype([("row", "<u8"), ("score", "<f8"), ("ticks", "<i8")]), count=n)
    return Hits(a["row"].copy(), a["score"].copy(), a["ticks"].copy())


class RecallShard:
    """Owns one orr_store (one GPU).  Thread-safe for concurrent search()."""

    def __init__(self, dim: int, capacity_rows: int, *, device: int = 0, term_slots: int = 64,
                 row_base: int = 0, w_cos: float = 0.7, w_kw: float = 0.2, w_rec: float = 0.1,
                 recency_days: float = 30.0):
        L = N.lib()
        cfg = N.OrrConfig()
        L.orr_config_default(C.byref(cfg))
        cfg.device, cfg.dim, cfg.term_slots = device, dim, term_slots
        cfg.capacity_rows, cfg.row_base = capacity_rows, row_base
        cfg.w_cos, cfg.w_kw, cfg.w_rec, cfg.recency_days = w_cos, w_kw, w_rec, recency_days
        self.cfg = cfg
        self.dim, self.term_slots, self.row_base, self.device = dim, term_slots, row_base, device
        self._h = C.c_void_p()
        N.check(L.orr_store_create(C.byref(cfg), C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            N.lib().orr_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- mutation ---------------------------------------------------------------------------
    def upsert_document_chunks(self, doc_key: int, emb: Optional[np.ndarray], ticks: np.ndarray,
                               term_hashes: Optional[Sequence[np.ndarray]] = None,
                               has_emb: Optional[np.ndarray] = None,
                               texts_lower: Optional[Sequence[str]] = None) -> np.ndarray:
        """orr_store_upsert_document_chunks(_text).  `texts_lower`: each chunk's lower-cased Content, kept in
        HBM for text mode (search_text); give it for every chunk of the store or for none."""
        ticks = np.ascontiguousarray(ticks, dtype=np.int64)
        n = int(ticks.shape[0])
        if emb is not None:
            emb = np.ascontiguousarray(emb, dtype=np.float32)
            if emb.shape != (n, self.dim):
                raise ValueError(f"emb must be ({n}, {self.dim}), got {emb.shape}")
        if has_emb is not None:
            has_emb = np.ascontiguousarray(has_emb, dtype=np.uint8)
        flat = off = None
        if term_hashes is not None:
            off = np.zeros(n + 1, dtype=np.uint32)
            off[1:] = np.cumsum([len(t) for t in term_hashes])
            flat = (np.concatenate([np.asarray(t, dtype=np.uint64) for t in term_hashes])
                    if n and off[-1] else np.zeros(1, dtype=np.uint64))
            flat = np.ascontiguousarray(flat, dtype=np.uint64)
        out_rows = np.zeros(max(n, 1), dtype=np.uint64)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        if texts_lower is not None:
            if len(texts_lower) != n:
                raise ValueError(f"{len(texts_lower)} texts for {n} chunks")
            enc = [t.encode("utf-8") for t in texts_lower]
            toff = np.zeros(n + 1, dtype=np.uint64)
            toff[1:] = np.cumsum([len(b) for b in enc])
            blob = b"".join(enc) or b"\0"
            N.check(N.lib().orr_store_upsert_document_chunks_text(self._h, doc_key, n, p(emb), p(has_emb), p(ticks),
                                                                  p(flat), p(off), blob, p(toff), p(out_rows)))
        else:
            N.check(N.lib().orr_store_upsert_document_chunks(self._h, doc_key, n, p(emb), p(has_emb), p(ticks),
                                                             p(flat), p(off), p(out_rows)))
        return out_rows[:n]

    def upsert_document_texts(self, doc_key: int, emb: Optional[np.ndarray], ticks: np.ndarray, contents: Sequence[str],
                              has_emb: Optional[np.ndarray] = None) -> np.ndarray:
        """orr_store_upsert_document_texts: Content strings as the reference holds them; the library lower-cases,
        tokenises, hashes and maintains the live vocabulary (and keeps the text in HBM with option keep_text)."""
        ticks = np.ascontiguousarray(ticks, dtype=np.int64)
        n = int(ticks.shape[0])
        if len(contents) != n:
            raise ValueError(f"{len(contents)} contents for {n} chunks")
        if emb is not None:
            emb = np.ascontiguousarray(emb, dtype=np.float32)
            if emb.shape != (n, self.dim):
                raise ValueError(f"emb must be ({n}, {self.dim}), got {emb.shape}")
        if has_emb is not None:
            has_emb = np.ascontiguousarray(has_emb, dtype=np.uint8)
        enc = [(c or "").encode("utf-8") for c in contents]
        coff = np.zeros(n + 1, dtype=np.uint64)
        coff[1:] = np.cumsum([len(b) for b in enc])
        blob = b"".join(enc) or b"\0"
        out_rows = np.zeros(max(n, 1), dtype=np.uint64)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_store_upsert_document_texts(self._h, doc_key, n, p(emb), p(has_emb), p(ticks), blob, p(coff), p(out_rows)))
        return out_rows[:n]

    def upsert_documents_texts(self, doc_keys: Sequence[int], chunk_counts: Sequence[int], emb: Optional[np.ndarray],
                               ticks: np.ndarray, contents: Sequence[str], has_emb: Optional[np.ndarray] = None) -> np.ndarray:
        """orr_store_upsert_documents_texts: many documents in one call (bulk / warm load).  Document d owns the next
        chunk_counts[d] entries of the chunk arrays.  Returns the global row ids of all chunks."""
        ticks = np.ascontiguousarray(ticks, dtype=np.int64)
        total = int(ticks.shape[0])
        keys = np.ascontiguousarray(doc_keys, dtype=np.uint64)
        doff = np.zeros(len(keys) + 1, dtype=np.uint32)
        doff[1:] = np.cumsum(np.asarray(chunk_counts, dtype=np.int64))
        if int(doff[-1]) != total or len(contents) != total:
            raise ValueError(f"chunk_counts sum to {int(doff[-1])}, {total} ticks, {len(contents)} contents")
        if emb is not None:
            emb = np.ascontiguousarray(emb, dtype=np.float32)
            if emb.shape != (total, self.dim):
                raise ValueError(f"emb must be ({total}, {self.dim}), got {emb.shape}")
        if has_emb is not None:
            has_emb = np.ascontiguousarray(has_emb, dtype=np.uint8)
        enc = [(c or "").encode("utf-8") for c in contents]
        coff = np.zeros(total + 1, dtype=np.uint64)
        coff[1:] = np.cumsum([len(b) for b in enc])
        blob = b"".join(enc) or b"\0"
        out_rows = np.zeros(max(total, 1), dtype=np.uint64)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_store_upsert_documents_texts(self._h, len(keys), p(keys), p(doff), p(emb), p(has_emb), p(ticks), blob,
                                                         p(coff), p(out_rows)))
        return out_rows[:total]

    @property
    def vocab_size(self) -> int:
        return int(N.lib().orr_store_vocab_size(self._h))

    def delete_document(self, doc_key: int) -> None:
        N.check(N.lib().orr_store_delete_document(self._h, doc_key))

    def fill_synthetic(self, spec: "N.OrrSynthSpec", first_row: int, n: int) -> None:
        N.check(N.lib().orr_store_fill_synthetic(self._h, C.byref(spec), first_row, n))

    def compact(self) -> np.ndarray:
        """orr_store_compact: drops tombstoned rows; returns old_rows[new_local_row] (global ids)."""
        out = np.zeros(max(self.rows_used, 1), dtype=np.uint64)
        n = C.c_int64(0)
        N.check(N.lib().orr_store_compact(self._h, out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(n)))
        return out[: n.value].copy()

    def save(self, path: str) -> None:
        """orr_store_save: the shard's byte image (rows, tombstones, term tables, text, document table)."""
        N.check(N.lib().orr_store_save(self._h, path.encode()))

    def load(self, path: str) -> None:
        """orr_store_load into this (empty, same dim / term_slots) shard; row ids are preserved."""
        N.check(N.lib().orr_store_load(self._h, path.encode()))

    def set_option(self, name: str, value: float) -> None:
        """orr_store_set_option, e.g. ("batch_passes", 0 = auto | 1 | 3)."""
        N.check(N.lib().orr_store_set_option(self._h, name.encode(), float(value)))

    @property
    def count(self) -> int:
        return int(N.lib().orr_store_count(self._h))

    @property
    def rows_used(self) -> int:
        return int(N.lib().orr_store_rows_used(self._h))

    # -- search -----------------------------------------------------------------------------
    def search(self, q: Optional[np.ndarray], terms: QueryTerms, now_ticks: int, top_k: int,
               candidate_cap: int = 0) -> Hits:
        """orr_search: host buffers in, hits out (H2D/D2H inside the call)."""
        if q is None:
            q = np.zeros(0, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = max(1, int(top_k))
        out = (N.OrrHit * k)()
        n = C.c_int32(0)
        ph = np.ascontiguousarray(terms.probe_hash, dtype=np.uint64)
        pt = None if terms.probe_term is None else np.ascontiguousarray(terms.probe_term, dtype=np.int32)
        N.check(N.lib().orr_search(
            self._h, q.ctypes.data_as(C.c_void_p) if q.size else None, int(q.size), int(terms.n_terms),
            ph.ctypes.data_as(C.c_void_p) if ph.size else None,
            None if pt is None else pt.ctypes.data_as(C.c_void_p), int(ph.size),
            int(now_ticks), int(top_k), int(candidate_cap), C.cast(out, C.c_void_p), C.byref(n)))
        return _hits_from(out, n.value)

    def search_query(self, query: str, q: Optional[np.ndarray], now_ticks: int, top_k: int, candidate_cap: int = 0,
                     keyword_mode: int = 0) -> Hits:
        """orr_search_query: the query STRING in; tokenising, stop words, vocabulary expansion (GPU) and the search
        happen behind the C ABI.  keyword_mode 0 = auto, 1 = hashed probes only, 2 = text mode only."""
        if q is None:
            q = np.zeros(0, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = max(1, int(top_k))
        out = (N.OrrHit * k)()
        n = C.c_int32(0)
        b = query.encode("utf-8")
        N.check(N.lib().orr_search_query(self._h, b, len(b), q.ctypes.data_as(C.c_void_p) if q.size else None, int(q.size),
                                         int(now_ticks), int(top_k), int(candidate_cap), int(keyword_mode),
                                         C.cast(out, C.c_void_p), C.byref(n)))
        return _hits_from(out, n.value)

    def expand_query(self, query: str, cap: int = N.ORR_MAX_QUERY_PROBES):
        """orr_expand_query -> (QueryTerms, n_probes): the probes are truncated to `cap`; n_probes is the full count."""
        b = query.encode("utf-8")
        ph = np.zeros(max(cap, 1), dtype=np.uint64)
        pt = np.zeros(max(cap, 1), dtype=np.int32)
        nt, npb = C.c_int32(0), C.c_int32(0)
        N.check(N.lib().orr_expand_query(self._h, b, len(b), ph.ctypes.data_as(C.c_void_p), pt.ctypes.data_as(C.c_void_p), cap,
                                         C.byref(nt), C.byref(npb)))
        m = min(cap, npb.value)
        return QueryTerms(nt.value, ph[:m].copy(), pt[:m].copy()), npb.value

    def search_text(self, q: Optional[np.ndarray], terms_lower: Sequence[str], now_ticks: int, top_k: int,
                    candidate_cap: int = 0) -> Hits:
        """orr_search_text: the keyword predicate evaluated as an ordinal substring search of each (lower-cased,
        A-2 filtered) query term in the chunk text kept in HBM; exact fp64 scoring of every candidate row."""
        if q is None:
            q = np.zeros(0, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = max(1, int(top_k))
        out = (N.OrrHit * k)()
        n = C.c_int32(0)
        enc = [t.encode("utf-8") for t in terms_lower]
        toff = np.zeros(len(enc) + 1, dtype=np.uint32)
        toff[1:] = np.cumsum([len(b) for b in enc])
        blob = b"".join(enc) or b"\0"
        N.check(N.lib().orr_search_text(
            self._h, q.ctypes.data_as(C.c_void_p) if q.size else None, int(q.size), len(enc), blob,
            toff.ctypes.data_as(C.c_void_p), int(now_ticks), int(top_k), int(candidate_cap),
            C.cast(out, C.c_void_p), C.byref(n)))
        return _hits_from(out, n.value)

    def search_device(self, q_dev_ptr: int, terms: QueryTerms, now_ticks: int, top_k: int,
                      out_dev_ptr: int, status_dev_ptr: int, stream_ptr: int) -> None:
        """orr_search_device: every buffer already in HBM, enqueued on `stream_ptr`, no sync."""
        ph = np.ascontiguousarray(terms.probe_hash, dtype=np.uint64)
        pt = None if terms.probe_term is None else np.ascontiguousarray(terms.probe_term, dtype=np.int32)
        N.check(N.lib().orr_search_device(
            self._h, C.c_void_p(q_dev_ptr), self.dim, int(terms.n_terms),
            ph.ctypes.data_as(C.c_void_p) if ph.size else None,
            None if pt is None else pt.ctypes.data_as(C.c_void_p), int(ph.size),
            int(now_ticks), int(top_k), C.c_void_p(out_dev_ptr), C.c_void_p(status_dev_ptr),
            C.c_void_p(stream_ptr)))

    def search_batch(self, q: np.ndarray, terms, now_ticks: int, top_k: int) -> BatchHits:
        """orr_search_batch: q is [B, dim] in host memory; `terms` is None, a sequence of B QueryTerms,
        or a pre-packed BatchTerms.  Returns BatchHits (hits[b] -> Hits of query b)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B, qd = int(q.shape[0]), int(q.shape[1]) if q.ndim == 2 else 0
        k = max(1, int(top_k))
        raw = np.zeros((max(B, 1), k), dtype=_HIT_DTYPE)
        n_out = np.zeros(max(B, 1), dtype=np.int32)
        bt = None
        if terms is not None:
            bt = terms if isinstance(terms, BatchTerms) else BatchTerms.pack(terms)
            if bt.n_terms.shape[0] != B:
                raise ValueError(f"terms describe {bt.n_terms.shape[0]} queries, q has {B}")
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_search_batch(self._h, B, p(q) if q.size else None, qd,
                                         p(bt.n_terms) if bt else None, p(bt.probe_hash) if bt else None,
                                         p(bt.probe_term) if bt else None, p(bt.probe_offsets) if bt else None,
                                         int(now_ticks), int(top_k), p(raw), p(n_out)))
        return BatchHits(raw[:B], n_out[:B])

    def search_batch_device(self, q: np.ndarray, terms, now_ticks: int, top_k: int, out_dev_ptr: int, n_out_dev_ptr: int) -> bool:
        """orr_search_batch_device: the answers stay in HBM (out_dev_ptr: orr_hit[B][k], n_out_dev_ptr: int32[B] on this
        shard's device).  False = the batch has to go through search_batch instead (outside the single-launch tcgen05
        path, or a query the screen could not prove); the buffers then hold nothing usable."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B, qd = int(q.shape[0]), int(q.shape[1]) if q.ndim == 2 else 0
        bt = None
        if terms is not None:
            bt = terms if isinstance(terms, BatchTerms) else BatchTerms.pack(terms)
            if bt.n_terms.shape[0] != B:
                raise ValueError(f"terms describe {bt.n_terms.shape[0]} queries, q has {B}")
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = N.lib().orr_search_batch_device(self._h, B, p(q) if q.size else None, qd,
                                             p(bt.n_terms) if bt else None, p(bt.probe_hash) if bt else None,
                                             p(bt.probe_term) if bt else None, p(bt.probe_offsets) if bt else None,
                                             int(now_ticks), int(top_k), C.c_void_p(out_dev_ptr), C.c_void_p(n_out_dev_ptr))
        if rc == N.ORR_E_UNSUPPORTED:
            return False
        N.check(rc)
        return True

    def debug_scan_scores(self, q: np.ndarray, terms: QueryTerms, now_ticks: int) -> np.ndarray:
        """The fp32 score the fused scan computes for every physical row (orr_debug_scan_scores)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros(max(self.rows_used, 1), dtype=np.float32)
        ph = np.ascontiguousarray(terms.probe_hash, dtype=np.uint64)
        pt = None if terms.probe_term is None else np.ascontiguousarray(terms.probe_term, dtype=np.int32)
        N.check(N.lib().orr_debug_scan_scores(
            self._h, q.ctypes.data_as(C.c_void_p), int(q.size), int(terms.n_terms),
            ph.ctypes.data_as(C.c_void_p) if ph.size else None, None if pt is None else pt.ctypes.data_as(C.c_void_p),
            int(ph.size), int(now_ticks), out.ctypes.data_as(C.c_void_p), out.shape[0]))
        return out[: self.rows_used]

    def debug_batch_scores(self, q: np.ndarray, now_ticks: int, tile_stride: int = 1) -> np.ndarray:
        """Raw fused GEMM scores (w_cos*cos + w_rec*rec) of every stride-th 256-row tile: [B, n]."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B = int(q.shape[0])
        tiles = (self.rows_used + 255) // 256
        n_s = ((tiles + tile_stride - 1) // tile_stride) * 256
        out = np.zeros((B, n_s), dtype=np.float32)
        N.check(N.lib().orr_debug_batch_scores(self._h, B, q.ctypes.data_as(C.c_void_p), int(q.shape[1]), int(now_ticks),
                                               int(tile_stride), out.ctypes.data_as(C.c_void_p), n_s))
        return out

    def last_device_timing(self) -> dict:
        """CUDA-event durations of the last search_device call (waits for it)."""
        t = N.OrrTiming()
        N.check(N.lib().orr_search_device_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in N.OrrTiming._fields_}

    def last_timing(self) -> dict:
        t = N.OrrTiming()
        N.lib().orr_last_timing(C.byref(t))
        return {f: getattr(t, f) for f, _ in N.OrrTiming._fields_}
