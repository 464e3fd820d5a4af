"""RecallCluster — one host process driving several GPUs (orr_cluster_*, csrc/orr_cluster.cu): the form the
single-process .NET API uses.  One orr_store per device, exchange buffers attached to each other; a search is
`orr_search_device` + the fused peer-memory all-gather/merge kernel on every device, issued from the calling
thread, no NCCL and no torch.distributed."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .shard import BatchHits, BatchTerms, Hits, QueryTerms, _HIT_DTYPE, _hits_from

ROW_SHIFT = 40  # global row id = shard << 40 | local row


class RecallCluster:
    def __init__(self, dim: int, capacity_rows_per_device: int, devices: Sequence[int], *, term_slots: int = 64,
                 max_top_k: int = 128, w_cos: float = 0.7, w_kw: float = 0.2, w_rec: float = 0.1, recency_days: float = 30.0):
        L = N.lib()
        cfg = N.OrrConfig()
        L.orr_config_default(C.byref(cfg))
        cfg.dim, cfg.term_slots, cfg.capacity_rows = dim, term_slots, capacity_rows_per_device
        cfg.w_cos, cfg.w_kw, cfg.w_rec, cfg.recency_days = w_cos, w_kw, w_rec, recency_days
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        self.dim, self.devices = dim, list(devices)
        self._h = C.c_void_p()
        N.check(L.orr_cluster_create(C.byref(cfg), devs.ctypes.data_as(C.c_void_p), len(devs), max_top_k, C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            N.lib().orr_cluster_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def count(self) -> int:
        return int(N.lib().orr_cluster_count(self._h))

    def fill_synthetic(self, spec: "N.OrrSynthSpec", first_row: int, n_per_shard: int) -> None:
        N.check(N.lib().orr_cluster_fill_synthetic(self._h, C.byref(spec), first_row, n_per_shard))

    def upsert_document_chunks(self, doc_key: int, emb: Optional[np.ndarray], ticks: np.ndarray,
                               term_hashes: Optional[Sequence[np.ndarray]] = None,
                               texts_lower: Optional[Sequence[str]] = None) -> np.ndarray:
        ticks = np.ascontiguousarray(ticks, dtype=np.int64)
        n = int(ticks.shape[0])
        if emb is not None:
            emb = np.ascontiguousarray(emb, dtype=np.float32)
        flat = off = None
        if term_hashes is not None:
            off = np.zeros(n + 1, dtype=np.uint32)
            off[1:] = np.cumsum([len(t) for t in term_hashes])
            flat = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.uint64) for t in term_hashes])
                                        if n and off[-1] else np.zeros(1, dtype=np.uint64))
        blob = toff = None
        if texts_lower is not None:
            enc = [t.encode("utf-8") for t in texts_lower]
            toff = np.zeros(n + 1, dtype=np.uint64)
            toff[1:] = np.cumsum([len(b) for b in enc])
            blob = b"".join(enc) or b"\0"
        out_rows = np.zeros(max(n, 1), dtype=np.uint64)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_cluster_upsert_document_chunks(self._h, doc_key, n, p(emb), None, p(ticks), p(flat), p(off), blob,
                                                           p(toff), p(out_rows)))
        return out_rows[:n]

    def delete_document(self, doc_key: int) -> None:
        N.check(N.lib().orr_cluster_delete_document(self._h, doc_key))

    def search(self, q: Optional[np.ndarray], terms: QueryTerms, now_ticks: int, top_k: int) -> Hits:
        if q is None:
            q = np.zeros(0, dtype=np.float32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = max(1, int(top_k))
        out = (N.OrrHit * k)()
        n = C.c_int32(0)
        ph = np.ascontiguousarray(terms.probe_hash, dtype=np.uint64)
        pt = None if terms.probe_term is None else np.ascontiguousarray(terms.probe_term, dtype=np.int32)
        N.check(N.lib().orr_cluster_search(
            self._h, q.ctypes.data_as(C.c_void_p) if q.size else None, int(q.size), int(terms.n_terms),
            ph.ctypes.data_as(C.c_void_p) if ph.size else None, None if pt is None else pt.ctypes.data_as(C.c_void_p),
            int(ph.size), int(now_ticks), int(top_k), C.cast(out, C.c_void_p), C.byref(n)))
        return _hits_from(out, n.value)

    def search_many(self, q: np.ndarray, terms, now_ticks: int, top_k: int) -> BatchHits:
        """orr_cluster_search_many: a run of SINGLE queries (q is [n, dim]; `terms` a sequence of n QueryTerms or a packed
        BatchTerms), pipelined — the exchange of query i overlaps every device's scan of query i+1.  Same hits as n
        search() calls."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        n, qd = int(q.shape[0]), int(q.shape[1]) if q.ndim == 2 else 0
        k = max(1, int(top_k))
        raw = np.zeros((max(n, 1), k), dtype=_HIT_DTYPE)
        n_out = np.zeros(max(n, 1), dtype=np.int32)
        bt = None
        if terms is not None:
            bt = terms if isinstance(terms, BatchTerms) else BatchTerms.pack(terms)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_cluster_search_many(self._h, n, p(q) if q.size else None, qd,
                                                p(bt.n_terms) if bt else None, p(bt.probe_hash) if bt else None,
                                                p(bt.probe_term) if bt else None, p(bt.probe_offsets) if bt else None,
                                                int(now_ticks), int(top_k), p(raw), p(n_out)))
        return BatchHits(raw[:n], n_out[:n])

    def search_batch(self, q: np.ndarray, terms, now_ticks: int, top_k: int) -> BatchHits:
        """orr_cluster_search_batch: orr_search_batch on every shard (one host thread per GPU) + per-query k-way merge."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B, qd = int(q.shape[0]), int(q.shape[1]) if q.ndim == 2 else 0
        k = max(1, int(top_k))
        raw = np.zeros((max(B, 1), k), dtype=_HIT_DTYPE)
        n_out = np.zeros(max(B, 1), dtype=np.int32)
        bt = None
        if terms is not None:
            bt = terms if isinstance(terms, BatchTerms) else BatchTerms.pack(terms)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        N.check(N.lib().orr_cluster_search_batch(self._h, B, p(q) if q.size else None, qd,
                                                 p(bt.n_terms) if bt else None, p(bt.probe_hash) if bt else None,
                                                 p(bt.probe_term) if bt else None, p(bt.probe_offsets) if bt else None,
                                                 int(now_ticks), int(top_k), p(raw), p(n_out)))
        return BatchHits(raw[:B], n_out[:B])
