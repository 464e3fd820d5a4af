"""GpuIngestionStore — host-side mirror of the reference's IIngestionStore
(src/OmniRecall.Api/Services/IIngestionStore.cs:5-17) whose chunk rows live in HBM.

Python stands in for the C# host here (no .NET toolchain in this image); the C# class of
the same name in dotnet/GpuIngestionStore.cs makes the same calls through P/Invoke.  Method
names are the reference's, snake-cased; behaviour follows InMemoryIngestionStore.cs line by
line, including its quirks (whole batch filed under chunks[0].DocumentId, :22-23).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .shard import RecallShard, hash_term, tokenize_content


@dataclass(frozen=True)
class CosmosDocumentRecord:  # Data/Models/CosmosIngestionRecords.cs:5-17
    id: str = ""
    file_name: str = ""
    source_type: str = "file"
    blob_path: str = ""
    content_hash: str = ""
    chunk_count: int = 0
    created_at_utc: int = 0  # .NET ticks


@dataclass(frozen=True)
class CosmosChunkRecord:  # Data/Models/CosmosIngestionRecords.cs:19-30
    id: str = ""
    document_id: str = ""
    chunk_index: int = 0
    content: str = ""
    embedding: Optional[Sequence[float]] = None
    created_at_utc: int = 0  # .NET ticks


class GpuIngestionStore:
    """IIngestionStore over one RecallShard.  Document records, chunk text and ids stay on the host (they are only
    needed for the <= top_k citations); embeddings, timestamps, hashed term sets, the live vocabulary and (keep_text)
    the lower-cased Content are mirrored into HBM on every upsert, behind ONE C-ABI call per document
    (orr_store_upsert_document_texts: the library tokenises and hashes)."""

    def __init__(self, dim: int, capacity_rows: int = 1 << 16, *, device: int = 0, term_slots: int = 128,
                 keep_text: bool = True, text_bytes_per_row: int = 2048):
        self.shard = RecallShard(dim, capacity_rows, device=device, term_slots=term_slots)
        self.dim = dim
        self.keep_text = keep_text          # lower-cased Content mirrored into HBM (text mode, orr_search_text)
        if keep_text:
            self.shard.set_option("text_bytes_per_row", text_bytes_per_row)
            self.shard.set_option("keep_text", 1)
        self._lock = threading.RLock()
        self._documents: Dict[str, CosmosDocumentRecord] = {}
        self._chunks_by_document: Dict[str, List[CosmosChunkRecord]] = {}
        self._rows_by_document: Dict[str, np.ndarray] = {}
        self._chunk_by_row: Dict[int, CosmosChunkRecord] = {}
        self._synth = None                  # (spec, first_row, n): rows filled by fill_synthetic have no host records

    def close(self) -> None:
        self.shard.close()

    # -- IIngestionStore --------------------------------------------------------------------
    def upsert_document(self, document: CosmosDocumentRecord) -> CosmosDocumentRecord:  # :11-15
        with self._lock:
            self._documents[document.id] = document
        return document

    def upsert_chunks(self, chunks: Sequence[CosmosChunkRecord]) -> None:  # :17-25
        if len(chunks) == 0:
            return
        document_id = chunks[0].document_id
        ordered = sorted(chunks, key=lambda c: c.chunk_index)  # OrderBy is stable (:23)
        n = len(ordered)
        emb = np.zeros((n, self.dim), dtype=np.float32)
        has = np.zeros(n, dtype=np.uint8)
        ticks = np.zeros(n, dtype=np.int64)
        for i, c in enumerate(ordered):
            e = c.embedding
            # an Embedding that is null/empty/of another width scores cosine 0 against any
            # query of the store's width (RecallSearchService.cs:71-72)
            if e is not None and len(e) == self.dim:
                emb[i] = np.asarray(e, dtype=np.float32)
                has[i] = 1
            ticks[i] = c.created_at_utc
        with self._lock:
            rows = self.shard.upsert_document_texts(_doc_key(document_id), emb, ticks, [c.content or "" for c in ordered], has)
            self._forget_rows(document_id)
            self._chunks_by_document[document_id] = ordered
            self._rows_by_document[document_id] = rows
            for r, c in zip(rows, ordered):
                self._chunk_by_row[int(r)] = c

    def upsert_chunks_bulk(self, batches: Sequence[Sequence[CosmosChunkRecord]]) -> None:
        """Many UpsertChunksAsync calls as ONE native call (orr_store_upsert_documents_texts): warm load / hydration of the
        store from the full container.  Every batch keeps the single-call semantics (:17-25: filed under its first chunk's
        DocumentId, ordered by ChunkIndex); a document may appear once per call."""
        batches = [sorted(b, key=lambda c: c.chunk_index) for b in batches if len(b)]
        if not batches:
            return
        doc_ids = [b[0].document_id for b in batches]
        total = sum(len(b) for b in batches)
        emb = np.zeros((total, self.dim), dtype=np.float32)
        has = np.zeros(total, dtype=np.uint8)
        ticks = np.zeros(total, dtype=np.int64)
        contents = []
        i = 0
        for b in batches:
            for c in b:
                e = c.embedding
                if e is not None and len(e) == self.dim:
                    emb[i] = np.asarray(e, dtype=np.float32)
                    has[i] = 1
                ticks[i] = c.created_at_utc
                contents.append(c.content or "")
                i += 1
        with self._lock:
            rows = self.shard.upsert_documents_texts([_doc_key(d) for d in doc_ids], [len(b) for b in batches], emb, ticks, contents, has)
            at = 0
            for d, b in zip(doc_ids, batches):
                self._forget_rows(d)
                self._chunks_by_document[d] = list(b)
                self._rows_by_document[d] = rows[at:at + len(b)].copy()
                for r, c in zip(rows[at:at + len(b)], b):
                    self._chunk_by_row[int(r)] = c
                at += len(b)

    def get_document(self, document_id: str) -> Optional[CosmosDocumentRecord]:  # :27-31
        return self._documents.get(document_id)

    def list_documents(self, max_count: int) -> List[CosmosDocumentRecord]:  # :33-40
        with self._lock:
            docs = sorted(self._documents.values(), key=lambda d: -d.created_at_utc)
        return docs[: max(1, max_count)]

    def get_chunks_by_document_id(self, document_id: str) -> List[CosmosChunkRecord]:  # :42-48
        return list(self._chunks_by_document.get(document_id, []))

    def delete_document(self, document_id: str) -> None:  # :50-55
        with self._lock:
            self._documents.pop(document_id, None)
            if document_id in self._chunks_by_document:
                self._forget_rows(document_id)
                self.shard.delete_document(_doc_key(document_id))
                del self._chunks_by_document[document_id]

    def get_recent_chunks(self, max_count: int) -> List[CosmosChunkRecord]:  # :57-65
        with self._lock:
            flat = [c for chunks in self._chunks_by_document.values() for c in chunks]
        flat.sort(key=lambda c: -c.created_at_utc)
        return flat[: max(1, max_count)]

    def get_documents_by_ids(self, document_ids: Sequence[str]) -> Dict[str, CosmosDocumentRecord]:  # :67-76
        wanted = set(document_ids)
        with self._lock:
            return {k: v for k, v in self._documents.items() if k in wanted}

    # -- maintenance: not in IIngestionStore; the HBM layout needs them -------------------------
    def compact(self) -> int:
        """Drops the rows that replace/delete tombstoned and remaps the host's row tables.  Returns the
        number of rows reclaimed."""
        with self._lock:
            before = self.shard.rows_used
            old_rows = self.shard.compact()
            base = self.shard.row_base
            new_of_old = {int(o): base + i for i, o in enumerate(old_rows)}
            self._chunk_by_row = {new_of_old[r]: c for r, c in self._chunk_by_row.items()}
            self._rows_by_document = {d: np.array([new_of_old[int(r)] for r in rows], dtype=np.uint64)
                                      for d, rows in self._rows_by_document.items()}
            return before - len(old_rows)

    def save(self, directory: str) -> None:
        """HBM image + vocabulary (orr_store_save) and the host-side records as JSON, so a restart does not re-ingest."""
        import json
        import os

        os.makedirs(directory, exist_ok=True)
        with self._lock:
            self.shard.save(os.path.join(directory, "shard.orrsnap"))
            host = {
                "documents": [vars(d) for d in self._documents.values()],
                "chunks": {d: [dict(vars(c), embedding=None) for c in cs] for d, cs in self._chunks_by_document.items()},
                "rows": {d: [int(r) for r in rows] for d, rows in self._rows_by_document.items()},
            }
            with open(os.path.join(directory, "host.json"), "w", encoding="utf-8") as f:
                json.dump(host, f)

    def load(self, directory: str) -> None:
        """Plain data only (JSON): nothing in a snapshot directory is executed.  Embeddings live in the HBM image;
        the host records come back without them (they are never read on the host)."""
        import json
        import os

        with self._lock:
            self.shard.load(os.path.join(directory, "shard.orrsnap"))
            with open(os.path.join(directory, "host.json"), encoding="utf-8") as f:
                h = json.load(f)
            self._documents = {d["id"]: CosmosDocumentRecord(**d) for d in h["documents"]}
            self._chunks_by_document = {d: [CosmosChunkRecord(**c) for c in cs] for d, cs in h["chunks"].items()}
            self._rows_by_document = {d: np.array(rows, dtype=np.uint64) for d, rows in h["rows"].items()}
            self._chunk_by_row = {int(r): c for d, rows in self._rows_by_document.items()
                                  for r, c in zip(rows, self._chunks_by_document[d])}

    # -- synthetic corpora (bench / tests): rows generated on the device, records derived on demand ---------------
    def fill_synthetic(self, spec, n: int, first_row: int = 0) -> None:
        """The bench corpus behind the service API: rows come from orr_store_fill_synthetic (with the chunk text if the
        store keeps text), the 2^20 synthetic tokens are registered as the vocabulary, and a hit's CosmosChunkRecord is
        rebuilt from the generator when a citation needs it (document = the run of rows sharing one timestamp)."""
        with self._lock:
            if self.shard.rows_used != 0:
                raise ValueError("fill_synthetic needs an empty store")
            self.shard.set_option("synth_vocab", 1)
            if self.keep_text:
                self.shard.set_option("synth_text", 1)
            self.shard.fill_synthetic(spec, first_row, n)
            self._synth = (spec, first_row, n)

    def _synthetic_chunk(self, row: int) -> CosmosChunkRecord:
        import ctypes as C

        from . import _native as N

        spec, first_row, _ = self._synth
        L = N.lib()
        g = first_row + (row - self.shard.row_base)
        if getattr(self, "_synth_buf", None) is None:
            self._synth_buf = C.create_string_buffer(9 * max(spec.terms_per_chunk, 1))
            self._synth_spec_ref = C.byref(spec)
        ticks, first = C.c_int64(0), C.c_uint64(0)
        L.orr_synth_row_info(self._synth_spec_ref, g, C.byref(ticks), C.byref(first))
        ln = L.orr_synth_row_text(self._synth_spec_ref, g, self._synth_buf, len(self._synth_buf))
        doc_first = first.value
        return CosmosChunkRecord(f"synth-{doc_first}:{g - doc_first:04d}", f"synth-{doc_first}", g - doc_first,
                                 self._synth_buf.raw[:ln].decode("ascii"), None, ticks.value)

    # -- used by GpuRecallSearchService -------------------------------------------------------
    def chunk_of_row(self, row: int) -> CosmosChunkRecord:
        c = self._chunk_by_row.get(int(row))
        if c is None and self._synth is not None:
            return self._synthetic_chunk(int(row))
        if c is None:
            raise KeyError(row)
        return c

    @property
    def vocabulary_size(self) -> int:
        """Distinct tokens held by live chunks (kept in HBM; query terms are expanded over it on the GPU)."""
        return self.shard.vocab_size

    def _forget_rows(self, document_id: str) -> None:
        rows = self._rows_by_document.pop(document_id, None)
        if rows is None:
            return
        for r in rows:
            self._chunk_by_row.pop(int(r), None)


def _doc_key(document_id: str) -> int:
    return hash_term("doc:" + document_id)


_WS = {0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000, *range(0x09, 0x0E), *range(0x2000, 0x200B)}


def _lower_invariant(s: str) -> str:
    """ToLowerInvariant (RecallSearchService.cs:96,110): the simple lower-case mapping per code point; U+0130, the one
    code point whose full lower-casing is two code points, stays as it is (.NET invariant casing)."""
    if s.isascii():
        return s.lower()
    out = []
    for ch in s:
        lo = ch.lower()
        out.append(lo if len(lo) == 1 else ch)
    return "".join(out)


def _distinct_lower_tokens(content: str) -> List[str]:
    """Host copy of orr_tokenize_content's token list (needed as strings for the vocabulary)."""
    toks: List[str] = []
    seen = set()
    cur: List[str] = []
    for ch in content or "":
        if ord(ch) in _WS:
            if cur:
                t = _lower_invariant("".join(cur))
                if t not in seen:
                    seen.add(t)
                    toks.append(t)
                cur = []
        else:
            cur.append(ch)
    if cur:
        t = _lower_invariant("".join(cur))
        if t not in seen:
            toks.append(t)
    return toks
