"""GpuRecallSearchService — host-side mirror of the reference's IRecallSearchService
(src/OmniRecall.Api/Services/RecallSearchService.cs:6-57) with the scoring loop and the
ordering moved behind the C ABI (orr_search).

What stays on the host, unchanged from the reference: argument validation (:22-23), the
query embedding call (:25), the document lookup (:39) and the citation DTO build (:41-54,
TextSnippetHelper.cs:5-11, Math.Round(score, 4)).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Protocol, Sequence

import numpy as np

from .shard import QueryTerms, hash_term  # noqa: F401
from .store import GpuIngestionStore, _distinct_lower_tokens, _WS  # noqa: F401
from . import _native as N

STOP_WORDS = frozenset(  # RecallSearchService.cs:13-18
    "a an and are as at be by for from how in is it of on or that the to was what when where "
    "which who why with".split())


@dataclass(frozen=True)
class EmbeddingResult:  # IEmbeddingClient.cs:12-16
    vector: Sequence[float]
    status: str = "Success"
    model: Optional[str] = None
    message: Optional[str] = None


class IEmbeddingClient(Protocol):  # IEmbeddingClient.cs:18-21
    def embed(self, text: str) -> EmbeddingResult: ...


class NoOpEmbeddingClient:  # NoOpEmbeddingClient.cs:5-8 — the reference's default provider
    def embed(self, text: str) -> EmbeddingResult:
        return EmbeddingResult([], "NotSupported", None, "Embeddings provider disabled.")


@dataclass(frozen=True)
class RecallCitationDto:  # Contracts/RecallDtos.cs:5-12
    document_id: str
    file_name: str
    chunk_id: str
    chunk_index: int
    snippet: str
    score: float
    created_at_utc: int


@dataclass(frozen=True)
class RecallSearchResponseDto:  # Contracts/RecallDtos.cs:14-16
    query: str
    citations: List[RecallCitationDto]


class UnsupportedQueryError(ValueError):
    """The query needs more term probes than the kernels take (ORR_MAX_QUERY_PROBES)."""


def build_snippet(content: str, max_length: int) -> str:  # TextSnippetHelper.cs:5-11
    normalized = content.replace("\n", " ").replace("\r", " ")
    s, e = 0, len(normalized)
    while s < e and ord(normalized[s]) in _WS:
        s += 1
    while e > s and ord(normalized[e - 1]) in _WS:
        e -= 1
    normalized = normalized[s:e]
    return normalized if len(normalized) <= max_length else normalized[:max_length] + "..."


def math_round4(x: float) -> float:
    """Math.Round(double, 4): scale by 1e4, round half to even, unscale (:51)."""
    if x != x or x in (float("inf"), float("-inf")):
        return x
    return float(np.rint(np.float64(x) * 10000.0) / 10000.0)


class GpuRecallSearchService:
    """IRecallSearchService.  `now_ticks` is injectable (the reference reads DateTime.UtcNow
    per chunk, :117); `candidate_cap` defaults to the reference's 300 (:26)."""

    def __init__(self, store: GpuIngestionStore, embedding_client: IEmbeddingClient, *,
                 candidate_cap: int = 300, clock=None, keyword_mode: str = "auto"):
        """keyword_mode: "hashed" = hashed term table + vocabulary expansion (the fused scan; raises
        UnsupportedQueryError when a term expands to more probes than the kernels take), "text" = substring
        search on the chunk text in HBM (orr_search_text; any term, slower), "auto" = hashed when the
        expansion fits, else text."""
        if keyword_mode not in ("auto", "hashed", "text"):
            raise ValueError("keyword_mode must be auto, hashed or text")
        self.store = store
        self.embedding_client = embedding_client
        self.candidate_cap = candidate_cap
        self.keyword_mode = keyword_mode
        self._clock = clock or _utc_now_ticks

    @staticmethod
    def filtered_terms(query: str) -> List[str]:
        """KeywordScore's query side (:95-108): lower-cased distinct tokens minus stop words (or all of
        them if that leaves nothing)."""
        raw = _distinct_lower_tokens(query)
        return [t for t in raw if t not in STOP_WORDS] or raw

    _MODES = {"auto": 0, "hashed": 1, "text": 2}

    def query_terms(self, query: str) -> QueryTerms:
        """The hashed keyword side the library derives from `query` (orr_expand_query): KeywordScore's query side
        (:95-108) + substring expansion over the live vocabulary (:111).  search() does not need it (orr_search_query
        does both in one call); it is here for inspection and tests."""
        qt, n_probes = self.store.shard.expand_query(query)
        if qt.n_terms > N.ORR_MAX_QUERY_TERMS or n_probes > N.ORR_MAX_QUERY_PROBES:
            raise UnsupportedQueryError(
                f"{qt.n_terms} terms / {n_probes} vocabulary probes exceed the kernel limits "
                f"({N.ORR_MAX_QUERY_TERMS} / {N.ORR_MAX_QUERY_PROBES})")
        return qt

    def search(self, query: str, top_k: int) -> RecallSearchResponseDto:  # SearchAsync :20-57
        if query is None or all(ord(ch) in _WS for ch in query):
            raise ValueError("Query is required.")  # ArgumentException (:22-23)
        query_embedding = self.embedding_client.embed(query)  # :25
        qvec = np.asarray(query_embedding.vector, dtype=np.float32)
        try:
            # :26-37 behind the C ABI: tokenise + stop words (:95-108), vocabulary expansion on the GPU (:110-111),
            # fused scan, exact re-score, ordering
            hits = self.store.shard.search_query(query, qvec, self._clock(), top_k, candidate_cap=self.candidate_cap,
                                                 keyword_mode=self._MODES[self.keyword_mode])
        except N.OrrError as e:
            if e.code == N.ORR_E_UNSUPPORTED:
                raise UnsupportedQueryError(str(e)) from e
            raise
        scored = [(self.store.chunk_of_row(int(r)), float(s)) for r, s in zip(hits.rows, hits.scores)]
        documents = self.store.get_documents_by_ids(list({c.document_id for c, _ in scored}))  # :39
        citations = []
        for chunk, score in scored:  # :41-54
            doc = documents.get(chunk.document_id)
            citations.append(RecallCitationDto(
                chunk.document_id, doc.file_name if doc else "unknown", chunk.id, chunk.chunk_index,
                build_snippet(chunk.content, 180), math_round4(score), chunk.created_at_utc))
        return RecallSearchResponseDto(query, citations)


def _utc_now_ticks() -> int:
    import time
    return 621_355_968_000_000_000 + int(time.time() * 10_000_000)
