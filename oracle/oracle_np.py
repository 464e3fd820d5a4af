"""numpy restatement of the reference's hybrid recall scorer — TEST INFRASTRUCTURE ONLY.

A second, independent restatement of SURVEY.md Appendix A (the first is oracle/orr_oracle.c)
so the two can be triangulated against each other, because the C# reference itself cannot be
run in this image.  Nothing under omni_recall_rag_b200/ may import this module.

Citations are to /root/reference/src/OmniRecall.Api/Services/RecallSearchService.cs unless
another file is named.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

TICKS_PER_DAY = 864_000_000_000

STOP_WORDS = frozenset(  # :13-18
    "a an and are as at be by for from how in is it of on or that the to was what when "
    "where which who why with".split()
)

# char.IsWhiteSpace — separators of string.Split((char[])null) (:95) and IsNullOrWhiteSpace (:92)
_WS = set(
    [0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000]
    + list(range(0x09, 0x0E))
    + list(range(0x2000, 0x200B))
)


def is_ws(ch: str) -> bool:
    return ord(ch) in _WS


def to_lower_invariant(s: str) -> str:
    """ToLowerInvariant (:96,:110): per-code-point SIMPLE lower-case mapping (no context rules, no
    expansions); U+0130 stays unchanged, as .NET's invariant casing leaves it."""
    out = []
    for ch in s:
        lo = ch.lower()
        out.append(lo if len(lo) == 1 else ch)
    return "".join(out)


def split_ws(s: str) -> list[str]:
    toks, cur = [], []
    for ch in s:
        if is_ws(ch):
            if cur:
                toks.append("".join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        toks.append("".join(cur))
    return toks


def is_null_or_whitespace(s: Optional[str]) -> bool:
    return s is None or all(is_ws(ch) for ch in s)


def query_terms(query: str) -> list[str]:
    """A-2 (:95-108): split, lower, ordinal-distinct, drop stop words unless that empties."""
    if is_null_or_whitespace(query):
        return []
    raw: list[str] = []
    for t in split_ws(query):
        t = to_lower_invariant(t)
        if t not in raw:
            raw.append(t)
    if not raw:
        return []
    kept = [t for t in raw if t not in STOP_WORDS]
    return kept if kept else raw


def keyword_score(query: str, content: Optional[str]) -> float:
    """KeywordScore (:90-113)."""
    if is_null_or_whitespace(query) or is_null_or_whitespace(content):
        return 0.0
    terms = query_terms(query)
    if not terms:
        return 0.0
    lowered = to_lower_invariant(content)
    matches = sum(1 for t in terms if t in lowered)  # ordinal substring (:111)
    return matches / len(terms)


def cosine(a: Sequence[float], b: Optional[Sequence[float]]) -> float:
    """CosineSimilarity (:69-88): fp32 products, fp64 sequential accumulation."""
    a32 = np.asarray(a, dtype=np.float32)
    if b is None:
        return 0.0
    b32 = np.asarray(b, dtype=np.float32)
    if a32.size == 0 or b32.size == 0 or a32.size != b32.size:  # :71-72
        return 0.0
    with np.errstate(all="ignore"):
        # np.cumsum is a strict left-to-right recurrence, i.e. the loop at :77-82
        dot = np.cumsum((a32 * b32).astype(np.float64))[-1]
        na = np.cumsum((a32 * a32).astype(np.float64))[-1]
        nb = np.cumsum((b32 * b32).astype(np.float64))[-1]
        if na <= 0.0 or nb <= 0.0:  # :84-85 (false for NaN -> NaN propagates)
            return 0.0
        return float(dot / (np.sqrt(na) * np.sqrt(nb)))  # :87


def recency(now_ticks: int, created_ticks: int) -> float:
    """RecencyScore (:115-119) with the clock injected."""
    age_days = float(now_ticks - created_ticks) / 864000000000.0
    age_days = max(0.0, age_days)
    return math.exp(-age_days / 30.0)


def fuse(cos_v: float, kw: float, rec: float) -> float:
    """ScoreChunk (:66), evaluated left to right in fp64."""
    return (cos_v * 0.7 + kw * 0.2) + rec * 0.1


def round4(x: float) -> float:
    """Math.Round(score, 4) (:51): banker's rounding of the scaled double."""
    if math.isnan(x) or math.isinf(x):
        return x
    return float(np.rint(x * 10000.0) / 10000.0)


def build_snippet(content: str, max_length: int = 180) -> str:
    """TextSnippetHelper.BuildSnippet (TextSnippetHelper.cs:5-11)."""
    normalized = content.replace("\n", " ").replace("\r", " ")
    # .NET Trim() strips char.IsWhiteSpace from both ends
    s, e = 0, len(normalized)
    while s < e and is_ws(normalized[s]):
        s += 1
    while e > s and is_ws(normalized[e - 1]):
        e -= 1
    normalized = normalized[s:e]
    if len(normalized) <= max_length:
        return normalized
    return normalized[:max_length] + "..."


@dataclass
class Chunk:
    """CosmosChunkRecord (Data/Models/CosmosIngestionRecords.cs:19-30), scoring fields only."""

    content: str
    embedding: Optional[Sequence[float]]
    ticks: int
    row: int = -1


def _sort_key_desc(score: float):
    # Comparer<double>: NaN below everything -> last under OrderByDescending (:34)
    return (1, 0.0) if math.isnan(score) else (0, -score)


def search(chunks: Sequence[Chunk], query: str, qvec: Sequence[float], now_ticks: int,
           top_k: int, candidate_cap: int = 300):
    """SearchAsync's scoring and ordering (:26-37) + GetRecentChunksAsync
    (InMemoryIngestionStore.cs:57-65).  Returns [(row, score, ticks)].
    candidate_cap=0 scores every chunk (the north-star extension)."""
    cands = list(enumerate(chunks))
    # stable OrderByDescending(CreatedAtUtc) (InMemoryIngestionStore.cs:61)
    cands.sort(key=lambda ic: -ic[1].ticks)
    if candidate_cap != 0:
        cands = cands[: max(1, candidate_cap)]
    scored = []
    for i, c in cands:
        s = fuse(cosine(qvec, c.embedding), keyword_score(query, c.content), recency(now_ticks, c.ticks))
        scored.append((i, s, c.ticks))
    # stable OrderByDescending(score).ThenByDescending(ticks) (:34-35)
    scored.sort(key=lambda t: (_sort_key_desc(t[1]), -t[2]))
    return scored[: max(1, top_k)]
