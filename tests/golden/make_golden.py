#!/usr/bin/env python
"""Regenerates tests/golden/reference_fixtures.json from the reference's own test sources.

Run in the build container only (needs /root/reference; the GPU box has no reference tree):
    python tests/golden/make_golden.py

What is PINNED BY THE REFERENCE: the fixture rows (parsed out of the C# test files below),
the queries, and the assertion each test makes (identity of Citations[0], or emptiness).
What is DERIVED: the full order and the fp64 scores, computed here with the numpy
restatement (oracle/oracle_np.py) at age 0 (the fixtures stamp every chunk with `now`).
The C oracle and the CUDA path are then tested against this file.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_np as onp  # noqa: E402

REF = "/root/reference/tests/OmniRecall.Api.Tests"
NOW = 639_000_000_000_000_000  # arbitrary fixed tick value; fixtures use age 0


def parse_seed_chunks(path):
    """CosmosChunkRecord initialisers inside SeedAsync (RecallSearchServiceTests.cs:88-116)."""
    src = open(path, encoding="utf-8").read()
    chunks = []
    for m in re.finditer(r"new CosmosChunkRecord\s*\{(.*?)\}", src, re.S):
        body = m.group(1)
        doc = re.search(r'DocumentId\s*=\s*"([^"]*)"', body).group(1)
        idx = int(re.search(r"ChunkIndex\s*=\s*(\d+)", body).group(1))
        content = re.search(r'Content\s*=\s*"([^"]*)"', body).group(1)
        emb = [float(x.rstrip("f")) for x in re.search(r"Embedding\s*=\s*\[([^\]]*)\]", body).group(1).split(",")]
        chunks.append({"document_id": doc, "chunk_index": idx, "content": content, "embedding": emb})
    files = dict(re.findall(r'Id\s*=\s*"(doc-\d)",\s*FileName\s*=\s*"([^"]*)"', src))
    return chunks, files


def run_case(chunks, query, qvec, top_k):
    recs = [onp.Chunk(c["content"], c["embedding"], NOW, i) for i, c in enumerate(chunks)]
    hits = onp.search(recs, query, qvec, NOW, top_k, candidate_cap=300)
    return [{"row": r, "score": s, "score_hex": float(s).hex(), "rounded": onp.round4(s)} for r, s, _ in hits]


def main():
    svc = os.path.join(REF, "Services", "RecallSearchServiceTests.cs")
    chunks, files = parse_seed_chunks(svc)
    assert len(chunks) == 3, chunks
    cases = []

    def add(name, cite, chunks_, query, qvec, top_k, asserted_first_row, asserted_file=None):
        hits = run_case(chunks_, query, qvec, top_k)
        if asserted_first_row is not None:
            assert hits and hits[0]["row"] == asserted_first_row, (name, hits)
        cases.append({"name": name, "reference_test": cite, "chunks": chunks_, "query": query,
                      "query_embedding": qvec, "top_k": top_k, "now_ticks": NOW,
                      "asserted_first_row": asserted_first_row, "asserted_file_name": asserted_file,
                      "derived_hits": hits})

    add("with_embeddings_most_similar_first", "Services/RecallSearchServiceTests.cs:9-21",
        chunks, "azure", [1.0, 0.0], 3, 0, files.get("doc-1"))
    add("no_query_embedding_falls_back_to_keyword", "Services/RecallSearchServiceTests.cs:24-35",
        chunks, "kubernetes", [], 3, 1, files.get("doc-2"))
    add("stop_words_do_not_dilute", "Services/RecallSearchServiceTests.cs:38-49",
        chunks, "what is the kubernetes", [], 3, 1, files.get("doc-2"))

    ep = open(os.path.join(REF, "Endpoints", "RecallEndpointTests.cs"), encoding="utf-8").read()
    ep = ep[ep.index("SearchRecall_AfterUpload_ReturnsCitations"):]
    text = re.search(r'var text = "([^"]*)"', ep).group(1)
    q = re.search(r'new RecallSearchRequestDto\("([^"]*)",\s*(\d+)\)', ep)
    add("endpoint_after_upload", "Endpoints/RecallEndpointTests.cs:11-30",
        [{"document_id": "doc_upload", "chunk_index": 0, "content": text, "embedding": []}],
        q.group(1), [], int(q.group(2)), 0, "nebula-notes.md")

    ch = open(os.path.join(REF, "Endpoints", "ChatEndpointTests.cs"), encoding="utf-8").read()
    ch = ch[ch.index("PostChat_AfterUpload_ReturnsCitations"):]
    vec = [float(x.rstrip("f")) for x in re.search(r"DeterministicEmbeddingClient\(\[([^\]]*)\]\)", ch).group(1).split(",")]
    text = re.search(r'var text = "([^"]*)"', ch).group(1)
    prompt = re.search(r'new ChatRequestDto\("([^"]*)"\)', ch).group(1)
    # ChatRequestDto's TopK default is 5 (Contracts/ChatDtos.cs); one chunk -> one citation
    add("chat_after_upload_citation_passes_guard", "Endpoints/ChatEndpointTests.cs:60-100",
        [{"document_id": "doc_upload", "chunk_index": 0, "content": text, "embedding": vec}],
        prompt, vec, 5, 0, "decision-log.md")
    add("empty_store_no_citations", "Endpoints/ChatEndpointTests.cs:26-58", [], "anything", [], 5, None)

    out = {"generated_by": "tests/golden/make_golden.py",
           "pinned_by_reference": "fixture rows, queries, asserted_first_row / asserted_file_name",
           "derived": "derived_hits (order beyond #1 and fp64 scores; numpy restatement, age 0)",
           "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1)
    for c in cases:
        print(c["name"], [(h["row"], h["score"]) for h in c["derived_hits"]])


if __name__ == "__main__":
    main()
