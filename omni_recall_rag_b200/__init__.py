"""omni_recall_rag_b200 — B200-native hybrid recall scorer (drop-in for the scoring path of
fchchen/omni-recall-rag's POST /api/recall/search).

    csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/orr.h) -> liborr.so
    shard.py   RecallShard: typed ctypes wrapper over the C ABI (one GPU shard)
    store.py   GpuIngestionStore: mirror of the reference's IIngestionStore
    recall.py  GpuRecallSearchService: mirror of the reference's IRecallSearchService
    sharded.py row-sharded multi-GPU search (torch.distributed all-gather of per-GPU top-k)
    synth.py   synthetic corpora of the benchmark shapes

There is no CPU fallback anywhere in this package.
"""
from . import _native  # noqa: F401
from .cluster import RecallCluster  # noqa: F401
from .shard import BatchHits, BatchTerms, Hits, QueryTerms, RecallShard, hash_term, merge_hits, tokenize_content, tokenize_query  # noqa: F401

__all__ = ["RecallCluster", "BatchHits", "BatchTerms", "Hits", "QueryTerms", "RecallShard", "hash_term", "merge_hits", "tokenize_content", "tokenize_query"]
