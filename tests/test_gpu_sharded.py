"""Row-sharded search on real GPUs.  With one GPU the shards are emulated in one process
(several RecallShard objects on cuda:0 + orr_merge_hits / orr_merge_hits_device); the
torchrun script tools/sharded_check.py covers the NCCL path on 2+ GPUs."""
import numpy as np
import pytest

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import sharded, synth
from tests.util import assert_same_ranking, oracle_search_synth

pytestmark = pytest.mark.gpu
NOW = synth.NOW_TICKS


@pytest.mark.parametrize("world", [2, 8])
def test_virtual_ranks_on_one_gpu_match_the_single_shard_oracle(world):
    import torch

    dim, total, k = 768, 24_001, 10
    spec = synth.make_spec(dim, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, total)
    shards = []
    for r in range(world):
        base, n_local = sharded.shard_rows(total, world, r)
        sh = orr.RecallShard(dim, n_local, row_base=base)
        sh.fill_synthetic(spec, base, n_local)
        shards.append(sh)
    try:
        dev = torch.device("cuda", 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        for qi in range(5):
            q = synth.query_host(spec, qi, total, n_terms=4)
            er, es, et = oracle_search_synth(rows, q, NOW, k)
            # host merge of the per-shard lists
            lists = [sh.search(q.q, q.terms, NOW, k) for sh in shards]
            merged = orr.merge_hits(lists, k)
            assert_same_ranking(merged.rows, merged.scores, er, es, what=f"host merge q={qi}")
            # device path: per-shard device search, concatenated like an all-gather, device merge
            q_dev = torch.from_numpy(q.q).to(dev)
            all_hits = torch.zeros(world * k * 24, dtype=torch.uint8, device=dev)
            all_status = torch.zeros(world * 2, dtype=torch.int32, device=dev)
            for r, sh in enumerate(shards):
                sh.search_device(q_dev.data_ptr(), q.terms, NOW, k, all_hits[r * k * 24:].data_ptr(),
                                 all_status[r * 2:].data_ptr(), stream)
            out_hits = torch.zeros(k * 24, dtype=torch.uint8, device=dev)
            out_status = torch.zeros(2, dtype=torch.int32, device=dev)
            N.check(N.lib().orr_merge_hits_device(0, all_hits.data_ptr(), all_status.data_ptr(), world, k, k,
                                                  out_hits.data_ptr(), out_status.data_ptr(), stream))
            torch.cuda.synchronize()
            got, flags = sharded.hits_from_device(out_hits, out_status)
            assert flags == 0
            assert got.rows.tolist() == merged.rows.tolist() and got.scores.tolist() == merged.scores.tolist()
    finally:
        for sh in shards:
            sh.close()
