#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/sharded_check.py — parity of the NCCL row-sharded path.

Each rank owns a row block of one synthetic corpus on its own GPU; every rank must return
exactly what the CPU oracle returns for the whole corpus, through both the host-buffer path
(ShardedRecall.search) and the device-resident path (search_device + NCCL all-gather +
device merge)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import omni_recall_rag_b200 as orr  # noqa: E402
from omni_recall_rag_b200 import sharded, synth  # noqa: E402
from tests.util import assert_same_ranking, oracle_search_synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dim, total, k = 3072, 40_003, 10
    spec = synth.make_spec(dim, dup_row_ppm=5000)
    base, n_local = sharded.shard_rows(total, world, rank)
    sh = orr.RecallShard(dim, n_local, device=local, row_base=base)
    sh.fill_synthetic(spec, base, n_local)
    sr = sharded.ShardedRecall(sh)
    rows = synth.rows_host(spec, 0, total) if rank == 0 else None
    ok = 0
    for qi in range(8):
        q = synth.query_host(spec, qi, total, n_terms=4)
        got = sr.search(q.q, q.terms, spec.now_ticks, k)
        hd, sd = sr.search_device(torch.from_numpy(q.q).to(dev), q.terms, spec.now_ticks, k)
        torch.cuda.synchronize()
        got_dev, flags = sharded.hits_from_device(hd, sd)
        assert flags == 0
        assert got_dev.rows.tolist() == got.rows.tolist() and got_dev.scores.tolist() == got.scores.tolist()
        # every rank holds the same answer
        t = torch.from_numpy(got.rows.astype(np.int64)).to(dev)
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        if rank == 0:
            er, es, _ = oracle_search_synth(rows, q, spec.now_ticks, k)
            assert_same_ranking(got.rows, got.scores, er, es, what=f"sharded x{world} q={qi}")
        ok += 1
    dist.barrier()
    if rank == 0:
        print(f"sharded parity ok: world={world}, {ok} queries, {total} x {dim}")
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
