#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/sharded_check.py — parity of the row-sharded path (both exchanges).

Each rank owns a row block of one synthetic corpus on its own GPU; every rank must return
exactly what the CPU oracle returns for the whole corpus, through both the host-buffer path
(ShardedRecall.search) and the device-resident path (search_device + the fused peer-memory
all-gather/merge kernel, and the NCCL all-gather + device merge form), which are also timed."""
import os
import sys
import time

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
    sr = sharded.ShardedRecall(sh, exchange="p2p")
    sr_nccl = sharded.ShardedRecall(sh, exchange="nccl")
    assert sr.exchange == "p2p" and sr_nccl.exchange == "nccl"
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
        hd2, sd2 = sr_nccl.search_device(torch.from_numpy(q.q).to(dev), q.terms, spec.now_ticks, k)
        torch.cuda.synchronize()
        got_nccl, flags2 = sharded.hits_from_device(hd2, sd2)
        assert flags2 == 0 and got_nccl.rows.tolist() == got.rows.tolist() and got_nccl.scores.tolist() == got.scores.tolist()
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
        print(f"sharded parity ok: world={world}, {ok} queries, {total} x {dim}, exchanges p2p + nccl")
    # batched queries over the sharded corpus: all-gather of the B x k answers + per-query device merge
    B, kb = 32, 20
    qs = [synth.query_host(spec, 100 + i, total, n_terms=4) for i in range(B)]
    Q = np.stack([q.q for q in qs])
    gotb = sr.search_batch(Q, [q.terms for q in qs], spec.now_ticks, kb)
    if rank == 0:
        for b in range(0, B, 3):
            er, es, _ = oracle_search_synth(rows, qs[b], spec.now_ticks, kb)
            assert_same_ranking(gotb[b].rows, gotb[b].scores, er, es, what=f"sharded batch x{world} b={b}")
        print(f"sharded batch parity ok: {B} queries, top-{kb}")
    # exchange cost: the same 200 device-resident queries through each exchange (small shard => exchange-dominated)
    q = synth.query_host(spec, 3, total, n_terms=4)
    qd = torch.from_numpy(q.q).to(dev)
    for name, s_ in (("p2p", sr), ("nccl", sr_nccl), ("p2p", sr), ("nccl", sr_nccl)):
        for _ in range(20):
            s_.search_device(qd, q.terms, spec.now_ticks, k)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            s_.search_device(qd, q.terms, spec.now_ticks, k)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"  {name:5s}: {t.item() * 1000:.1f} us per query (scan + re-score + exchange, {n_local} rows/GPU, max over ranks)")
    sr.close()
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
