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


@pytest.mark.parametrize("world", [2, 4])
def test_fused_peer_memory_exchange_with_virtual_ranks(world):
    """orr_xchg_allgather_merge with `world` ranks emulated on one GPU (same-process attach, one stream per rank so
    the ranks' one-CTA kernels can wait for each other): every rank ends with the oracle's global top-k."""
    import ctypes as C
    import torch

    dim, total, k, kmax = 256, 9_001, 10, 16
    spec = synth.make_spec(dim, gen_dim=dim, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, total)
    L = N.lib()
    shards, xs = [], []
    dev = torch.device("cuda", 0)
    try:
        for r in range(world):
            base, n_local = sharded.shard_rows(total, world, r)
            sh = orr.RecallShard(dim, n_local, row_base=base)
            sh.fill_synthetic(spec, base, n_local)
            shards.append(sh)
            x = C.c_void_p()
            N.check(L.orr_xchg_create(0, world, r, kmax, C.byref(x)))
            xs.append(x)
        for r in range(world):
            for p in range(world):
                if p != r:
                    N.check(L.orr_xchg_attach_peer(xs[r], p, xs[p]))
        streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
        local = [torch.zeros(k * 24, dtype=torch.uint8, device=dev) for _ in range(world)]
        status = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
        out = [torch.zeros(k * 24, dtype=torch.uint8, device=dev) for _ in range(world)]
        out_status = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
        for qi in range(9):                                   # > ORR_XCHG_SLOTS queries: slots are reused
            q = synth.query_host(spec, qi, total, n_terms=4)
            er, es, _ = oracle_search_synth(rows, q, NOW, k)
            q_dev = torch.from_numpy(q.q).to(dev)
            main = torch.cuda.current_stream(dev).cuda_stream
            for r, sh in enumerate(shards):
                sh.search_device(q_dev.data_ptr(), q.terms, NOW, k, local[r].data_ptr(), status[r].data_ptr(), main)
            torch.cuda.synchronize()
            for r in range(world):
                N.check(L.orr_xchg_allgather_merge(xs[r], local[r].data_ptr(), status[r].data_ptr(), k, out[r].data_ptr(),
                                                   out_status[r].data_ptr(), streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(world):
                got, flags = sharded.hits_from_device(out[r], out_status[r])
                assert flags == 0, f"rank {r}: flags {flags}"
                assert_same_ranking(got.rows, got.scores, er, es, what=f"xchg world={world} rank={r} q={qi}")
    finally:
        torch.cuda.synchronize()
        for x in xs:
            L.orr_xchg_destroy(x)
        for sh in shards:
            sh.close()


def test_exchange_reports_a_missing_peer_instead_of_hanging():
    """A rank whose peer never publishes gets status flag ORR_STATUS_XCHG_TIMEOUT after the time-out."""
    import ctypes as C
    import torch

    L = N.lib()
    dev = torch.device("cuda", 0)
    xs = []
    try:
        for r in range(2):
            x = C.c_void_p()
            N.check(L.orr_xchg_create(0, 2, r, 4, C.byref(x)))
            xs.append(x)
        N.check(L.orr_xchg_attach_peer(xs[0], 1, xs[1]))
        N.check(L.orr_xchg_attach_peer(xs[1], 0, xs[0]))
        hits = torch.zeros(4 * 24, dtype=torch.uint8, device=dev)
        st = torch.tensor([0, 0], dtype=torch.int32, device=dev)
        out = torch.zeros(4 * 24, dtype=torch.uint8, device=dev)
        out_st = torch.zeros(2, dtype=torch.int32, device=dev)
        N.check(L.orr_xchg_allgather_merge(xs[0], hits.data_ptr(), st.data_ptr(), 4, out.data_ptr(), out_st.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream))       # rank 1 never calls
        torch.cuda.synchronize()
        assert int(out_st.cpu()[1]) & N.STATUS_XCHG_TIMEOUT
    finally:
        for x in xs:
            L.orr_xchg_destroy(x)


def test_batched_search_over_virtual_shards_merges_to_the_oracle():
    """Row-sharded orr_search_batch: every shard answers every query, the [shards][B][k] lists are merged per
    query on the device (orr_merge_hits_batch_device) -> the oracle's global top-k for each query."""
    import torch

    dim, total, k, B, world = 128, 12_003, 20, 24, 3
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, total)
    qs = [synth.query_host(spec, qi, total, n_terms=3, frequent_terms=1) for qi in range(B)]
    Q = np.stack([q.q for q in qs])
    dev = torch.device("cuda", 0)
    shards = []
    try:
        raws, ns = [], []
        for r in range(world):
            base, n_local = sharded.shard_rows(total, world, r)
            sh = orr.RecallShard(dim, n_local, row_base=base)
            sh.fill_synthetic(spec, base, n_local)
            shards.append(sh)
            got = sh.search_batch(Q, [q.terms for q in qs], NOW, k)
            assert sh.last_timing()["path"] & 0xff == N.PATH_BATCH
            raws.append(got.raw.copy()); ns.append(got.n_out.astype(np.int32))
        allh = torch.from_numpy(np.stack(raws).view(np.uint8).reshape(-1)).to(dev)
        alln = torch.from_numpy(np.stack(ns).reshape(-1)).to(dev)
        out = torch.zeros(B * k * 24, dtype=torch.uint8, device=dev)
        out_n = torch.zeros(B, dtype=torch.int32, device=dev)
        N.check(N.lib().orr_merge_hits_batch_device(0, allh.data_ptr(), alln.data_ptr(), world, B, k, out.data_ptr(),
                                                    out_n.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
        a = np.frombuffer(out.cpu().numpy().tobytes(), dtype=np.dtype([("row", "<u8"), ("score", "<f8"), ("ticks", "<i8")])).reshape(B, k)
        assert out_n.cpu().tolist() == [k] * B
        for b, q in enumerate(qs):
            er, es, _ = oracle_search_synth(rows, q, NOW, k)
            assert_same_ranking(a[b]["row"], a[b]["score"], er, es, what=f"sharded batch b={b}")
    finally:
        for sh in shards:
            sh.close()


def test_single_process_cluster_matches_the_oracle():
    """orr_cluster_*: one process, several shards (here all on GPU 0), fused scan + peer-memory exchange per shard.
    Synthetic corpus split contiguously over the shards -> the oracle's global ranking (global row ids are
    shard << 40 | local row); then document mutations routed to one shard each."""
    dim, per, world, k = 256, 3_000, 3, 10
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, per * world)
    with orr.RecallCluster(dim, per + 64, [0] * world, max_top_k=32) as cl:
        cl.fill_synthetic(spec, 0, per)
        assert cl.count == per * world
        to_global = lambda r: ((int(r) // per) << 40) | (int(r) % per)
        for qi in range(8):
            q = synth.query_host(spec, qi, per * world, n_terms=3)
            got = cl.search(q.q, q.terms, NOW, k)
            er, es, _ = oracle_search_synth(rows, q, NOW, k)
            assert_same_ranking(got.rows, got.scores, [to_global(r) for r in er], es, what=f"cluster q={qi}")
        # no embedding -> per-shard exact path + host merge
        q = synth.query_host(spec, 3, per * world, n_terms=3, frequent_terms=1)
        got = cl.search(None, q.terms, NOW, k)
        q0 = synth.HostQuery(np.zeros(0, np.float32), q.term_ids, q.text, q.terms)
        er, es, _ = oracle_search_synth(rows, q0, NOW, k)
        assert [int(r) for r in got.rows] == [to_global(r) for r in er]
        # mutations: a new document lands on one shard, wins, and disappears again
        q = synth.query_host(spec, 5, per * world, n_terms=2)
        new_rows = cl.upsert_document_chunks(77, q.q[None, :].copy(), np.array([NOW - 864_000_000_000]),
                                             [np.asarray(q.terms.probe_hash, dtype=np.uint64)])
        assert cl.search(q.q, q.terms, NOW, 3).rows[0] == new_rows[0]
        cl.upsert_document_chunks(77, -q.q[None, :].copy(), np.array([NOW - 864_000_000_000]))   # replace: same shard
        assert cl.count == per * world + 1
        cl.delete_document(77)
        assert cl.count == per * world
        got = cl.search(q.q, q.terms, NOW, k)
        er, es, _ = oracle_search_synth(rows, q, NOW, k)
        assert_same_ranking(got.rows, got.scores, [to_global(r) for r in er], es, what="cluster after delete")


def test_pipelined_device_search_equals_the_blocking_form_on_one_rank():
    """search_device_pipelined with world == 1 is search_device plus a completion event (the N>1 form is checked
    against the blocking exchange inside bench.py's multi-GPU run and tools/sharded_check.py)."""
    import torch

    dim, n, k = 256, 5_000, 10
    spec = synth.make_spec(dim, gen_dim=dim)
    with orr.RecallShard(dim, n) as sh:
        sh.fill_synthetic(spec, 0, n)
        sr = sharded.ShardedRecall(sh)
        dev = torch.device("cuda", 0)
        for qi in range(4):
            q = synth.query_host(spec, qi, n, n_terms=4)
            q_dev = torch.from_numpy(q.q).to(dev)
            h, st, done = sr.search_device_pipelined(q_dev, q.terms, NOW, k)
            done.synchronize()
            a, fa = sharded.hits_from_device(h, st)
            b = sh.search(q.q, q.terms, NOW, k)
            assert fa == 0 and a.rows.tolist() == b.rows.tolist() and a.scores.tolist() == b.scores.tolist()
        sr.close()


def test_single_process_cluster_batch_matches_the_oracle():
    """orr_cluster_search_batch: the tcgen05 batched path on every shard (one host thread per GPU) and a per-query k-way
    merge under the reference tie chain -> the oracle's global ranking for every query of the batch, duplicates included."""
    dim, per, world, k, B = 256, 3_000, 3, 10, 24
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, per * world)
    to_global = lambda r: ((int(r) // per) << 40) | (int(r) % per)
    with orr.RecallCluster(dim, per + 64, [0] * world, max_top_k=32) as cl:
        cl.fill_synthetic(spec, 0, per)
        qs = [synth.query_host(spec, 100 + i, per * world, n_terms=1 + i % 4) for i in range(B)]
        got = cl.search_batch(np.stack([q.q for q in qs]), [q.terms for q in qs], NOW, k)
        assert len(got) == B
        for b, q in enumerate(qs):
            er, es, _ = oracle_search_synth(rows, q, NOW, k)
            assert_same_ranking(got[b].rows, got[b].scores, [to_global(r) for r in er], es, what=f"cluster batch b={b}")
        one = cl.search(qs[5].q, qs[5].terms, NOW, k)                       # and the same hits as the single-query form
        assert got[5].rows.tolist() == one.rows.tolist() and got[5].scores.tolist() == one.scores.tolist()


def test_cluster_search_many_pipelines_a_run_of_queries_to_the_same_hits():
    """orr_cluster_search_many: a run of single queries with three in flight (exchange of query i on a side stream while the
    devices scan query i+1) returns, query by query, exactly what orr_cluster_search returns — also for runs shorter than
    the ring, for queries without an embedding (per-shard exact path) and after a mutation."""
    dim, per, world, k = 256, 3_000, 3, 10
    spec = synth.make_spec(dim, gen_dim=dim, terms_per_chunk=16, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, per * world)
    to_global = lambda r: ((int(r) // per) << 40) | (int(r) % per)
    with orr.RecallCluster(dim, per + 64, [0] * world, max_top_k=32) as cl:
        cl.fill_synthetic(spec, 0, per)
        qs = [synth.query_host(spec, 500 + i, per * world, n_terms=i % 4) for i in range(11)]
        for n in (11, 2, 1, 0):
            many = cl.search_many(np.stack([q.q for q in qs[:n]]) if n else np.zeros((0, dim), np.float32), [q.terms for q in qs[:n]], NOW, k)
            assert len(many) == n
            for i in range(n):
                one = cl.search(qs[i].q, qs[i].terms, NOW, k)
                assert many[i].rows.tolist() == one.rows.tolist() and many[i].scores.tolist() == one.scores.tolist(), (n, i)
        many = cl.search_many(np.stack([q.q for q in qs]), [q.terms for q in qs], NOW, k)
        for i in (1, 7, 10):                                                # and the oracle's global ranking
            er, es, _ = oracle_search_synth(rows, qs[i], NOW, k)
            assert_same_ranking(many[i].rows, many[i].scores, [to_global(r) for r in er], es, what=f"many q={i}")
        # larger k than the fused path takes -> the one-at-a-time fallback inside the same call
        big = cl.search_many(np.stack([q.q for q in qs[:3]]), [q.terms for q in qs[:3]], NOW, 40)
        for i in range(3):
            one = cl.search(qs[i].q, qs[i].terms, NOW, 40)
            assert big[i].rows.tolist() == one.rows.tolist()
