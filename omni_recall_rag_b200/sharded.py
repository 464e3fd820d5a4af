"""Row-sharded recall across the GPUs of one box: one process per GPU (torch.distributed),
each rank owns rows [row_base, row_base + n_local) of the corpus, computes its exact local
top-k, and the small candidate lists are all-gathered (NCCL over NVLink on GPUs, gloo in
the CPU tests) and merged with the reference tie chain.

Correctness: every chunk's score depends only on (query, row, now) and the exact score of a
row is computed by the same deterministic function on every rank, so
global top-k  ⊆  ∪ local top-k   and the merge (score desc, ticks desc, global row asc) gives
exactly what the single-shard scorer would return.

The only exchange on the data path is the all-gather of k·24 B per rank.  On GPUs the device-resident
path does it with liborr's own fused kernel (orr_xchg_allgather_merge: every rank stores its list straight
into its peers' HBM over NVLink and merges what arrived, one launch); `exchange="nccl"` keeps the NCCL
all_gather_into_tensor + merge-kernel form for comparison.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

from . import _native as N
from .shard import BatchHits, BatchTerms, Hits, QueryTerms, RecallShard, merge_hits, _HIT_DTYPE

HIT_BYTES = 24  # sizeof(orr_hit)


def shard_rows(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: (row_base, n_local).  Blocks differ by at most 1 row."""
    base, rem = divmod(total_rows, world_size)
    n_local = base + (1 if rank < rem else 0)
    row_base = rank * base + min(rank, rem)
    return row_base, n_local


def _pack(h: Hits, k: int) -> np.ndarray:
    """Hits -> int64[k, 3] (row bits, score bits, ticks) + valid count in the last slot."""
    buf = np.zeros((k + 1, 3), dtype=np.int64)
    n = len(h)
    buf[:n, 0] = h.rows.view(np.int64)
    buf[:n, 1] = h.scores.view(np.int64)
    buf[:n, 2] = h.ticks
    buf[k, 0] = n
    return buf


def _unpack(buf: np.ndarray, k: int) -> Hits:
    n = int(buf[k, 0])
    return Hits(buf[:n, 0].copy().view(np.uint64), buf[:n, 1].copy().view(np.float64), buf[:n, 2].copy())


class ShardedRecall:
    """Search over a row-sharded corpus.  `local_search(q, terms, now_ticks, top_k) -> Hits`
    defaults to this rank's RecallShard; tests inject a CPU scorer to exercise the
    gather/merge logic under gloo."""

    def __init__(self, shard: Optional[RecallShard] = None, *, group=None,
                 local_search: Optional[Callable[..., Hits]] = None, exchange: str = "p2p", max_top_k: int = 128):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.shard = shard
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if local_search is None:
            if shard is None:
                raise ValueError("ShardedRecall needs a RecallShard or a local_search callable")
            local_search = lambda q, terms, now, k: shard.search(q, terms, now, k)  # noqa: E731
        self.local_search = local_search
        self._dev_bufs = {}
        self._xchg = None
        self.max_top_k = max_top_k
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange if (self.world > 1 and shard is not None) else "none"
        if self.exchange == "p2p" and dist.get_backend(group) != "nccl":
            self.exchange = "nccl"                       # CPU tests (gloo): there is no peer memory
        if self.exchange == "p2p":
            self._open_exchange()

    def _open_exchange(self) -> None:
        """One exchange buffer per rank; the 64-byte CUDA IPC handles travel over torch.distributed."""
        import torch

        L = N.lib()
        x = C.c_void_p()
        N.check(L.orr_xchg_create(self.shard.device, self.world, self.rank, self.max_top_k, C.byref(x)))
        self._xchg = x
        mine = (C.c_ubyte * N.XCHG_HANDLE_BYTES)()
        N.check(L.orr_xchg_get_handle(x, C.cast(mine, C.c_void_p)))
        dev = torch.device("cuda", self.shard.device)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        allh = torch.empty((self.world, N.XCHG_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(allh, t, group=self.group)
        allh = allh.cpu().numpy()
        for r in range(self.world):
            if r != self.rank:
                h = (C.c_ubyte * N.XCHG_HANDLE_BYTES)(*allh[r].tolist())
                N.check(L.orr_xchg_open_peer(x, r, C.cast(h, C.c_void_p)))
        self.dist.barrier(group=self.group)            # every rank has every peer mapped before the first push

    def resync(self) -> None:
        """Recovery after an exchange time-out (a peer issued a different number of searches): a collective.  Every
        rank drains its device, the ranks agree on a sequence number above any in use, and a barrier precedes the next
        exchange (orr_xchg_resync)."""
        if self._xchg is None:
            return
        import torch

        torch.cuda.synchronize()
        dev = torch.device("cuda", self.shard.device)
        t = torch.tensor([int(N.lib().orr_xchg_sequence(self._xchg))], dtype=torch.int64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        N.check(N.lib().orr_xchg_resync(self._xchg, (int(t.item()) + 16) & 0x7fffffff))
        self.dist.barrier(group=self.group)

    def set_exchange_timeout_ms(self, ms: float) -> None:
        if self._xchg is not None:
            N.check(N.lib().orr_xchg_set_timeout_ms(self._xchg, float(ms)))

    def last_timing(self) -> dict:
        """Kernel durations (CUDA events) of this rank's part of the last search()."""
        return dict(getattr(self, "_last_timing", {}))

    def close(self) -> None:
        if self._xchg is not None:
            import torch

            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)        # no peer may still be pushing into this rank's buffer
            N.lib().orr_xchg_destroy(self._xchg)
            self._xchg = None

    # -- host-buffer path (end-to-end: query in host memory, hits back in host memory) --------
    def search(self, q: Optional[np.ndarray], terms: QueryTerms, now_ticks: int, top_k: int) -> Hits:
        import torch

        k = max(1, int(top_k))
        if (self.world > 1 and self._xchg is not None and q is not None and len(q) == self.shard.dim
                and k <= min(self.max_top_k, 224)):
            # GPUs: query up once, fused scan + exact re-score + peer-memory exchange on the device, hits down once
            dev = torch.device("cuda", self.shard.device)
            q_dev = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(dev, non_blocking=True)
            hits_dev, status_dev = self.search_device(q_dev, terms, now_ticks, top_k)
            hits, flags = hits_from_device(hits_dev, status_dev)
            self._last_timing = self.shard.last_device_timing()
            if flags == 0:
                return hits
            if flags & N.STATUS_XCHG_TIMEOUT:
                raise RuntimeError("sharded search: a peer rank never published its hit list (exchange time-out)")
            # a shard could not prove its fp32 selection: every rank sees the same OR-ed flag and re-runs below,
            # where orr_search escalates to the exact path
        local = self.local_search(q, terms, now_ticks, top_k)
        self._last_timing = self.shard.last_timing() if self.shard is not None else {}
        if self.world == 1:
            return local
        mine = torch.from_numpy(_pack(local, k))
        backend = self.dist.get_backend(self.group)
        if backend == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            mine = mine.to(dev)
            gathered = torch.empty((self.world, k + 1, 3), dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
            gathered = gathered.cpu().numpy()
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.group)
            gathered = np.stack([p.numpy() for p in parts])
        return merge_hits([_unpack(gathered[r], k) for r in range(self.world)], top_k)

    # -- batched queries over the sharded corpus -------------------------------------------------------
    def search_batch(self, q: np.ndarray, terms, now_ticks: int, top_k: int) -> BatchHits:
        """Every rank runs orr_search_batch (tcgen05 contraction + exact re-rank) on its shard; the B x k x 24 B
        answers are all-gathered (a bandwidth-type exchange: NCCL) and merged per query on the device
        (orr_merge_hits_batch_device, one CTA per query) with the reference tie chain."""
        import torch

        if self.world == 1:
            return self.shard.search_batch(q, terms, now_ticks, top_k)
        if self.dist.get_backend(self.group) != "nccl":
            raise RuntimeError("sharded search_batch needs the nccl backend (device merge)")
        dev = torch.device("cuda", self.shard.device)
        B, k = int(np.asarray(q).shape[0]), max(1, int(top_k))
        bt = terms if (terms is None or isinstance(terms, BatchTerms)) else BatchTerms.pack(terms)
        key = ("batch", B, k)
        if key not in self._dev_bufs:                      # device buffers and the pinned landing zone, once per shape
            nb = B * k * HIT_BYTES
            self._dev_bufs[key] = dict(
                mine=torch.empty(nb, dtype=torch.uint8, device=dev), mine_n=torch.empty(B, dtype=torch.int32, device=dev),
                allh=torch.empty(self.world * nb, dtype=torch.uint8, device=dev),
                alln=torch.empty(self.world * B, dtype=torch.int32, device=dev),
                out=torch.empty(nb, dtype=torch.uint8, device=dev), out_n=torch.empty(B, dtype=torch.int32, device=dev),
                h_out=torch.empty(nb, dtype=torch.uint8).pin_memory(), h_n=torch.empty(B, dtype=torch.int32).pin_memory())
        bufs = self._dev_bufs[key]
        mine, mine_n = bufs["mine"], bufs["mine_n"]
        # the local answers stay in HBM and go straight into the all-gather (orr_search_batch_device); a batch that path
        # does not take (or could not prove) comes back through host memory instead
        if not self.shard.search_batch_device(q, bt, now_ticks, top_k, mine.data_ptr(), mine_n.data_ptr()):
            local = self.shard.search_batch(q, bt, now_ticks, top_k)
            mine.copy_(torch.from_numpy(local.raw.view(np.uint8).reshape(-1)), non_blocking=True)
            mine_n.copy_(torch.from_numpy(np.ascontiguousarray(local.n_out, dtype=np.int32)), non_blocking=True)
        self.dist.all_gather_into_tensor(bufs["allh"], mine, group=self.group)
        self.dist.all_gather_into_tensor(bufs["alln"], mine_n, group=self.group)
        N.check(N.lib().orr_merge_hits_batch_device(self.shard.device, bufs["allh"].data_ptr(), bufs["alln"].data_ptr(), self.world, B, k,
                                                    bufs["out"].data_ptr(), bufs["out_n"].data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
        bufs["h_out"].copy_(bufs["out"], non_blocking=True)
        bufs["h_n"].copy_(bufs["out_n"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        raw = bufs["h_out"].numpy().view(_HIT_DTYPE).reshape(B, k).copy()     # one host copy: the pinned buffer is reused
        return BatchHits(raw, bufs["h_n"].numpy().copy())

    def search_batch_async(self, q: np.ndarray, terms, now_ticks: int, top_k: int) -> Callable[[], BatchHits]:
        """Throughput form of search_batch for a stream of batches: the local tcgen05 search runs now (the call returns when
        this rank's answers are in HBM), the all-gather + per-query merge + copy to pinned memory are enqueued on a side
        stream, and the returned function waits for them and hands out the BatchHits.  Calling it for batch i+1 before
        collecting batch i overlaps i's exchange — and its wait for the slowest rank — with i+1's search.  At most two
        batches may be outstanding (two buffer sets); every rank must issue the same sequence of batches."""
        import torch

        if self.world == 1:
            local = self.shard.search_batch(q, terms, now_ticks, top_k)
            return lambda: local
        if self.dist.get_backend(self.group) != "nccl":
            raise RuntimeError("sharded search_batch needs the nccl backend (device merge)")
        dev = torch.device("cuda", self.shard.device)
        B, k = int(np.asarray(q).shape[0]), max(1, int(top_k))
        bt = terms if (terms is None or isinstance(terms, BatchTerms)) else BatchTerms.pack(terms)
        key = ("batch_async", B, k)
        if key not in self._dev_bufs:
            nb = B * k * HIT_BYTES
            self._dev_bufs[key] = dict(side=torch.cuda.Stream(device=dev), idx=0, slots=[dict(
                mine=torch.empty(nb, dtype=torch.uint8, device=dev), mine_n=torch.empty(B, dtype=torch.int32, device=dev),
                allh=torch.empty(self.world * nb, dtype=torch.uint8, device=dev),
                alln=torch.empty(self.world * B, dtype=torch.int32, device=dev),
                out=torch.empty(nb, dtype=torch.uint8, device=dev), out_n=torch.empty(B, dtype=torch.int32, device=dev),
                h_out=torch.empty(nb, dtype=torch.uint8).pin_memory(), h_n=torch.empty(B, dtype=torch.int32).pin_memory(),
                done=torch.cuda.Event(), pending=False) for _ in range(2)])
        pipe = self._dev_bufs[key]
        b = pipe["slots"][pipe["idx"] % 2]
        pipe["idx"] += 1
        if b["pending"]:
            raise RuntimeError("search_batch_async: collect the batch issued two calls ago before issuing another")
        if not self.shard.search_batch_device(q, bt, now_ticks, top_k, b["mine"].data_ptr(), b["mine_n"].data_ptr()):
            local = self.shard.search_batch(q, bt, now_ticks, top_k)
            b["mine"].copy_(torch.from_numpy(local.raw.view(np.uint8).reshape(-1)))
            b["mine_n"].copy_(torch.from_numpy(np.ascontiguousarray(local.n_out, dtype=np.int32)))
            torch.cuda.current_stream(dev).synchronize()
        side = pipe["side"]                                   # the local answers are complete (the call above synchronised)
        with torch.cuda.stream(side):
            self.dist.all_gather_into_tensor(b["allh"], b["mine"], group=self.group)
            self.dist.all_gather_into_tensor(b["alln"], b["mine_n"], group=self.group)
            N.check(N.lib().orr_merge_hits_batch_device(self.shard.device, b["allh"].data_ptr(), b["alln"].data_ptr(), self.world, B, k,
                                                        b["out"].data_ptr(), b["out_n"].data_ptr(), side.cuda_stream))
            b["h_out"].copy_(b["out"], non_blocking=True)
            b["h_n"].copy_(b["out_n"], non_blocking=True)
            b["done"].record(side)
        b["pending"] = True

        def collect() -> BatchHits:
            b["done"].synchronize()
            b["pending"] = False
            return BatchHits(b["h_out"].numpy().view(_HIT_DTYPE).reshape(B, k).copy(), b["h_n"].numpy().copy())
        return collect

    # -- device-resident path (query and hits stay in HBM; nothing synchronises the host) -----
    def search_device(self, q_dev, terms: QueryTerms, now_ticks: int, top_k: int):
        """q_dev: torch float32 CUDA tensor [dim].  Returns (hits_dev uint8[k*24], status_dev
        int32[2] = {n_out, flags}) on the current stream: local fused scan + exact re-score
        (orr_search_device), then the exchange: the fused peer-memory all-gather + merge kernel
        (orr_xchg_allgather_merge), or with exchange="nccl" all_gather_into_tensor of the k·24 B lists
        followed by the one-CTA merge (orr_merge_hits_device)."""
        import torch

        k = max(1, int(top_k))
        key = (k, q_dev.device.index)
        if key not in self._dev_bufs:
            dev = q_dev.device
            self._dev_bufs[key] = dict(
                hits=torch.zeros(k * HIT_BYTES, dtype=torch.uint8, device=dev),
                status=torch.zeros(2, dtype=torch.int32, device=dev),
                all_hits=torch.zeros(self.world * k * HIT_BYTES, dtype=torch.uint8, device=dev),
                all_status=torch.zeros(self.world * 2, dtype=torch.int32, device=dev),
                out_hits=torch.zeros(k * HIT_BYTES, dtype=torch.uint8, device=dev),
                out_status=torch.zeros(2, dtype=torch.int32, device=dev))
        b = self._dev_bufs[key]
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        self.shard.search_device(q_dev.data_ptr(), terms, now_ticks, top_k, b["hits"].data_ptr(),
                                 b["status"].data_ptr(), stream)
        if self.world == 1:
            return b["hits"], b["status"]
        if self._xchg is not None:
            if k > self.max_top_k:
                raise ValueError(f"top_k {k} > max_top_k {self.max_top_k} of the exchange buffers")
            N.check(N.lib().orr_xchg_allgather_merge(self._xchg, b["hits"].data_ptr(), b["status"].data_ptr(), top_k,
                                                     b["out_hits"].data_ptr(), b["out_status"].data_ptr(), stream))
            return b["out_hits"], b["out_status"]
        self.dist.all_gather_into_tensor(b["all_hits"], b["hits"], group=self.group)
        self.dist.all_gather_into_tensor(b["all_status"], b["status"], group=self.group)
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        N.check(N.lib().orr_merge_hits_device(
            q_dev.device.index, b["all_hits"].data_ptr(), b["all_status"].data_ptr(), self.world, k, top_k,
            b["out_hits"].data_ptr(), b["out_status"].data_ptr(), stream))
        return b["out_hits"], b["out_status"]


    def search_device_pipelined(self, q_dev, terms: QueryTerms, now_ticks: int, top_k: int, depth: int = 3):
        """Throughput form of search_device for a stream of queries: the exchange of query i runs on a side
        stream, so this rank's scan of query i+1 starts as soon as its own local top-k of query i exists
        instead of waiting for the slowest peer.  Every query is still a complete search (local exact top-k,
        all-gather over peer memory, merge); only the waiting moves off the scan's stream.  Returns
        (hits_dev, status_dev, done_event): the buffers are valid once `done_event` has completed and stay
        untouched for the next `depth - 1` calls.  The exchange ring (ORR_XCHG_SLOTS = 4 sequence-tagged slots
        per rank) bounds how far ranks may drift apart; kernels of one rank's exchanges run in order."""
        import torch

        if self.world == 1 or self._xchg is None:
            h, st = self.search_device(q_dev, terms, now_ticks, top_k)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(q_dev.device))
            return h, st, ev
        k = max(1, int(top_k))
        if k > self.max_top_k:
            raise ValueError(f"top_k {k} > max_top_k {self.max_top_k} of the exchange buffers")
        dev = q_dev.device
        key = ("pipe", k, dev.index, depth)
        if key not in self._dev_bufs:
            self._dev_bufs[key] = dict(
                side=torch.cuda.Stream(device=dev), idx=0,
                slots=[dict(hits=torch.zeros(k * HIT_BYTES, dtype=torch.uint8, device=dev),
                            status=torch.zeros(2, dtype=torch.int32, device=dev),
                            out_hits=torch.zeros(k * HIT_BYTES, dtype=torch.uint8, device=dev),
                            out_status=torch.zeros(2, dtype=torch.int32, device=dev),
                            ready=torch.cuda.Event(), done=torch.cuda.Event(), used=False) for _ in range(max(2, depth))])
        pipe = self._dev_bufs[key]
        b = pipe["slots"][pipe["idx"] % len(pipe["slots"])]
        pipe["idx"] += 1
        main = torch.cuda.current_stream(dev)
        if b["used"]:
            main.wait_event(b["done"])                 # the slot's previous exchange has read hits/status
        self.shard.search_device(q_dev.data_ptr(), terms, now_ticks, top_k, b["hits"].data_ptr(),
                                 b["status"].data_ptr(), main.cuda_stream)
        b["ready"].record(main)
        side = pipe["side"]
        side.wait_event(b["ready"])
        N.check(N.lib().orr_xchg_allgather_merge(self._xchg, b["hits"].data_ptr(), b["status"].data_ptr(), top_k,
                                                 b["out_hits"].data_ptr(), b["out_status"].data_ptr(), side.cuda_stream))
        b["done"].record(side)
        b["used"] = True
        return b["out_hits"], b["out_status"], b["done"]


def hits_from_device(hits_dev, status_dev) -> Tuple[Hits, int]:
    """D2H of a device hit list -> (Hits, flags)."""
    st = status_dev.cpu().numpy()
    raw = hits_dev.cpu().numpy()
    a = np.frombuffer(raw.tobytes(), dtype=np.dtype([("row", "<u8"), ("score", "<f8"), ("ticks", "<i8")]))
    n = int(st[0])
    return Hits(a["row"][:n].copy(), a["score"][:n].copy(), a["ticks"][:n].copy()), int(st[1])
