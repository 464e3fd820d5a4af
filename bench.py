#!/usr/bin/env python
"""bench.py — hybrid recall QPS at 1M x 3072 fp32, top-10 (BASELINE.json's metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one hybrid query (0.7 cosine + 0.2 keyword + 0.1 recency, 4 query terms, top-10)
over the whole HBM-resident corpus.  At N=1 the corpus is BASELINE.json configs[1]
(1M chunks x 3072 fp32, ~12.3 GB).  For N>1 the corpus is row-sharded, 1M rows per GPU (weak
scaling; --rows-per-gpu 5000000 gives configs[3], 40M rows on 8 GPUs): every rank scans its
shard, computes its exact local top-10, and the lists are all-gathered over NCCL and merged.

    value  = (total rows / 1M) x queries / second, device-timed (CUDA events, max over ranks),
             query already in HBM: "1M x 3072 scans per second".  At N=1 this is plain QPS.
    e2e    = the same through the public host API (ShardedRecall.search -> orr_search): the
             query is copied from pinned host memory and the hits are read back every step.
    roofline = algorithmic bytes of the scan kernel / its CUDA-event duration vs the
             measured HBM copy bandwidth (MEASURED_PEAKS.json).
    cpu_baseline = the C oracle (a port of the reference's C# scorer; the C# cannot run
             here) timed on this box's host cores on a bounded sample of the same corpus.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "hybrid recall QPS at 1M x 3072 fp32, top-10; HBM GB/s vs ~8 TB/s peak"
UNIT = "queries/s (1M x 3072 scans per second)"
DIM = 3072
TOP_K = 10
N_TERMS = 4
TERM_SLOTS = 64
FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.samples.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
                for name, v in zip(names, s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(sample_rows: int, n_queries: int, threads: int, corpus_rows: int, min_seconds: float = 0.0):
    """Times the oracle port on `sample_rows` rows of the same synthetic corpus; returns
    1M x 3072 scans per second (linear in rows: the scorer is a per-row loop)."""
    import numpy as np  # noqa: F401
    from omni_recall_rag_b200 import synth
    from oracle import oracle_c

    spec = synth.make_spec(DIM)
    rows = synth.rows_host(spec, 0, sample_rows)
    blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
    queries = [synth.query_host(spec, qi, corpus_rows, n_terms=N_TERMS) for qi in range(n_queries)]
    oracle_c.search(emb=rows.emb, dim=DIM, ticks=rows.ticks, content_blob=blob, content_off=off,
                    query=queries[0].text, qvec=queries[0].q, now_ticks=spec.now_ticks, top_k=TOP_K, threads=threads)
    t0 = time.perf_counter()
    done = 0
    while True:                       # bounded by time: keep going until ~min_seconds of CPU work
        q = queries[done % n_queries]
        oracle_c.search(emb=rows.emb, dim=DIM, ticks=rows.ticks, content_blob=blob, content_off=off,
                        query=q.text, qvec=q.q, now_ticks=spec.now_ticks, top_k=TOP_K, threads=threads)
        done += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds and done >= n_queries:
            break
    return (done / dt) * (sample_rows / 1.0e6), dt, done


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The C# cannot be
    built here (no dotnet), so this is the oracle port on all host threads; each step is one
    query over a bounded sample of the corpus, scaled linearly to 1M rows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c

    threads = oracle_c.max_threads()
    sample = args.cpu_sample_rows
    steps = max(1, args.steps)
    from omni_recall_rag_b200 import synth

    spec = synth.make_spec(DIM)
    rows = synth.rows_host(spec, 0, sample)
    blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
    total_rows = args.rows_per_gpu * args.gpus
    qs = [synth.query_host(spec, qi, total_rows, n_terms=N_TERMS) for qi in range(steps + args.warmup)]

    def one(q):
        oracle_c.search(emb=rows.emb, dim=DIM, ticks=rows.ticks, content_blob=blob, content_off=off, query=q.text,
                        qvec=q.q, now_ticks=spec.now_ticks, top_k=TOP_K, threads=threads)

    for q in qs[: args.warmup]:
        one(q)
    t0 = time.perf_counter()
    for q in qs[args.warmup:]:
        one(q)
    dt = time.perf_counter() - t0
    value = (steps / dt) * (sample / 1.0e6)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / steps * (1.0e6 / sample), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} queries x {sample} rows x {DIM} (first rows of the same synthetic corpus), "
                                   f"scaled linearly to 1M rows; C port of RecallSearchService.cs (dotnet absent)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    total = args.rows_per_gpu * args.gpus
    return {"workload": f"{total} chunks x {DIM} fp32 row-major in HBM ({args.rows_per_gpu} rows/GPU), single query, "
                        f"hybrid 0.7/0.2/0.1, {N_TERMS} query terms, {TERM_SLOTS} hashed terms/chunk, top-{TOP_K}",
            "rows_total": total, "rows_per_gpu": args.rows_per_gpu, "dim": DIM, "top_k": TOP_K,
            "parallelism": f"row-sharded x{args.gpus}, NCCL all-gather of per-GPU top-{TOP_K}" if args.gpus > 1 else "1 GPU",
            "l2": "inputs (12.3 GB/GPU) exceed L2 (126 MB); no flush needed",
            "value_units": "queries/s x (rows_total / 1M)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborr has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    n_local = args.rows_per_gpu
    total_rows = n_local * world
    row_base = rank * n_local
    spec = synth.make_spec(DIM)
    shard = orr.RecallShard(DIM, n_local, device=local_rank, term_slots=TERM_SLOTS, row_base=row_base)
    shard.fill_synthetic(spec, row_base, n_local)
    sr = sharded.ShardedRecall(shard)

    n_q = steps + warmup
    queries = [synth.query_host(spec, qi, total_rows, n_terms=N_TERMS) for qi in range(n_q)]
    q_dev = [torch.from_numpy(q.q).to(dev) for q in queries]
    q_pinned = [torch.from_numpy(q.q).pin_memory() for q in queries]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    flags_seen = 0
    for i in range(warmup):
        sr.search_device(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    status_log = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for i in range(warmup, n_q):
            _, st = sr.search_device(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
        ev1.record()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
        # keep sampling under load for a moment longer so short runs still get samples
        t_end = time.time() + 0.6
        j = warmup
        while time.time() < t_end:
            sr.search_device(q_dev[j], queries[j].terms, spec.now_ticks, TOP_K)
            j = warmup + (j + 1 - warmup) % steps
        torch.cuda.synchronize()
    clock_summary = clocks.summary()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- end-to-end timing through the host API: `e2e`, and the scan kernel's own duration ----
    for i in range(warmup):
        sr.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K)
    barrier()
    scan_ms, fin_ms, escalated = [], [], 0
    t0 = time.perf_counter()
    for i in range(warmup, n_q):
        hits = sr.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K)
        tm = shard.last_timing()
        scan_ms.append(tm["scan_ms"]); fin_ms.append(tm["finalize_ms"])
        escalated += 1 if (tm["path"] & 0x100) else 0
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # sanity: the device-resident path and the host path return the same hits
    hd, sd = sr.search_device(q_dev[n_q - 1], queries[n_q - 1].terms, spec.now_ticks, TOP_K)
    torch.cuda.synchronize()
    got, flags = sharded.hits_from_device(hd, sd)
    assert got.rows.tolist() == hits.rows.tolist() and got.scores.tolist() == hits.scores.tolist(), "device/host paths disagree"
    flags_seen |= flags

    scale = total_rows / 1.0e6
    value = steps / (dev_ms / 1000.0) * scale
    e2e_value = steps / e2e_s * scale
    peak, peak_kind = measured_peak()
    bytes_per_launch = n_local * (4 * DIM + 8 + 4 * TERM_SLOTS)
    scan_avg_ms = sum(scan_ms) / len(scan_ms)
    achieved = bytes_per_launch / (scan_avg_ms / 1000.0) / 1.0e9
    launches_per_step = 2 + (1 if world > 1 else 0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 scan select + f64 exact re-score", "data": "synthetic",
        "config": workload_config(args),
        "corpus_qps": steps / (dev_ms / 1000.0),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * DIM + 12 * N_TERMS,
                "d2h_bytes_per_step": 24 * TOP_K + 8, "ms_per_step": 1000.0 * e2e_s / steps},
        "gpu_launches": launches_per_step * steps,
        "kernels_per_step": ["orr_scan_kernel<24,1>", "orr_rescore_kernel"] + (["orr_merge_kernel"] if world > 1 else []),
        "roofline": {"bound": "hbm", "kernel": "orr_scan_kernel<24,1>", "achieved": achieved, "peak": peak,
                     "peak_kind": f"{peak_kind} HBM copy GB/s (MEASURED_PEAKS.json)" if peak_kind == "measured" else "fallback 6650 GB/s",
                     "unit": "GB/s", "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0,
                     "bytes_per_launch": bytes_per_launch, "bytes_per_launch_emb_only": n_local * 4 * DIM,
                     "kernel_ms": scan_avg_ms, "finalize_kernel_ms": sum(fin_ms) / len(fin_ms),
                     "traffic": None},
        "clocks": clock_summary,
        "bound_check_escalations": escalated, "device_flags": flags_seen,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        v_all, dt_all, n_all = cpu_baseline_run(args.cpu_sample_rows, 8, threads, total_rows, min_seconds=10.0)
        v_one, dt_one, n_one = cpu_baseline_run(args.cpu_sample_rows, 3, 1, total_rows, min_seconds=8.0)
        line["cpu_baseline"] = {
            "value": v_all, "unit": UNIT, "cores": threads, "kind": "port",
            "value_1_thread": v_one,
            "sample": f"{n_all} queries x {args.cpu_sample_rows} rows x {DIM} (first rows of the same corpus) on {threads} "
                      f"threads ({dt_all:.1f} s) and {n_one} queries on 1 thread ({dt_one:.1f} s), scaled linearly to 1M "
                      f"rows; C port of RecallSearchService.cs:20-119 (no dotnet in the image)"}
    if rank == 0:
        print(json.dumps(line))
    shard.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
