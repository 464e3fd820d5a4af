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

--workload c3 | c5 (N=1 only) benches the batched path instead — BASELINE.json configs[2] and
configs[4]: 5M x 768, batch 1024 top-100 (4 terms) / batch 256 top-50 (16 terms, tie-heavy) — a
"step" is one orr_search_batch call; roofline.bound is "tensor" (useful 2*N*D*B flops of the main
tcgen05 pass against the measured bf16 throughput).  The default line stays the headline metric.

--workload c1 (N=1) is BASELINE.json configs[0], the reference's own CPU-runnable case, in full: 10k x 3072,
single query, top-10, with the reference's 300-most-recent pre-selection (InMemoryIngestionStore.cs:57-65,
candidate_cap=300 — the line's value) and over all rows (cap=0, under "all_rows"); the CPU port runs the
whole workload (no sampling) on 1 thread (the reference's own execution) and on all threads.
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
FALLBACK_BF16_TFLOPS = 1500.0
# dram__bytes_read.sum + dram__bytes_write.sum of orr_scan_kernel<24,1,0> per launch at 1M x 3072, 4 terms,
# from the ncu --set full capture summarised in profiles/r01_final_kernels.md (12.5533 GB read + 3.6 MB
# written); scales linearly with rows.
NCU_SCAN_TRAFFIC_BYTES_PER_ROW = 12556.9
BATCH_WORKLOADS = {
    "c3": dict(rows=5_000_000, dim=768, batch=1024, top_k=100, n_terms=4, frequent=0, dup_ppm=0,
               name="5M chunks x 768 fp32 (truncated embeddings), batch 1024 queries, 4 terms, top-100"),
    "c5": dict(rows=5_000_000, dim=768, batch=256, top_k=50, n_terms=16, frequent=8, dup_ppm=1000,
               name="keyword-heavy: 5M chunks x 768, 16-term queries (8 from the 1000 most frequent tokens), "
                    "planted duplicate rows, batch 256, top-50"),
}


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: anything a library prints to fd 1 meanwhile (NCCL's version banner
    under NCCL_DEBUG=VERSION is a raw printf) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def measured_tensor_peak():
    """(burst, sustained) dense bf16 TFLOP/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
            return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured"
    except Exception:
        return FALLBACK_BF16_TFLOPS, FALLBACK_BF16_TFLOPS, "fallback"


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
    if args.workload == "c1":
        raise SystemExit("--impl reference --workload c1: the c1 line times the CPU port itself (cpu_baseline); use --workload c1")
    if args.workload != "c2":
        wl = BATCH_WORKLOADS[args.workload]
        sample, steps = 200_000, max(1, args.steps if args.steps != 200 else 20)
        v, dt, n = oracle_batch_rate(wl, sample, steps, threads, 0.0)
        emit({
            "impl": "reference", "metric": f"hybrid recall QPS, {wl['name']}", "value": v, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wl["batch"] / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic", "config": batch_config(args, wl),
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                             "sample": f"{n} queries x {sample} rows x {wl['dim']}, scaled linearly to {wl['rows']} rows; C port of "
                                       f"RecallSearchService.cs (dotnet absent)"},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
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
    emit(line)


def workload_config(args):
    total = args.rows_per_gpu * args.gpus
    return {"workload": f"{total} chunks x {DIM} fp32 row-major in HBM ({args.rows_per_gpu} rows/GPU), single query, "
                        f"hybrid 0.7/0.2/0.1, {N_TERMS} query terms, {TERM_SLOTS} hashed terms/chunk, top-{TOP_K}",
            "rows_total": total, "rows_per_gpu": args.rows_per_gpu, "dim": DIM, "top_k": TOP_K,
            "parallelism": (f"row-sharded x{args.gpus}, per-GPU exact top-{TOP_K}, "
                            + ("fused peer-memory all-gather + merge kernel over NVLink (orr_xchg_allgather_merge); NCCL only "
                               "for set-up and the timing barrier" + ("" if args.no_pipeline else "; the exchange of query i runs on a "
                               "side stream while the rank scans query i+1") if args.exchange == "p2p" else
                               "NCCL all_gather_into_tensor + merge kernel")) if args.gpus > 1 else "1 GPU",
            "l2": "inputs (12.3 GB/GPU) exceed L2 (126 MB); no flush needed",
            "value_units": "queries/s x (rows_total / 1M)"}


def oracle_batch_rate(wl, sample_rows: int, n_queries: int, threads: int, min_seconds: float):
    """The oracle port on the first `sample_rows` rows of the batch workload's corpus: queries/s scaled
    linearly to the workload's row count (the scorer is a per-row loop, one query at a time)."""
    from omni_recall_rag_b200 import synth
    from oracle import oracle_c

    spec = synth.make_spec(wl["dim"], dup_row_ppm=wl["dup_ppm"])
    rows = synth.rows_host(spec, 0, sample_rows)
    blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
    qs = [synth.query_host(spec, qi, wl["rows"], n_terms=wl["n_terms"], frequent_terms=wl["frequent"]) for qi in range(n_queries)]

    def one(q):
        oracle_c.search(emb=rows.emb, dim=wl["dim"], ticks=rows.ticks, content_blob=blob, content_off=off, query=q.text,
                        qvec=q.q, now_ticks=spec.now_ticks, top_k=wl["top_k"], threads=threads)

    one(qs[0])
    t0 = time.perf_counter()
    done = 0
    while True:
        one(qs[done % n_queries])
        done += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds and done >= n_queries:
            break
    return (done / dt) * (sample_rows / float(wl["rows"])), dt, done


def batch_config(args, wl):
    return {"workload": wl["name"], "rows_total": wl["rows"], "rows_per_gpu": wl["rows"], "dim": wl["dim"], "batch": wl["batch"],
            "top_k": wl["top_k"], "n_terms": wl["n_terms"], "parallelism": "1 GPU",
            "split_precision_passes": args.batch_passes,
            "l2": "bf16 planes (15.4 GB) exceed L2 (126 MB); every step is a fresh batch of queries; no flush needed"}


def run_c1(args):
    """--workload c1: 10k x 3072, single query, top-10 — candidate_cap=300 (the reference's behaviour) and all rows."""
    import numpy as np
    import torch

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import _native as N  # noqa: F401
    from omni_recall_rag_b200 import synth

    if args.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--workload c1 is a single-GPU bench")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborr has no CPU path")
    rows_n = 10_000
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    spec = synth.make_spec(DIM)
    shard = orr.RecallShard(DIM, rows_n, device=0, term_slots=TERM_SLOTS)
    shard.fill_synthetic(spec, 0, rows_n)
    n_q = steps + warmup
    queries = [synth.query_host(spec, qi, rows_n, n_terms=N_TERMS) for qi in range(n_q)]
    q_pinned = [torch.from_numpy(q.q).pin_memory() for q in queries]
    res = {}
    with ClockSampler(0) as clocks:
        for cap in (300, 0):
            for i in range(warmup):
                shard.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K, candidate_cap=cap)
            torch.cuda.synchronize()
            dev_ms, wall = [], []
            t0 = time.perf_counter()
            for i in range(warmup, n_q):
                shard.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K, candidate_cap=cap)
                tm = shard.last_timing()
                dev_ms.append(tm["total_device_ms"]); wall.append(tm["wall_ms"])
            e2e_s = time.perf_counter() - t0
            res[cap] = dict(qps_device=steps / (sum(dev_ms) / 1000.0), qps_e2e=steps / e2e_s, device_ms=sum(dev_ms) / steps,
                            call_ms_median=statistics.median(wall), call_ms_p99=sorted(wall)[min(len(wall) - 1, int(0.99 * len(wall)))],
                            path=tm["path"])
    peak, peak_kind = measured_peak()
    bytes_all = rows_n * (4 * DIM + 8 + 4 * TERM_SLOTS)
    line = {
        "metric": "hybrid recall QPS at 10k x 3072 fp32, top-10 (the reference's own CPU-runnable case)", "value": res[300]["qps_device"],
        "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": res[300]["device_ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 scan select + f64 exact re-score", "data": "synthetic",
        "config": {"workload": "InMemory-store shape: 10000 chunks x 3072 fp32, single query, hybrid 0.7/0.2/0.1, 4 query terms, top-10, "
                               "scored over the 300 most recent chunks (candidate_cap=300, the reference's GetRecentChunksAsync cut)",
                   "rows_total": rows_n, "dim": DIM, "top_k": TOP_K, "candidate_cap": 300, "parallelism": "1 GPU",
                   "l2": "the 123 MB corpus fits L2 (126 MB): latency-bound, no flush (steady state of a small store)"},
        "e2e": {"value": res[300]["qps_e2e"], "unit": "queries/s", "h2d_bytes_per_step": 4 * DIM + 12 * N_TERMS,
                "d2h_bytes_per_step": 24 * TOP_K + 8, "call_ms": {"median": res[300]["call_ms_median"], "p99": res[300]["call_ms_p99"]}},
        "all_rows": {"candidate_cap": 0, "value": res[0]["qps_device"], "e2e": res[0]["qps_e2e"], "device_ms": res[0]["device_ms"],
                     "call_ms": {"median": res[0]["call_ms_median"], "p99": res[0]["call_ms_p99"]}},
        "gpu_launches": 2 * steps * 2,
        "roofline": {"bound": "hbm", "kernel": "orr_scan_kernel (all rows, cap=0)", "achieved": bytes_all / (res[0]["device_ms"] / 1000.0) / 1.0e9,
                     "peak": peak, "unit": "GB/s", "frac": bytes_all / (res[0]["device_ms"] / 1000.0) / 1.0e9 / peak, "traffic": None,
                     "note": "a 125 MB pass is launch/latency-bound (and L2-resident), not a bandwidth measurement; the headline roofline is the c2 line"},
        "clocks": clocks.summary(),
    }
    if not args.no_cpu_baseline:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        rows = synth.rows_host(spec, 0, rows_n)
        blob, off = oracle_c.pack_contents(synth.contents_of(rows.term_ids))
        cpu = {}
        for cap in (300, 0):
            for th in (1, threads):
                nq = 0
                t0 = time.perf_counter()
                while True:
                    q = queries[nq % n_q]
                    oracle_c.search(emb=rows.emb, dim=DIM, ticks=rows.ticks, content_blob=blob, content_off=off, query=q.text, qvec=q.q,
                                    now_ticks=spec.now_ticks, top_k=TOP_K, candidate_cap=cap, threads=th)
                    nq += 1
                    dt = time.perf_counter() - t0
                    if dt >= 3.0 and nq >= 8:
                        break
                cpu[(cap, th)] = nq / dt
        line["cpu_baseline"] = {"value": cpu[(300, 1)], "unit": "queries/s", "cores": 1, "kind": "port",
                                "value_all_threads": cpu[(300, threads)], "threads": threads,
                                "all_rows_1_thread": cpu[(0, 1)], "all_rows_all_threads": cpu[(0, threads)],
                                "sample": "the whole workload (10000 rows, no sampling), >= 3 s per figure; 1 thread is the reference's own "
                                          "sequential execution of a request; C port of RecallSearchService.cs:20-119 + "
                                          "InMemoryIngestionStore.cs:57-65 (no dotnet in the image)"}
    emit(line)
    shard.close()


def run_batch(args):
    """--workload c3 | c5: the tcgen05 batched path through orr_search_batch (N=1)."""
    import numpy as np
    import torch

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import _native as N
    from omni_recall_rag_b200 import synth

    wl = BATCH_WORKLOADS[args.workload]
    main_passes = 3 if args.batch_passes == 3 else 1
    if args.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--workload c3/c5 is a single-GPU bench (the batched path shards like the single-query path; not benched)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborr has no CPU path")
    steps = max(1, args.steps if args.steps != 200 else 20)
    warmup = max(3, args.warmup if args.warmup != 20 else 3)
    rows, dim, B, k = wl["rows"], wl["dim"], wl["batch"], wl["top_k"]
    spec = synth.make_spec(dim, dup_row_ppm=wl["dup_ppm"])
    shard = orr.RecallShard(dim, rows, device=0, term_slots=TERM_SLOTS)
    shard.fill_synthetic(spec, 0, rows)
    shard.set_option("batch_passes", args.batch_passes)
    n_b = steps + warmup
    Qs, Ts = [], []
    for i in range(n_b):                          # every step is a fresh batch: new vectors, new terms
        qs = [synth.query_host(spec, i * B + j, rows, n_terms=wl["n_terms"], frequent_terms=wl["frequent"]) for j in range(B)]
        Qs.append(torch.from_numpy(np.stack([q.q for q in qs])).pin_memory())
        Ts.append(orr.BatchTerms.pack([q.terms for q in qs]))

    def one(i):
        hits = shard.search_batch(Qs[i].numpy(), Ts[i], spec.now_ticks, k)
        return hits, shard.last_timing()

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    dev_ms, main_ms, redo, cascaded = [], [], 0, 0
    with ClockSampler(0) as clocks:
        t0 = time.perf_counter()
        for i in range(warmup, n_b):
            hits, tm = one(i)
            assert tm["path"] & 0xff == N.PATH_BATCH, tm
            dev_ms.append(tm["total_device_ms"]); redo += tm["n_survivors"] & 0xffff
            if tm["path"] & N.PATH_ESCALATED:
                cascaded += 1                    # a bf16x3 launch for the unproven queries followed; scan_ms holds both
            else:
                main_ms.append(tm["scan_ms"])
        e2e_s = time.perf_counter() - t0
        t_end = time.time() + 0.6                # a moment more under load for the clock samples
        while time.time() < t_end:
            one(warmup)
    assert all(int(n) == k for n in hits.n_out), "short hit lists"
    # the same batches again: every term bitmap is now cached (a service's steady state on a Zipf vocabulary)
    warm_ms = []
    for i in range(warmup, n_b):
        _, tm = one(i)
        warm_ms.append(tm["total_device_ms"])

    value = steps * B / (sum(dev_ms) / 1000.0)
    burst, sustained, peak_kind = measured_tensor_peak()
    useful = 2.0 * rows * dim * B
    main_avg = sum(main_ms) / max(1, len(main_ms)) if main_ms else float("nan")
    achieved = useful / (main_avg / 1000.0) / 1.0e12
    h2d = B * dim * 4 + B * 4 + (B + 1) * 4 + B * wl["n_terms"] * 8
    line = {
        "metric": f"hybrid recall QPS, {wl['name']}", "value": value, "unit": "queries/s", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": sum(dev_ms) / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("bf16 tcgen05 screen, bf16x3 split precision for unproven queries" if args.batch_passes == 0 else
                  f"bf16x{args.batch_passes} tcgen05 screen") + " (fp32 accumulate in TMEM) + f64 exact re-rank of the candidates",
        "data": "synthetic", "config": batch_config(args, wl),
        "e2e": {"value": steps * B / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * k * 24 + B * 8,
                "ms_per_step": 1000.0 * e2e_s / steps},
        "value_warm_terms": steps * B / (sum(warm_ms) / 1000.0),
        "gpu_launches": 7 * steps,
        "kernels_per_step": ["orr_prep_queries_kernel", "orr_build_rowaux_kernel", "orr_batch_term_bits_kernel (unseen terms only)",
                             f"orr_batch_gemm_kernel<{main_passes},1> (sampling pass)", "orr_batch_threshold_kernel",
                             f"orr_batch_gemm_kernel<{main_passes},0> (main pass)", "orr_batch_finalize_kernel"],
        "roofline": {"bound": "tensor", "kernel": f"orr_batch_gemm_kernel<{main_passes},0> (main pass)", "achieved": achieved,
                     "peak": sustained, "peak_kind": f"{peak_kind} cuBLAS bf16 TFLOP/s, sustained (kernel timed inside a long step); burst {burst}",
                     "unit": "TFLOP/s", "frac": achieved / sustained, "flops_per_launch": useful,
                     "issued_tflops": achieved * main_passes, "issued_frac": achieved * main_passes / sustained,
                     "kernel_ms": main_avg, "traffic": None,
                     "note": "achieved counts the useful 2*N*D*B flops once; the split-precision passes are not credited"},
        "clocks": clocks.summary(),
        "queries_rerun_singly": redo, "steps_with_bf16x3_cascade": cascaded,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        sample = 200_000
        v, dt, n = oracle_batch_rate(wl, sample, 8, threads, 10.0)
        line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"{n} queries x {sample} rows x {dim} (first rows of the same corpus) on {threads} threads "
                                          f"({dt:.1f} s), scaled linearly to {rows} rows; C port of RecallSearchService.cs:20-119"}
    emit(line)
    shard.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how the per-GPU top-k lists meet (fused peer-memory kernel, or NCCL all-gather + merge)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="N>1: every query is a blocking collective step (no exchange/scan overlap between consecutive queries)")
    ap.add_argument("--batch-passes", type=int, default=0, choices=[0, 1, 3],
                    help="0 = auto (bf16 screen, bf16x3 cascade for unproven queries; the library default), 1, 3")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "c1":
        run_c1(args)
        return
    if args.workload != "c2":
        run_batch(args)
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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    n_local = args.rows_per_gpu
    total_rows = n_local * world
    row_base = rank * n_local
    spec = synth.make_spec(DIM)
    shard = orr.RecallShard(DIM, n_local, device=local_rank, term_slots=TERM_SLOTS, row_base=row_base)
    shard.fill_synthetic(spec, row_base, n_local)
    sr = sharded.ShardedRecall(shard, exchange=args.exchange)

    n_q = steps + warmup
    queries = [synth.query_host(spec, qi, total_rows, n_terms=N_TERMS) for qi in range(n_q)]
    q_dev = [torch.from_numpy(q.q).to(dev) for q in queries]
    q_pinned = [torch.from_numpy(q.q).pin_memory() for q in queries]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    # N>1: a stream of queries with the exchange of query i on a side stream (ShardedRecall.search_device_pipelined),
    # so a rank's next scan does not wait for the slowest peer; --no-pipeline keeps every query a blocking collective
    # step.  Both forms are timed; `value` is the pipelined one unless --no-pipeline.
    flags_seen = 0
    main_stream = torch.cuda.current_stream(dev)

    def timed_loop(pipelined: bool):
        for i in range(warmup):
            if pipelined:
                _, _, done = sr.search_device_pipelined(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
            else:
                sr.search_device(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        done = None
        for i in range(warmup, n_q):
            if pipelined:
                _, st, done = sr.search_device_pipelined(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
            else:
                _, st = sr.search_device(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
        if done is not None:
            main_stream.wait_event(done)               # the last exchange (they complete in order) ends the region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        fl = int(st.cpu().numpy()[1])
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), fl

    pipelined = world > 1 and sr.exchange == "p2p" and not args.no_pipeline
    with ClockSampler(local_rank) as clocks:
        dev_ms_sync, fl = timed_loop(False)
        flags_seen |= fl
        dev_ms = dev_ms_sync
        if pipelined:
            dev_ms, fl = timed_loop(True)
            flags_seen |= fl
        # keep sampling under load for ~0.6 s more so short runs still get samples.  The count comes from the
        # all-reduced step time, so every rank issues the SAME number of searches: each one is a collective step
        # (the exchange is sequence-numbered), a time-based loop would let the ranks drift apart.
        extra = max(1, min(2000, int(600.0 / max(dev_ms / steps, 1.0e-3))))
        j = warmup
        for _ in range(extra):
            sr.search_device(q_dev[j], queries[j].terms, spec.now_ticks, TOP_K)
            j = warmup + (j + 1 - warmup) % steps
        torch.cuda.synchronize()
    clock_summary = clocks.summary()

    # ---- end-to-end timing through the host API: `e2e`, and the scan kernel's own duration ----
    for i in range(warmup):
        sr.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K)
    barrier()
    scan_ms, fin_ms, wall_ms, escalated = [], [], [], 0
    t0 = time.perf_counter()
    for i in range(warmup, n_q):
        tq = time.perf_counter()
        hits = sr.search(q_pinned[i].numpy(), queries[i].terms, spec.now_ticks, TOP_K)
        tm = sr.last_timing()
        scan_ms.append(tm["scan_ms"]); fin_ms.append(tm["finalize_ms"])
        wall_ms.append(tm["wall_ms"] if world == 1 else (time.perf_counter() - tq) * 1000.0)
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
    if pipelined:                                      # ... and so does the pipelined form, for a run of queries
        outs = []
        for i in range(n_q - 3, n_q):
            hp_, sp_, done = sr.search_device_pipelined(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
            outs.append((hp_, sp_, done))
        for (hp_, sp_, done), i in zip(outs, range(n_q - 3, n_q)):
            done.synchronize()
            gp, fl = sharded.hits_from_device(hp_, sp_)
            hd, sd = sr.search_device(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
            torch.cuda.synchronize()
            gs, _ = sharded.hits_from_device(hd, sd)
            assert gp.rows.tolist() == gs.rows.tolist() and gp.scores.tolist() == gs.scores.tolist(), "pipelined/blocking exchange disagree"
            flags_seen |= fl

    # per-rank scan kernel time: the exchange makes every query wait for the slowest shard
    scan_avg_local = sum(scan_ms) / len(scan_ms)
    per_rank_scan = [scan_avg_local]
    if world > 1:
        tl = torch.tensor([scan_avg_local], dtype=torch.float64, device=dev)
        tg = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(tg, tl)
        per_rank_scan = [round(float(x), 4) for x in tg.cpu().tolist()]

    scale = total_rows / 1.0e6
    value = steps / (dev_ms / 1000.0) * scale
    e2e_value = steps / e2e_s * scale
    peak, peak_kind = measured_peak()
    bytes_per_launch = n_local * (4 * DIM + 8 + 4 * TERM_SLOTS)
    scan_avg_ms = sum(scan_ms) / len(scan_ms)
    achieved = bytes_per_launch / (scan_avg_ms / 1000.0) / 1.0e9
    launches_per_step = 2 + (1 if world > 1 else 0)
    exchange_kernel = ["orr_xchg_merge_kernel"] if sr.exchange == "p2p" else ["orr_merge_kernel"]

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 scan select + f64 exact re-score", "data": "synthetic",
        "config": workload_config(args),
        "corpus_qps": steps / (dev_ms / 1000.0),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * DIM + 12 * N_TERMS,
                "d2h_bytes_per_step": 24 * TOP_K + 8, "ms_per_step": 1000.0 * e2e_s / steps,
                "call_ms": {"what": "orr_search wall clock" if world == 1 else "ShardedRecall.search wall clock", "median": statistics.median(wall_ms), "p99": sorted(wall_ms)[min(len(wall_ms) - 1, int(0.99 * len(wall_ms)))]}},
        "gpu_launches": launches_per_step * steps,
        "kernels_per_step": ["orr_scan_kernel<24,1>", "orr_rescore_kernel"] + (exchange_kernel if world > 1 else []),
        "exchange": sr.exchange, "per_rank_scan_kernel_ms": per_rank_scan,
        "pipelined_exchange": pipelined,
        "value_blocking_exchange": steps / (dev_ms_sync / 1000.0) * scale,
        "roofline": {"bound": "hbm", "kernel": "orr_scan_kernel<24,1>", "achieved": achieved, "peak": peak,
                     "peak_kind": f"{peak_kind} HBM copy GB/s (MEASURED_PEAKS.json)" if peak_kind == "measured" else "fallback 6650 GB/s",
                     "unit": "GB/s", "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0,
                     "bytes_per_launch": bytes_per_launch, "bytes_per_launch_emb_only": n_local * 4 * DIM,
                     "kernel_ms": scan_avg_ms, "finalize_kernel_ms": sum(fin_ms) / len(fin_ms),
                     "traffic": NCU_SCAN_TRAFFIC_BYTES_PER_ROW * n_local,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/r01_final_kernels.md), scaled by rows"},
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
        emit(line)
    sr.close()
    shard.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
