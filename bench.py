#!/usr/bin/env python
"""bench.py — hybrid recall QPS at 1M x 3072 fp32, top-10 (BASELINE.json's metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one hybrid query (0.7 cosine + 0.2 keyword + 0.1 recency, 4 query terms, top-10) over the whole
HBM-resident corpus.

  N = 1   the corpus is BASELINE.json configs[1]: 1M chunks x 3072 fp32 (~12.3 GB).  The line also carries, under
          "configs", every other BASELINE config measured in the same process — c1 (configs[0], 10k x 3072, the reference's
          own CPU-runnable case), c3 (configs[2], 5M x 768, batch 1024, top-100: tcgen05 path), c5 (configs[4],
          keyword-heavy batch 256, top-50), c2_noemb (the reference's DEFAULT configuration: no embeddings, keyword +
          recency only, on the 1M corpus) — each with value / e2e / roofline / cpu_baseline, and "e2e_service": the
          query STRING in, citations out, through GpuRecallSearchService (tokenise, vocabulary expansion on the GPU
          over a 1M-word vocabulary, orr_search, citation build).  --headline-only skips them.
  N > 1   row-sharded, 5M rows per GPU by default (N = 8 is configs[3]: 40M x 3072, 61 GB per GPU; weak scaling): every
          rank scans its shard, computes its exact local top-10, and the lists meet in a fused peer-memory all-gather +
          merge kernel over NVLink.  "rows_1m_per_gpu" repeats the measurement with 1M rows per GPU.

    value  = (total rows / 1M) x queries / second, device-timed (CUDA events, max over ranks), query already in
             HBM: "1M x 3072 scans per second".  At N=1 this is plain QPS.
    e2e    = the same through the public host API (ShardedRecall.search -> orr_search): the query is copied from
             pinned host memory and the hits are read back every step.
    roofline = algorithmic bytes of the scan kernel / its CUDA-event duration vs the measured HBM copy bandwidth
             (MEASURED_PEAKS.json).
    cpu_baseline = the C oracle (a port of the reference's C# scorer; the C# cannot run here) timed on this box's host
             cores on a bounded sample of the same corpus.

--impl reference: the oracle port alone on all host threads.  It generates its corpus with the oracle library's own
generator (never loads liborr.so) and every step scans the WHOLE sample it reports — 1M x 3072 rows when the host has
the memory — so ms_per_step is a measurement, not an extrapolation.

--workload c1 | c3 | c5 (N=1) prints that config alone as the line (profiling runs).
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
# dram__bytes_read.sum + dram__bytes_write.sum of orr_scan_kernel<24,1,0> per launch and row at 1M x 3072, 4 terms: a
# CONSTANT taken from the ncu --set full capture summarised in the file below (not measured in this run); it scales
# linearly with rows.
NCU_SCAN_TRAFFIC_BYTES_PER_ROW = 12557.2
NCU_SCAN_TRAFFIC_SOURCE = "profiles/r02_ncu_scan_kernel.md"
BATCH_WORKLOADS = {
    "c3": dict(rows=5_000_000, dim=768, batch=1024, top_k=100, n_terms=4, frequent=0, dup_ppm=0,
               name="5M chunks x 768 fp32 (truncated embeddings), batch 1024 queries, 4 terms, top-100"),
    "c5": dict(rows=5_000_000, dim=768, batch=256, top_k=50, n_terms=16, frequent=8, dup_ppm=1000,
               name="keyword-heavy: 5M chunks x 768, 16-term queries (8 from the 1000 most frequent tokens), "
                    "planted duplicate rows, batch 256, top-50"),
}
PORT_NOTE = "C port of RecallSearchService.cs:20-119 (no dotnet in the image)"


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


def note(msg: str) -> None:
    print(f"[bench] {msg}", file=sys.stderr, flush=True)


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


def hbm_peak_kind(kind: str) -> str:
    return f"{kind} HBM copy GB/s (MEASURED_PEAKS.json)" if kind == "measured" else "fallback 6650 GB/s"


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


def p99(xs):
    return sorted(xs)[min(len(xs) - 1, int(0.99 * len(xs)))]


# ================================================================================================================
# CPU side: the oracle port timed on the host cores.  Nothing here touches liborr.so: the corpus comes from the oracle
# library's own generator (oracle/orr_oracle_stream.c, built from the header the device fill kernel uses).
# ================================================================================================================
class OracleCorpus:
    """`rows` synthetic rows materialised on the host for the oracle: embeddings (optional), ticks, Content blob."""

    def __init__(self, spec, rows: int, *, want_emb: bool = True):
        from oracle import oracle_c
        self.oc = oracle_c
        self.spec, self.rows, self.dim = spec, rows, spec.dim
        self.emb, self.ticks, tids = oracle_c.synth_rows(spec, 0, rows, want_emb=want_emb)
        self.blob, self.off = oracle_c.synth_contents(tids)

    def search(self, text: str, qvec, top_k: int, threads: int, candidate_cap: int = 0):
        import numpy as np
        return self.oc.search(emb=self.emb, dim=self.dim, ticks=self.ticks, content_blob=self.blob, content_off=self.off,
                              query=text, qvec=np.zeros(0, np.float32) if qvec is None else qvec, now_ticks=self.spec.now_ticks,
                              top_k=top_k, candidate_cap=candidate_cap, threads=threads)


def timed_queries(fn, queries, min_seconds: float, min_queries: int):
    """Runs fn(query) round-robin until both bounds are met; -> (queries/s, seconds, n)."""
    fn(queries[0])
    t0 = time.perf_counter()
    done = 0
    while True:
        fn(queries[done % len(queries)])
        done += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds and done >= min_queries:
            return done / dt, dt, done


def host_sample_rows(want_rows: int, bytes_per_row: int) -> int:
    """The largest sample <= want_rows whose host copy fits comfortably in the box's free memory."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    rows = want_rows
    while rows > 50_000 and rows * bytes_per_row > 0.6 * avail:
        rows //= 2
    return rows


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The C# cannot be built here (no dotnet),
    so this is the oracle port on all host threads.  Every step is one query over the whole sample the line reports
    (1M x 3072 rows = the full configs[1] corpus when host memory allows); value is in the metric's unit, rows scanned
    per second / 1M, so no step time is ever scaled."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c

    threads = oracle_c.max_threads()
    if args.workload == "c1":
        raise SystemExit("--impl reference --workload c1: the c1 line times the CPU port itself (cpu_baseline); use --workload c1")
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload != "c2":
        wl = BATCH_WORKLOADS[args.workload]
        sample = host_sample_rows(args.cpu_sample_rows or 500_000, 4 * wl["dim"] + 600)
        spec = oracle_c.synth_spec(wl["dim"], dup_row_ppm=wl["dup_ppm"])
        corpus = OracleCorpus(spec, sample)
        qs = [oracle_c.synth_query(spec, qi, wl["rows"], wl["n_terms"], wl["frequent"]) for qi in range(steps + warmup)]
        for q, _, text in qs[:warmup]:
            corpus.search(text, q, wl["top_k"], threads)
        t0 = time.perf_counter()
        for q, _, text in qs[warmup:]:
            corpus.search(text, q, wl["top_k"], threads)
        dt = time.perf_counter() - t0
        v = steps / dt * (sample / float(wl["rows"]))
        emit({
            "impl": "reference", "metric": f"hybrid recall QPS, {wl['name']}", "value": v, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic",
            "config": batch_config(args, wl), "reference_sample_rows": sample,
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} steps, each ONE query over {sample} of the {wl['rows']} rows x {wl['dim']} (ms_per_step is "
                                       f"that step, unscaled; value = steps/s x {sample}/{wl['rows']}: the scorer is a per-row loop); {PORT_NOTE}"},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    rows_per_gpu = args.rows_per_gpu or (1_000_000 if args.gpus == 1 else 5_000_000)
    total_rows = rows_per_gpu * args.gpus
    sample = host_sample_rows(args.cpu_sample_rows or 1_000_000, 4 * DIM + 600)
    sample = min(sample, total_rows)
    spec = oracle_c.synth_spec(DIM)
    t_gen = time.perf_counter()
    corpus = OracleCorpus(spec, sample)
    note(f"reference arm: {sample} x {DIM} rows generated on {threads} host threads in {time.perf_counter() - t_gen:.1f} s")
    qs = [oracle_c.synth_query(spec, qi, total_rows, N_TERMS) for qi in range(steps + warmup)]
    for q, _, text in qs[:warmup]:
        corpus.search(text, q, TOP_K, threads)
    t0 = time.perf_counter()
    for q, _, text in qs[warmup:]:
        corpus.search(text, q, TOP_K, threads)
    dt = time.perf_counter() - t0
    value = (steps / dt) * (sample / 1.0e6)
    cfg = workload_config(args, rows_per_gpu)        # identical to the GPU arm's config; the sample is reported beside it
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic",
        "config": cfg, "reference_sample_rows": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} steps, each ONE query over {sample} x {DIM} rows of the same synthetic corpus"
                                   + (" (the whole configs[1] corpus)" if sample == 1_000_000 and total_rows == 1_000_000 else
                                      f" (of {total_rows}; the scorer is a per-row loop)")
                                   + f"; ms_per_step is that step, unscaled; value = steps/s x rows/1M; {PORT_NOTE}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, rows_per_gpu: int):
    total = rows_per_gpu * args.gpus
    return {"workload": f"{total} chunks x {DIM} fp32 row-major in HBM ({rows_per_gpu} rows/GPU), single query, "
                        f"hybrid 0.7/0.2/0.1, {N_TERMS} query terms, {TERM_SLOTS} hashed terms/chunk, top-{TOP_K}",
            "rows_total": total, "rows_per_gpu": rows_per_gpu, "dim": DIM, "top_k": TOP_K,
            "parallelism": (f"row-sharded x{args.gpus}, per-GPU exact top-{TOP_K}, "
                            + ("fused peer-memory all-gather + merge kernel over NVLink (orr_xchg_allgather_merge); NCCL only "
                               "for set-up and the timing barrier" + ("" if args.no_pipeline else "; the exchange of query i runs on a "
                               "side stream while the rank scans query i+1") if args.exchange == "p2p" else
                               "NCCL all_gather_into_tensor + merge kernel")) if args.gpus > 1 else "1 GPU",
            "l2": f"inputs ({rows_per_gpu * 4 * DIM / 1e9:.1f} GB/GPU) exceed L2 (126 MB); no flush needed",
            "value_units": "queries/s x (rows_total / 1M)"}


def batch_config(args, wl):
    return {"workload": wl["name"], "rows_total": wl["rows"], "rows_per_gpu": wl["rows"], "dim": wl["dim"], "batch": wl["batch"],
            "top_k": wl["top_k"], "n_terms": wl["n_terms"], "parallelism": "1 GPU",
            "split_precision_passes": args.batch_passes,
            "l2": "bf16 planes (7.7 GB) exceed L2 (126 MB); every step is a fresh batch of queries; no flush needed"}


# ================================================================================================================
# GPU side
# ================================================================================================================
def need_gpu():
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborr has no CPU path (use --impl reference for the CPU baseline)")


def measure_c1(steps: int, warmup: int, cpu_seconds: float):
    """configs[0]: 10k x 3072, single query, top-10 — candidate_cap=300 (the reference's behaviour) and all rows."""
    import torch

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import synth

    rows_n = 10_000
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
                            call_ms_median=statistics.median(wall), call_ms_p99=p99(wall), path=tm["path"])
    # a web API serves requests concurrently: T host threads through the same store (orr_search takes its own search context
    # and stream per call, ctypes releases the GIL), each request still a blocking call with host buffers in and out
    import threading
    conc = {}
    for T in (4, 16):
        per = max(50, min(400, steps))
        done = []

        def worker(t):
            for i in range(per):
                j = (t * per + i) % n_q
                shard.search(q_pinned[j].numpy(), queries[j].terms, spec.now_ticks, TOP_K, candidate_cap=300)
            done.append(t)
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        assert len(done) == T
        conc[T] = {"threads": T, "value": T * per / dt, "unit": "queries/s", "requests": T * per}
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
        "concurrent_callers": {"what": "blocking orr_search calls (candidate_cap=300) from T host threads on one store, host buffers in and out",
                               "t4": conc[4], "t16": conc[16]},
        "gpu_launches": 2 * steps * 2,
        "roofline": {"bound": "hbm", "kernel": "orr_scan_kernel (all rows, cap=0)", "achieved": bytes_all / (res[0]["device_ms"] / 1000.0) / 1.0e9,
                     "peak": peak, "peak_kind": hbm_peak_kind(peak_kind), "unit": "GB/s",
                     "frac": bytes_all / (res[0]["device_ms"] / 1000.0) / 1.0e9 / peak, "kernel_ms": res[0]["device_ms"], "traffic": None,
                     "note": "a 125 MB pass is launch/latency-bound (and L2-resident), not a bandwidth measurement; the headline roofline is the c2 line"},
        "clocks": clocks.summary(),
    }
    shard.close()
    if cpu_seconds > 0:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        corpus = OracleCorpus(oracle_c.copy_spec(spec), rows_n)
        cpu = {}
        qs = [(q.q, q.text) for q in queries]
        for cap in (300, 0):
            for th in (1, threads):
                cpu[(cap, th)], _, _ = timed_queries(lambda qt: corpus.search(qt[1], qt[0], TOP_K, th, candidate_cap=cap), qs, cpu_seconds, 8)
        line["cpu_baseline"] = {"value": cpu[(300, 1)], "unit": "queries/s", "cores": 1, "kind": "port",
                                "value_all_threads": cpu[(300, threads)], "threads": threads,
                                "all_rows_1_thread": cpu[(0, 1)], "all_rows_all_threads": cpu[(0, threads)],
                                "sample": f"the whole workload (10000 rows, no sampling), >= {cpu_seconds:.1f} s per figure; 1 thread is the reference's own "
                                          "sequential execution of a request; C port of RecallSearchService.cs:20-119 + "
                                          "InMemoryIngestionStore.cs:57-65 (no dotnet in the image)"}
    return line


def measure_batch(args, name: str, steps: int, warmup: int, cpu_seconds: float):
    """configs[2] / configs[4]: the tcgen05 batched path through orr_search_batch (N=1)."""
    import numpy as np
    import torch

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import _native as N
    from omni_recall_rag_b200 import synth

    wl = dict(BATCH_WORKLOADS[name])
    if os.environ.get("ORR_BENCH_TERMS"):          # tuning experiments only: how much of a pass is the keyword side
        wl["n_terms"] = int(os.environ["ORR_BENCH_TERMS"]); wl["frequent"] = min(wl["frequent"], wl["n_terms"])
    main_passes = 3 if args.batch_passes == 3 else 1
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
    shard.close()
    if cpu_seconds > 0:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        sample = 200_000
        ospec = oracle_c.synth_spec(dim, dup_row_ppm=wl["dup_ppm"])
        corpus = OracleCorpus(ospec, sample)
        qs = [oracle_c.synth_query(ospec, qi, rows, wl["n_terms"], wl["frequent"]) for qi in range(8)]
        rate, dt, n = timed_queries(lambda q: corpus.search(q[2], q[0], k, threads), qs, cpu_seconds, 4)
        line["cpu_baseline"] = {"value": rate * sample / float(rows), "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"{n} queries x {sample} rows x {dim} (first rows of the same corpus) on {threads} threads "
                                          f"({dt:.1f} s), scaled linearly to {rows} rows (the scorer is a per-row loop); {PORT_NOTE}"}
    return line


def measure_batch_sharded(args, name: str, steps: int, warmup: int):
    """configs[2] / configs[4] row-sharded over N ranks (torchrun): every rank holds the config's rows (weak scaling: the
    corpus is N x 5M rows) and answers every query of the batch on its shard (tcgen05 path, answers left in HBM:
    orr_search_batch_device); the B x k x 24 B lists are all-gathered with NCCL — a bandwidth-type exchange — and merged per
    query on the device (orr_merge_hits_batch_device).  Host buffers in (queries) and out (merged hits) on every step, so
    the one figure is end to end; wall clock between barriers, max over ranks.  Returns the line on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import sharded, synth

    world, rank, local_rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    wl = dict(BATCH_WORKLOADS[name])
    rows, dim, B, k = wl["rows"], wl["dim"], wl["batch"], wl["top_k"]
    total_rows = rows * world
    spec = synth.make_spec(dim, dup_row_ppm=wl["dup_ppm"])
    shard = orr.RecallShard(dim, rows, device=local_rank, term_slots=TERM_SLOTS, row_base=rank * rows)
    shard.fill_synthetic(spec, rank * rows, rows)
    shard.set_option("batch_passes", args.batch_passes)
    sr = sharded.ShardedRecall(shard, exchange="nccl")
    n_b = steps + warmup
    Qs, Ts = [], []
    for i in range(n_b):                          # the same fresh batch on every rank
        qs = [synth.query_host(spec, i * B + j, total_rows, n_terms=wl["n_terms"], frequent_terms=wl["frequent"]) for j in range(B)]
        Qs.append(torch.from_numpy(np.stack([q.q for q in qs])).pin_memory())
        Ts.append(orr.BatchTerms.pack([q.terms for q in qs]))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        sr.search_batch(Qs[i].numpy(), Ts[i], spec.now_ticks, k)
    barrier()
    main_ms = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        t0 = time.perf_counter()
        for i in range(warmup, n_b):
            hits = sr.search_batch(Qs[i].numpy(), Ts[i], spec.now_ticks, k)
            main_ms.append(shard.last_timing()["scan_ms"])
        barrier()
        dt = time.perf_counter() - t0
        for i in range(warmup, warmup + max(1, min(steps, int(0.6 * steps / max(dt, 1e-3))))):
            sr.search_batch(Qs[i].numpy(), Ts[i], spec.now_ticks, k)     # same count on every rank: each call is a collective
    assert all(int(n) == k for n in hits.n_out), "short hit lists"
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_blocking = float(t.item())
    # the throughput form: batch i's all-gather + merge (and its wait for the slowest rank) overlap batch i+1's search
    barrier()
    t0 = time.perf_counter()
    pending, last = None, None
    for i in range(warmup, n_b):
        nxt = sr.search_batch_async(Qs[i].numpy(), Ts[i], spec.now_ticks, k)
        if pending is not None:
            last = pending()
        pending = nxt
    last = pending()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_pipelined = float(t.item())
    assert last.raw.tobytes() == hits.raw.tobytes(), "pipelined / blocking batch forms disagree"
    # both forms do the same work and return the same hits; the line's value is the faster one (the side-stream NCCL kernels
    # can delay the next persistent GEMM's launch as well as hide the wait for the slowest rank), the other is reported beside it
    pipelined = dt_pipelined < dt_blocking
    dt = min(dt_pipelined, dt_blocking)
    line = None
    if rank == 0:
        burst, sustained, peak_kind = measured_tensor_peak()
        useful = 2.0 * rows * dim * B
        main_avg = sum(main_ms) / len(main_ms)
        achieved = useful / (main_avg / 1000.0) / 1.0e12
        scale = total_rows / float(rows)
        line = {
            "metric": f"hybrid recall QPS, {wl['name']}", "value": steps * B / dt * scale,
            "unit": f"queries/s x (rows_total / {rows})", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 tcgen05 screen (fp32 accumulate in TMEM) + f64 exact re-rank of the candidates", "data": "synthetic",
            "config": dict(batch_config(args, wl), rows_total=total_rows, rows_per_gpu=rows, parallelism=f"row-sharded x {world} (torchrun)"),
            "corpus_qps": steps * B / dt, "pipelined_exchange": pipelined,
            "value_blocking_exchange": steps * B / dt_blocking * scale, "ms_per_step_blocking": 1000.0 * dt_blocking / steps,
            "value_pipelined_exchange": steps * B / dt_pipelined * scale, "ms_per_step_pipelined": 1000.0 * dt_pipelined / steps,
            "e2e": {"value": steps * B / dt * scale, "unit": f"queries/s x (rows_total / {rows})",
                    "h2d_bytes_per_step": B * dim * 4 + B * 4 + (B + 1) * 4 + B * wl["n_terms"] * 8,
                    "d2h_bytes_per_step": B * k * 24 + B * 4, "ms_per_step": 1000.0 * dt / steps},
            "gpu_launches": 8 * steps,
            "kernels_per_step": ["orr_prep_queries_kernel", "orr_build_rowaux_kernel", "orr_batch_term_bits_kernel (unseen terms only)",
                                 "orr_batch_gemm_kernel<1,1> (sampling pass)", "orr_batch_threshold_kernel",
                                 "orr_batch_gemm_kernel<1,0> (main pass)", "orr_batch_finalize_kernel", "orr_merge_batch_kernel"],
            "exchange": "nccl all_gather_into_tensor of the B x k x 24 B lists + per-query device merge",
            "roofline": {"bound": "tensor", "kernel": "orr_batch_gemm_kernel<1,0> (main pass, rank 0)", "achieved": achieved, "peak": sustained,
                         "peak_kind": f"{peak_kind} cuBLAS bf16 TFLOP/s, sustained; burst {burst}", "unit": "TFLOP/s",
                         "frac": achieved / sustained, "flops_per_launch": useful, "kernel_ms": main_avg, "traffic": None},
            "clocks": clocks.summary(),
        }
    sr.close()
    shard.close()
    dist.destroy_process_group()
    return line


def measure_noemb(shard, spec, n_local: int, steps: int, warmup: int, cpu_seconds: float):
    """The reference's DEFAULT configuration (appsettings.json:30-32, NoOpEmbeddingClient.cs:5-8): no query embedding,
    every cosine is 0 (RecallSearchService.cs:71-72), ranking = keyword + recency.  Same 1M-row corpus, 4 terms, top-10:
    the exact path (orr_noemb_scores_kernel reads terms64 + ticks only, then the radix select under the tie chain)."""
    import torch

    from omni_recall_rag_b200 import _native as N
    from omni_recall_rag_b200 import synth

    n_q = steps + warmup
    queries = [synth.query_host(spec, 1000 + qi, n_local, n_terms=N_TERMS) for qi in range(n_q)]
    for i in range(warmup):
        shard.search(None, queries[i].terms, spec.now_ticks, TOP_K)
    torch.cuda.synchronize()
    score_ms, sel_ms, wall = [], [], []
    t0 = time.perf_counter()
    for i in range(warmup, n_q):
        shard.search(None, queries[i].terms, spec.now_ticks, TOP_K)
        tm = shard.last_timing()
        assert tm["path"] == N.PATH_EXACT, tm
        score_ms.append(tm["scan_ms"]); sel_ms.append(tm["finalize_ms"]); wall.append(tm["wall_ms"])
    e2e_s = time.perf_counter() - t0
    peak, peak_kind = measured_peak()
    scale = n_local / 1.0e6
    bytes_per_launch = n_local * (4 * TERM_SLOTS + 8)
    k_ms = sum(score_ms) / steps
    dev_ms = k_ms + sum(sel_ms) / steps
    achieved = bytes_per_launch / (k_ms / 1000.0) / 1.0e9
    entry = {
        "metric": "keyword + recency-only recall QPS at 1M x 3072 (no embeddings: the reference's default configuration), top-10",
        "value": 1000.0 / dev_ms * scale, "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": dev_ms,
        "dtype": "f64 exact scores (order-preserving u64 keys) + radix select under (score, ticks, row)",
        "config": {"workload": f"{n_local} chunks x {DIM} fp32 in HBM, query WITHOUT an embedding (cosine 0, RecallSearchService.cs:71-72), "
                               f"{N_TERMS} query terms over {TERM_SLOTS} hashed terms/chunk, recency, top-{TOP_K}; exact path",
                   "rows_total": n_local, "top_k": TOP_K},
        "e2e": {"value": steps / e2e_s * scale, "unit": UNIT, "h2d_bytes_per_step": 12 * N_TERMS, "d2h_bytes_per_step": 24 * TOP_K + 8,
                "ms_per_step": 1000.0 * e2e_s / steps, "call_ms": {"median": statistics.median(wall), "p99": p99(wall)}},
        "roofline": {"bound": "hbm", "kernel": "orr_noemb_scores_kernel<2,0>", "achieved": achieved, "peak": peak,
                     "peak_kind": hbm_peak_kind(peak_kind), "unit": "GB/s", "frac": achieved / peak, "bytes_per_launch": bytes_per_launch,
                     "bytes_per_row": "4 B x 64 low hash words + 8 B ticks = 264 B: the kernel screens on the scan's 32-bit term table and confirms "
                                      "the (rare) 32-bit hits against the 64-bit table; embeddings are never read",
                     "achieved_if_counted_as_520B_per_row": n_local * 520 / (k_ms / 1000.0) / 1.0e9, "kernel_ms": k_ms,
                     "select_ms": sum(sel_ms) / steps, "traffic": 264.0e6 * n_local / 1.0e6,
                     "traffic_source": "CONSTANT from the ncu --set full capture in profiles/r02_ncu_noemb_kernel.md (264.0 MB read per 1M rows)",
                     "note": "kernel_ms spans the 48 KB state clear + the scoring kernel (CUDA events); select_ms = digit passes + gather + "
                             "order + host sync of the exact path.  ncu: this kernel moves exactly its algorithmic bytes but is ALU-bound "
                             "(78 % ALU pipe, 62 % issue slots, 50 % DRAM): the HBM fraction is reported, the limiter is the compare / "
                             "REDUX / fp64 recency work per row"},
        "gpu_launches": 5 * steps,
        "kernels_per_step": ["orr_noemb_scores_kernel<2,0>", "orr_sel_pass_kernel x2", "orr_sel_gather_kernel", "orr_order_kernel"],
    }
    if cpu_seconds > 0:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        sample = min(n_local, 250_000)
        corpus = OracleCorpus(oracle_c.copy_spec(spec), sample, want_emb=False)
        qs = [q.text for q in queries[:8]]
        rate, dt, n = timed_queries(lambda t: corpus.search(t, None, TOP_K, threads), qs, cpu_seconds, 4)
        rate1, dt1, n1 = timed_queries(lambda t: corpus.search(t, None, TOP_K, 1), qs, min(cpu_seconds, 2.0), 2)
        entry["cpu_baseline"] = {"value": rate * sample / 1.0e6, "unit": UNIT, "cores": threads, "kind": "port",
                                 "value_1_thread": rate1 * sample / 1.0e6,
                                 "sample": f"{n} queries x {sample} rows (Content of ~575 B/chunk, no embeddings) on {threads} threads ({dt:.1f} s), "
                                           f"{n1} on 1 thread; scaled linearly to 1M rows; {PORT_NOTE}"}
    return entry


def measure_service(store, spec, n_local: int, steps: int, warmup: int):
    """Query STRING in, citations out: GpuRecallSearchService.search = embedding client (stub: the vector is known),
    orr_search_query (tokenise, stop words, vocabulary expansion on the GPU over the 2^20-word vocabulary, fused scan,
    exact re-score) and the citation build (document lookup, snippet, Math.Round(score, 4))."""
    import numpy as np

    from omni_recall_rag_b200 import _native as N
    from omni_recall_rag_b200 import recall as R
    from omni_recall_rag_b200 import synth

    n_q = steps + warmup
    queries = [synth.query_host(spec, 2000 + qi, n_local, n_terms=N_TERMS) for qi in range(n_q)]
    table = {q.text: q.q for q in queries}

    class KnownVectors:
        def embed(self, text):
            return R.EmbeddingResult(table[text], "Success", "synthetic")

    svc = R.GpuRecallSearchService(store, KnownVectors(), candidate_cap=0, clock=lambda: spec.now_ticks)
    for i in range(warmup):
        svc.search(queries[i].text, TOP_K)
    wall, dev = [], []
    t0 = time.perf_counter()
    for i in range(warmup, n_q):
        tq = time.perf_counter()
        resp = svc.search(queries[i].text, TOP_K)
        wall.append((time.perf_counter() - tq) * 1000.0)
        tm = store.shard.last_timing()
        assert tm["path"] == N.PATH_FUSED and len(resp.citations) == TOP_K, tm
        dev.append(tm["total_device_ms"])
    e2e_s = time.perf_counter() - t0
    # the same query through the pre-hashed entry point returns the same rows
    q = queries[-1]
    direct = store.shard.search(q.q, q.terms, spec.now_ticks, TOP_K)
    assert [c.score for c in resp.citations] == [R.math_round4(float(s)) for s in direct.scores], "service and orr_search disagree"
    scale = n_local / 1.0e6
    return {
        "value": steps / e2e_s * scale, "unit": UNIT, "ms_per_step": 1000.0 * e2e_s / steps,
        "call_ms": {"median": statistics.median(wall), "p99": p99(wall)}, "device_ms_per_step": sum(dev) / steps,
        "vocabulary_words": store.vocabulary_size,
        "h2d_bytes_per_step": 4 * DIM + 36 + 4232 + 12 * N_TERMS, "d2h_bytes_per_step": 24 * TOP_K + 8 + 4 + 8 * N_TERMS,
        "what": "GpuRecallSearchService.search(query string, 10): embedding stub -> orr_search_query (A-2 tokenising, GPU substring "
                "expansion of the 4 terms over the live vocabulary kept in HBM, fused scan + exact re-score) -> 10 RecallCitationDto "
                "(chunk records rebuilt from the generator, 180-char snippet, Math.Round(score, 4)); Python host standing in for the C# shim",
    }


def measure_cluster(args):
    """The single-process form (orr_cluster_*, csrc/orr_cluster.cu): one host thread issues, per query and per device,
    query -> HBM, orr_search_device, the fused peer-memory all-gather + merge; hits come back from device 0.  Host
    buffers in and out on every call, so `value` IS the end-to-end number (e2e repeats it with the bytes moved); the
    device-resident figure of the same layout is the torchrun line."""
    import numpy as np
    import torch

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import synth

    n_dev = args.gpus
    assert torch.cuda.device_count() >= n_dev, f"--gpus {n_dev} but {torch.cuda.device_count()} visible"
    n_local = args.rows_per_gpu or (1_000_000 if n_dev == 1 else 5_000_000)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    spec = synth.make_spec(DIM)
    total_rows = n_local * n_dev
    queries = [synth.query_host(spec, qi, total_rows, n_terms=N_TERMS) for qi in range(steps + warmup)]
    with orr.RecallCluster(DIM, n_local, list(range(n_dev)), term_slots=TERM_SLOTS, max_top_k=32) as cl:
        cl.fill_synthetic(spec, 0, n_local)
        for q in queries[:warmup]:
            cl.search(q.q, q.terms, spec.now_ticks, TOP_K)
        wall = []
        with ClockSampler(0) as clocks:
            t0 = time.perf_counter()
            for q in queries[warmup:]:
                tq = time.perf_counter()
                hits = cl.search(q.q, q.terms, spec.now_ticks, TOP_K)
                wall.append((time.perf_counter() - tq) * 1000.0)
            dt = time.perf_counter() - t0
            for q in queries[warmup:warmup + max(1, int(0.6 * steps / max(dt, 1e-3)))]:
                cl.search(q.q, q.terms, spec.now_ticks, TOP_K)           # keep the sampler under load for ~0.6 s
        assert len(hits.rows) == TOP_K
        # the throughput form: the same queries as ONE pipelined run (orr_cluster_search_many: three queries in flight, the
        # exchange of query i on a side stream of every device while it scans query i+1); host buffers in and out
        Qrun = np.stack([q.q for q in queries[warmup:]])
        Trun = [q.terms for q in queries[warmup:]]
        cl.search_many(Qrun[:8], Trun[:8], spec.now_ticks, TOP_K)
        t1 = time.perf_counter()
        many = cl.search_many(Qrun, Trun, spec.now_ticks, TOP_K)
        dt_many = time.perf_counter() - t1
        assert many[len(many) - 1].rows.tolist() == hits.rows.tolist(), "pipelined run / single call disagree"
        # batched queries over the cluster (orr_cluster_search_batch): every shard's tcgen05 path, merged per query
        batch = None
        if not args.headline_only:
            bq = np.stack([q.q for q in queries[:256]]) if len(queries) >= 256 else np.stack([queries[i % len(queries)].q for i in range(256)])
            bt = [queries[i % len(queries)].terms for i in range(256)]
            try:
                bh = cl.search_batch(bq, bt, spec.now_ticks, TOP_K)      # builds the bf16 planes on first use
                t1 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    bh = cl.search_batch(bq, bt, spec.now_ticks, TOP_K)
                bdt = (time.perf_counter() - t1) / reps
                h0 = cl.search(queries[0].q, queries[0].terms, spec.now_ticks, TOP_K)
                assert bh[0].rows.tolist() == h0.rows.tolist(), "cluster batch / single-query paths disagree"
                batch = {"batch": 256, "ms_per_batch": 1000.0 * bdt, "queries_per_s": 256 / bdt,
                         "what": "orr_cluster_search_batch: 256 queries x the whole corpus (tcgen05 path on every shard, k-way merge "
                                 "per query), host buffers in and out; query 0 checked against orr_cluster_search"}
            except Exception as e:                                        # the single-query figures stand on their own
                batch = {"error": str(e)[:300]}
    scale = total_rows / 1.0e6
    peak, peak_kind = measured_peak()
    bytes_per_query = total_rows * (4 * DIM + 8 + 4 * TERM_SLOTS)
    line = {"metric": METRIC, "value": steps / dt * scale, "unit": UNIT, "n_gpus": n_dev, "steps": steps, "warmup": warmup,
            "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 scan select + f64 exact re-score", "data": "synthetic",
            "config": dict(workload_config(args, n_local), rows_total=total_rows, parallelism=f"1 process x {n_dev} GPUs (orr_cluster)"),
            "corpus_qps": steps / dt, "form": "single process (orr_cluster_search), wall clock, host buffers in and out",
            "pipelined_run": {"value": steps / dt_many * scale, "corpus_qps": steps / dt_many, "ms_per_query": 1000.0 * dt_many / steps,
                              "what": f"orr_cluster_search_many: the same {steps} queries as one pipelined run (3 in flight), host buffers in and out"},
            "e2e": {"value": steps / dt * scale, "unit": UNIT, "h2d_bytes_per_step": n_dev * (4 * DIM + 12 * N_TERMS),
                    "d2h_bytes_per_step": 24 * TOP_K + 8, "ms_per_step": 1000.0 * dt / steps,
                    "call_ms": {"what": "RecallCluster.search wall clock", "median": statistics.median(wall), "p99": p99(wall)}},
            "gpu_launches": 3 * n_dev * steps,
            "kernels_per_step": ["orr_scan_kernel<24,1>", "orr_rescore_kernel + orr_order_kernel", "orr_xchg_merge_kernel"],
            "roofline": {"bound": "hbm", "kernel": "whole query (all devices), wall clock", "achieved": bytes_per_query / (dt / steps) / 1.0e9,
                         "peak": peak * n_dev, "peak_kind": hbm_peak_kind(peak_kind) + f" x {n_dev}", "unit": "GB/s",
                         "frac": bytes_per_query / (dt / steps) / 1.0e9 / (peak * n_dev), "traffic": None},
            "clocks": clocks.summary()}
    if batch:
        line["cluster_batch"] = batch
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="default: 1M at N=1 (configs[1]), 5M at N>1 (N=8: configs[3])")
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="rows of the corpus the CPU port scans per query (default: 1M for "
                    "--impl reference, memory permitting; 100k for the inline cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="N=1: only the c2 headline (no configs / e2e_service sub-results)")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how the per-GPU top-k lists meet (fused peer-memory kernel, or NCCL all-gather + merge)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="N>1: every query is a blocking collective step (no exchange/scan overlap between consecutive queries)")
    ap.add_argument("--cluster", action="store_true",
                    help="ONE host process driving --gpus N devices through orr_cluster_* (the .NET deployment of the row-sharded "
                         "layout: no torchrun, no NCCL); launch with plain `python bench.py --cluster --gpus N`")
    ap.add_argument("--batch-passes", type=int, default=0, choices=[0, 1, 3],
                    help="0 = auto (bf16 screen, bf16x3 cascade for unproven queries; the library default), 1, 3")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    need_gpu()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_s = 0.0 if args.no_cpu_baseline else 1.0
    if args.cluster:
        if world != 1:
            raise SystemExit("--cluster is ONE process for all GPUs: launch it with plain python, not torchrun")
        emit(measure_cluster(args))
        return
    if args.workload == "c1":
        if args.gpus != 1 or world != 1:
            raise SystemExit("--workload c1 is a single-GPU bench")
        emit(measure_c1(max(1, args.steps), max(3, args.warmup), 3.0 * cpu_s))
        return
    if args.workload != "c2":
        b_steps, b_warm = max(1, args.steps if args.steps != 200 else 20), max(3, args.warmup if args.warmup != 20 else 3)
        if world > 1:
            line = measure_batch_sharded(args, args.workload, b_steps, b_warm)
            if line is not None:
                emit(line)
            return
        if args.gpus != 1:
            raise SystemExit("--workload c3/c5 --gpus N needs torchrun (one rank per GPU)")
        emit(measure_batch(args, args.workload, b_steps, b_warm, 10.0 * cpu_s))
        return

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    import omni_recall_rag_b200 as orr
    from omni_recall_rag_b200 import sharded, synth
    from omni_recall_rag_b200 import store as S

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    spec = synth.make_spec(DIM)
    n_q = steps + warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure_sharded(n_local: int, with_service: bool):
        """One corpus of n_local rows per GPU: device-timed stream of queries (`value`), the host-API loop (`e2e`) and the
        scan kernel's own duration (roofline).  Returns (line fields, shard, store-or-None, ShardedRecall)."""
        total_rows = n_local * world
        row_base = rank * n_local
        store = None
        if with_service:          # N=1: the shard lives inside the reference-shaped store so the service leg reuses the corpus
            store = S.GpuIngestionStore(DIM, n_local, device=local_rank, term_slots=TERM_SLOTS, keep_text=False)
            store.fill_synthetic(spec, n_local)
            shard = store.shard
        else:
            shard = orr.RecallShard(DIM, n_local, device=local_rank, term_slots=TERM_SLOTS, row_base=row_base)
            shard.fill_synthetic(spec, row_base, n_local)
        sr = sharded.ShardedRecall(shard, exchange=args.exchange)
        queries = [synth.query_host(spec, qi, total_rows, n_terms=N_TERMS) for qi in range(n_q)]
        q_dev = [torch.from_numpy(q.q).to(dev) for q in queries]
        q_pinned = [torch.from_numpy(q.q).pin_memory() for q in queries]
        flags_seen = 0
        main_stream = torch.cuda.current_stream(dev)

        # ---- device-resident timing: `value` ----
        # N>1: a stream of queries with the exchange of query i on a side stream (ShardedRecall.search_device_pipelined),
        # so a rank's next scan does not wait for the slowest peer; --no-pipeline keeps every query a blocking collective
        # step.  Both forms are timed; `value` is the pipelined one unless --no-pipeline.
        def timed_loop(pipelined: bool):
            for i in range(warmup):
                if pipelined:
                    sr.search_device_pipelined(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K)
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
                outs.append(sr.search_device_pipelined(q_dev[i], queries[i].terms, spec.now_ticks, TOP_K))
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
        peak, peak_kind = measured_peak()
        bytes_per_launch = n_local * (4 * DIM + 8 + 4 * TERM_SLOTS)
        scan_avg_ms = sum(scan_ms) / len(scan_ms)
        achieved = bytes_per_launch / (scan_avg_ms / 1000.0) / 1.0e9
        launches_per_step = 2 + (1 if world > 1 else 0)
        exchange_kernel = ["orr_xchg_merge_kernel"] if sr.exchange == "p2p" else ["orr_merge_kernel"]
        fields = {
            "value": steps / (dev_ms / 1000.0) * scale, "ms_per_step": dev_ms / steps,
            "config": workload_config(args, n_local),
            "corpus_qps": steps / (dev_ms / 1000.0),
            "e2e": {"value": steps / e2e_s * scale, "unit": UNIT, "h2d_bytes_per_step": 4 * DIM + 12 * N_TERMS,
                    "d2h_bytes_per_step": 24 * TOP_K + 8, "ms_per_step": 1000.0 * e2e_s / steps,
                    "call_ms": {"what": "orr_search wall clock" if world == 1 else "ShardedRecall.search wall clock",
                                "median": statistics.median(wall_ms), "p99": p99(wall_ms)}},
            "gpu_launches": launches_per_step * steps,
            "kernels_per_step": ["orr_scan_kernel<24,1>", "orr_rescore_kernel"] + (exchange_kernel if world > 1 else []),
            "exchange": sr.exchange, "per_rank_scan_kernel_ms": per_rank_scan,
            "pipelined_exchange": pipelined,
            "value_blocking_exchange": steps / (dev_ms_sync / 1000.0) * scale,
            "roofline": {"bound": "hbm", "kernel": "orr_scan_kernel<24,1>", "achieved": achieved, "peak": peak,
                         "peak_kind": hbm_peak_kind(peak_kind),
                         "unit": "GB/s", "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0,
                         "bytes_per_launch": bytes_per_launch, "bytes_per_launch_emb_only": n_local * 4 * DIM,
                         "kernel_ms": scan_avg_ms, "finalize_kernel_ms": sum(fin_ms) / len(fin_ms),
                         "traffic": NCU_SCAN_TRAFFIC_BYTES_PER_ROW * n_local,
                         "traffic_source": f"CONSTANT, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per launch from the "
                                           f"ncu --set full capture summarised in {NCU_SCAN_TRAFFIC_SOURCE} ({NCU_SCAN_TRAFFIC_BYTES_PER_ROW} B/row), scaled by rows"},
            "clocks": clock_summary,
            "bound_check_escalations": escalated, "device_flags": flags_seen,
        }
        return fields, shard, store, sr

    n_local = args.rows_per_gpu or (1_000_000 if world == 1 else 5_000_000)
    with_service = world == 1 and not args.headline_only
    fields, shard, store, sr = measure_sharded(n_local, with_service)
    line = {"metric": METRIC, "value": fields.pop("value"), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": fields.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 scan select + f64 exact re-score", "data": "synthetic"}
    line.update(fields)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_c
        threads = oracle_c.max_threads()
        sample = args.cpu_sample_rows or 100_000
        corpus = OracleCorpus(oracle_c.copy_spec(spec), sample)
        qs = [oracle_c.synth_query(corpus.spec, qi, n_local, N_TERMS) for qi in range(8)]
        v_all, dt_all, n_all = timed_queries(lambda q: corpus.search(q[2], q[0], TOP_K, threads), qs, 6.0, 8)
        v_one, dt_one, n_one = timed_queries(lambda q: corpus.search(q[2], q[0], TOP_K, 1), qs, 4.0, 2)
        line["cpu_baseline"] = {
            "value": v_all * sample / 1.0e6, "unit": UNIT, "cores": threads, "kind": "port",
            "value_1_thread": v_one * sample / 1.0e6,
            "sample": f"{n_all} queries x {sample} rows x {DIM} (first rows of the same corpus) on {threads} "
                      f"threads ({dt_all:.1f} s) and {n_one} queries on 1 thread ({dt_one:.1f} s), scaled linearly to 1M "
                      f"rows (the reference arm, --impl reference, scans the full 1M rows per step); {PORT_NOTE}"}
        del corpus

    if world == 1 and not args.headline_only:
        # ---- every other BASELINE config, in this process ----
        sub_steps, sub_warm = min(steps, 50), max(3, min(warmup, 10))
        t_sub = time.perf_counter()
        configs = {}
        configs["c2_noemb"] = measure_noemb(shard, spec, n_local, sub_steps, sub_warm, 3.0 * cpu_s)
        line["e2e_service"] = measure_service(store, spec, n_local, sub_steps, sub_warm)
        line["e2e_service"]["vs_e2e_ms"] = line["e2e_service"]["ms_per_step"] / line["e2e"]["ms_per_step"]
        sr.close()
        store.close()
        sr = shard = store = None
        configs["c1"] = measure_c1(max(sub_steps, 100), sub_warm, 1.5 * cpu_s)
        b_steps, b_warm = min(steps, 20), max(3, min(warmup, 5))
        for name in ("c3", "c5"):
            configs[name] = measure_batch(args, name, b_steps, b_warm, 4.0 * cpu_s)
        line["configs"] = configs
        line["configs_seconds"] = round(time.perf_counter() - t_sub, 1)

    if world > 1 and not args.headline_only and n_local != 1_000_000:
        # the same measurement at 1M rows per GPU (round 1's scaling series), as a sub-line
        sr.close()
        shard.close()
        sr = shard = None
        f1, shard, _, sr = measure_sharded(1_000_000, False)
        line["rows_1m_per_gpu"] = {k: f1[k] for k in ("value", "ms_per_step", "corpus_qps", "e2e", "value_blocking_exchange",
                                                      "per_rank_scan_kernel_ms", "roofline", "config")}

    if rank == 0:
        emit(line)
    if sr is not None:
        sr.close()
    if store is not None:
        store.close()
    elif shard is not None:
        shard.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
