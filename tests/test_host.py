"""Host-side logic and the C-ABI surface, without any GPU compute."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import omni_recall_rag_b200 as orr
from omni_recall_rag_b200 import _native as N
from omni_recall_rag_b200 import recall as R
from omni_recall_rag_b200 import sharded, synth
from omni_recall_rag_b200.store import _distinct_lower_tokens
from oracle import oracle_c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DAY = 864_000_000_000


def _header_functions():
    src = open(os.path.join(ROOT, "include", "orr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 20
    lib = N.lib()
    for n in names:
        assert hasattr(lib, n), f"liborr.so does not export {n}"
    assert sorted(names) == N.declared_symbols(), "ctypes table and include/orr.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", N.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (orr_\w+)", out))
    assert set(names) <= exported


def test_every_declared_entry_point_has_a_ctypes_signature():
    """include/orr.h is the boundary; the Python host (and the tests through it) may only call what it declares, with
    argument types spelled out — no entry point reaches ctypes with default int marshalling."""
    names = _header_functions()
    missing = [n for n in names if n not in N._SIGNATURES]
    extra = [n for n in N._SIGNATURES if n not in names]
    assert not missing and not extra, (missing, extra)


def test_library_is_sm100a_only_with_tma_bulk_copies():
    """The scan kernel must be real Blackwell code: UBLKCP (cp.async.bulk) in sm_100a SASS."""
    r = subprocess.run(["cuobjdump", "-lelf", N.library_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    sass = subprocess.run(["cuobjdump", "-sass", N.library_path()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass and "LDS.128" in sass


def test_no_cpu_fallback_store_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(N.OrrError) as e:
        orr.RecallShard(64, 16)
    assert e.value.code == N.ORR_E_CUDA


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "omni_recall_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.lower().replace("# oracle", "").replace("(oracle input)", "").replace("for the oracle", "").replace("the cpu oracle", ""), f


QUERIES = ["azure", "what is the kubernetes", "What backend did we choose?", "the of and", "  \t ",
           "Azure AZURE azure cosmos", "a b c", "ÀÉÎ Straße ΣΟΦΙΑ Привет", "x" * 300, "tab\tsep\nnl"]


@pytest.mark.parametrize("q", QUERIES)
def test_tokenize_query_matches_oracle_terms(q):
    got = orr.tokenize_query(q)
    exp = [orr.hash_term(t) for t in oracle_c.query_terms(q)]
    assert list(got) == exp


@settings(max_examples=100, deadline=None)
@given(st.text(alphabet=st.sampled_from(list("abAB zZéÉ?\t\n  яЯσΣ")), max_size=40))
def test_tokenizers_agree_on_arbitrary_text(text):
    assert list(orr.tokenize_query(text)) == [orr.hash_term(t) for t in oracle_c.query_terms(text)]
    assert list(orr.tokenize_content(text)) == [orr.hash_term(t) for t in _distinct_lower_tokens(text)]


def _py_lower_invariant(ch: str) -> str:
    lo = ch.lower()                      # full mapping; == the simple mapping wherever it is one code point
    return lo if len(lo) == 1 else ch    # U+0130 -> "i̇" (2 code points) stays U+0130 under .NET's invariant casing


def test_to_lower_invariant_covers_every_code_point():
    """ToLowerInvariant (RecallSearchService.cs:96,110) is the Unicode simple lower-case mapping of EVERY code point, not
    just Latin/Greek/Cyrillic: the library's run table, the C oracle's pair table and the numpy oracle are each compared
    with Python's unicodedata over all 0x110000 code points (tokens "<cp>q", 400 per query string)."""
    from oracle import oracle_np
    from omni_recall_rag_b200.store import _lower_invariant
    cps = [c for c in range(0x80, 0x110000) if not (0xD800 <= c <= 0xDFFF) and chr(c).lower() != chr(c)]
    assert len(cps) > 1300 and 0x130 in cps and 0x10400 in cps and 0x1E900 in cps
    cps += [0xDF, 0x131, 0x149, 0x3C2, 0x4E2D, 0x1F600, 0x10FFFF]       # unmapped neighbours
    for at in range(0, len(cps), 400):
        part = cps[at:at + 400]
        text = " ".join(chr(c) + "q" for c in part)
        exp = list(dict.fromkeys(_py_lower_invariant(chr(c)) + "q" for c in part))
        assert oracle_c.query_terms(text) == exp
        assert oracle_np.query_terms(text) == exp
        assert list(orr.tokenize_query(text)) == [orr.hash_term(t) for t in exp]
        assert list(orr.tokenize_content(text)) == [orr.hash_term(t) for t in exp]
        assert _distinct_lower_tokens(text) == exp
        assert _lower_invariant(text) == " ".join(_py_lower_invariant(chr(c)) + "q" for c in part)
    assert oracle_c.query_terms("İstanbul KELVINK \U00010400") == ["İstanbul", "kelvink", "\U00010428"]
    assert oracle_c.keyword("ⱥ", "xȺx") == 1.0                          # 2-byte upper -> 3-byte lower in UTF-8


def test_hash_properties_on_the_synthetic_vocabulary():
    ids = np.arange(0, 1 << 20, 37, dtype=np.uint32)
    h = np.array([orr.hash_term(synth.term_text(i)) for i in ids], dtype=np.uint64)
    assert np.all(h != 0)
    assert len(np.unique(h)) == len(h)                       # no 64-bit collisions
    assert orr.hash_term("t0000001") != orr.hash_term("t0000010")


def test_synth_rows_are_deterministic_unit_and_grouped():
    spec = synth.make_spec(3072)
    a = synth.rows_host(spec, 1000, 96)
    b = synth.rows_host(spec, 1000 + 32, 32)
    assert np.array_equal(a.emb[32:64], b.emb) and np.array_equal(a.ticks[32:64], b.ticks)
    assert np.array_equal(a.term_ids[32:64], b.term_ids)
    norms = np.linalg.norm(a.emb.astype(np.float64), axis=1)
    live = norms > 0
    assert np.allclose(norms[live], 1.0, atol=1e-6)
    # documents: rows sharing doc_first_row share one timestamp (DocumentIngestionService.cs:102)
    for d in np.unique(a.doc_first_row):
        assert len(np.unique(a.ticks[a.doc_first_row == d])) == 1
    assert np.all(a.ticks <= spec.now_ticks) and np.all(a.ticks > spec.now_ticks - 365 * DAY)
    for row in a.term_ids:
        assert len(set(row.tolist())) == 64
    # truncated embeddings = prefix of the unit 3072-d vector
    t = synth.rows_host(synth.make_spec(768), 1000, 8)
    assert np.array_equal(t.emb, a.emb[:8, :768])


def test_synth_zero_rows_duplicates_and_zipf():
    spec = synth.make_spec(64, gen_dim=64, zero_row_ppm=50000, dup_row_ppm=100000, terms_per_chunk=16)
    r = synth.rows_host(spec, 0, 4000)
    zero = np.all(r.emb == 0, axis=1)
    assert 0.02 < zero.mean() < 0.09
    # duplicates: some later row repeats an earlier row's embedding and terms exactly
    key = {}
    dups = 0
    for i in range(4000):
        k = (r.emb[i].tobytes(), r.term_ids[i].tobytes())
        if k in key and not zero[i]:
            dups += 1
        key.setdefault(k, i)
    assert 200 < dups < 700
    # Zipf-like: about half of all draws come from the 1023 most frequent tokens
    assert 0.35 < (r.term_ids < 1023).mean() < 0.65


def test_synth_query_kinds():
    spec = synth.make_spec(256, gen_dim=256)
    planted = 0
    for qi in range(60):
        q = synth.query_host(spec, qi, 5000, n_terms=4)
        assert abs(np.linalg.norm(q.q.astype(np.float64)) - 1.0) < 1e-6
        assert len(set(q.term_ids.tolist())) == 4 and q.terms.n_terms == 4
        src = synth.rows_host(spec, 0, 5000).emb if qi == 0 else None
        if src is not None:
            all_rows = src
        sims = all_rows.astype(np.float64) @ q.q.astype(np.float64)
        planted += sims.max() > 0.9
    assert 1 <= planted <= 15                                   # ~10 % of queries sit next to a corpus row
    q16 = synth.query_host(spec, 3, 5000, n_terms=16, frequent_terms=8)
    assert (q16.term_ids[:8] < 1023).all() and len(set(q16.term_ids.tolist())) == 16


def test_merge_hits_applies_the_reference_tie_chain():
    nan = float("nan")
    a = orr.Hits(np.array([5, 9, 2], dtype=np.uint64), np.array([0.9, 0.5, nan]), np.array([10, 7, 3], dtype=np.int64))
    b = orr.Hits(np.array([105, 100, 101], dtype=np.uint64), np.array([0.9, 0.5, 0.5]), np.array([12, 7, 7], dtype=np.int64))
    m = orr.merge_hits([a, b], 6)
    assert m.rows.tolist() == [105, 5, 9, 100, 101, 2]          # score desc, ticks desc, row asc, NaN last
    assert orr.merge_hits([a, b], 0).rows.tolist() == [105]     # Math.Max(1, topK)
    assert len(orr.merge_hits([orr.Hits(np.zeros(0, np.uint64), np.zeros(0), np.zeros(0, np.int64))], 5)) == 0


def test_citation_helpers_match_the_oracle():
    for x in [0.30000000000000004, 0.8999999999999998, 0.12345, 0.00005, 2.5e-4, 0.99995, 1.0, 0.0]:
        assert R.math_round4(x) == oracle_c.round4(x)
    for s in ["  a\nb\r\n c  ", "x" * 200, "short", " " * 5, "é" * 181]:
        assert R.build_snippet(s, 180) == oracle_c.snippet(s, 180)


def test_shard_rows_partition():
    for total in [0, 1, 7, 1000, 40_000_000]:
        for world in [1, 2, 4, 8]:
            spans = [sharded.shard_rows(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (b0, n0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + n0 == b1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _gloo_worker(rank, world, port, total_rows, top_k, out_q):
    import torch.distributed as dist
    from tests.util import oracle_search_synth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = synth.make_spec(64, gen_dim=64, terms_per_chunk=8, dup_row_ppm=20000)
    base, n_local = sharded.shard_rows(total_rows, world, rank)
    rows = synth.rows_host(spec, base, n_local)

    def local_search(q, terms, now, k):
        r, s, t = oracle_search_synth(rows, query, now, k)
        return orr.Hits(r + np.uint64(base), s, t)

    sr = sharded.ShardedRecall(local_search=local_search)
    results = []
    for qi in range(4):
        query = synth.query_host(spec, qi, total_rows, n_terms=3)
        h = sr.search(query.q, query.terms, spec.now_ticks, top_k)
        results.append((h.rows.tolist(), h.scores.tolist(), h.ticks.tolist()))
    out_q.put((rank, results))
    dist.destroy_process_group()


def test_sharded_search_world2_gloo_matches_single_shard():
    """N>1 host logic on CPU: two ranks, each scoring its row block (the oracle stands in for
    the local GPU scan), all-gather over gloo, merge == the single-shard oracle result."""
    import torch.multiprocessing as mp
    from tests.util import oracle_search_synth

    total, top_k, world = 3001, 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, total, top_k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = synth.make_spec(64, gen_dim=64, terms_per_chunk=8, dup_row_ppm=20000)
    rows = synth.rows_host(spec, 0, total)
    for qi in range(4):
        query = synth.query_host(spec, qi, total, n_terms=3)
        r, s, t = oracle_search_synth(rows, query, spec.now_ticks, top_k)
        for rank in range(world):
            assert got[rank][qi] == (r.tolist(), s.tolist(), t.tolist())


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU port on the host cores) needs no GPU; stdout must be ONE JSON line with
    the contract's keys even when a library prints to fd 1 meanwhile (bench.py routes such output to stderr)."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample-rows", "1500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["metric"].startswith("hybrid recall QPS at 1M x 3072") and j["unit"].startswith("queries/s")
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["n_gpus"] == 1 and j["steps"] == 2


def test_bench_without_a_gpu_fails_loudly_instead_of_falling_back():
    """The product arm has no CPU path: without CUDA bench.py exits non-zero and says so."""
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def _build_shim_driver(tmp_path):
    import subprocess

    exe = str(tmp_path / "shim_driver")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "shim", "shim_driver.c"), "-ldl", "-o", exe])
    return exe


def test_csharp_struct_declarations_match_the_c_layout(tmp_path):
    """SURVEY 8 f3 without a .NET SDK: the structs DECLARED in dotnet/OrrNative.cs ([StructLayout(Sequential)], parsed from
    the C# source) must lay out exactly as the C compiler lays out orr_config / orr_hit (tests/shim/shim_driver.c
    --layout, compiled against include/orr.h) and as the ctypes binding declares them."""
    import ctypes as C
    import json
    import re
    import subprocess

    c_layout = json.loads(subprocess.check_output([_build_shim_driver(tmp_path), "--layout"]))
    src = open(os.path.join(ROOT, "dotnet", "OrrNative.cs"), encoding="utf-8").read()
    size_of = {"int": 4, "uint": 4, "long": 8, "ulong": 8, "double": 8, "float": 4, "byte": 1}

    def cs_fields(struct):
        body = re.search(r"\[StructLayout\(LayoutKind\.Sequential\)\]\s*internal struct " + struct + r"\s*\{(.*?)\}", src, re.S).group(1)
        fields = []
        for typ, names in re.findall(r"public\s+(\w+)\s+([^;]+);", body):
            fields += [(n.strip(), size_of[typ]) for n in names.split(",")]
        return fields

    def sequential(fields):                       # .NET sequential layout = C natural alignment (Pack = 0)
        off, out, align = 0, [], 1
        for name, sz in fields:
            off = (off + sz - 1) // sz * sz
            out.append(off)
            off += sz
            align = max(align, sz)
        return out, (off + align - 1) // align * align

    snake = lambda s: re.sub(r"(?<!^)(?=[A-Z])", "_", s).lower()
    for cs_name, c_name, ctype in (("OrrConfig", "orr_config", N.OrrConfig), ("OrrHit", "orr_hit", N.OrrHit)):
        fields = cs_fields(cs_name)
        offs, size = sequential(fields)
        assert size == c_layout[f"sizeof.{c_name}"] == C.sizeof(ctype), cs_name
        assert len(fields) == len(ctype._fields_)
        for (name, _), off, (py_name, _) in zip(fields, offs, ctype._fields_):
            assert snake(name) == py_name, (name, py_name)                       # same field order on all three sides
            assert off == c_layout[f"{c_name}.{py_name}"] == getattr(ctype, py_name).offset, (cs_name, name)
    assert C.sizeof(N.OrrTiming) == c_layout["sizeof.orr_timing"]
    assert c_layout["ORR_ABI_VERSION"] == N.ORR_ABI_VERSION
    # every entry point the C# binding imports exists in the header and in the library
    imported = set(re.findall(r"partial\s+\w+\s+(orr_\w+)\s*\(", src))
    assert imported and imported <= set(N.declared_symbols()), imported - set(N.declared_symbols())
