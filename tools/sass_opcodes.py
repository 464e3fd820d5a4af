#!/usr/bin/env python
"""Static SASS opcode counts per kernel of the shipped library (the tcgen05 / TMA evidence the profiling recipe asks for).
python tools/sass_opcodes.py [liborr.so] > profiles/rNN_sass_opcodes.md      (needs cuobjdump and c++filt; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "omni_recall_rag_b200", "liborr.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
COLS = [("UTCHMMA", r"^UTCHMMA"), ("UTCBAR", r"^UTCBAR"), ("LDTM", r"^LDTM"), ("UTMALDG", r"^UTMALDG"), ("UBLKCP", r"^UBLKCP"),
        ("SYNCS", r"^SYNCS"), ("REDUX", r"^REDUX"), ("MATCH", r"^MATCH"), ("DADD+DFMA+DMUL", r"^(DADD|DFMA|DMUL)"),
        ("F2F.F64.F32", r"^F2F\.F64\.F32"), ("LDG.256", r"^LDG\..*\b256\b|^LDG\.E\.ENL2\.256"), ("HMMA/IMMA", r"^(HMMA|IMMA)")]
kern, cur = collections.OrderedDict(), None
spell = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["instr"] += 1
        for name, pat in COLS:
            if re.search(pat, op): kern[cur][name] += 1
        if re.match(r"(UTCHMMA|UTCBAR|LDTM|UTMALDG|UBLKCP|UTCATOMSWS|UTMAPF)", op): spell[op] += 1
names = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)          # argument list
    return n.replace("(int)", "").replace("(bool)", "")
print("# SASS opcode counts per kernel (`cuobjdump -sass omni_recall_rag_b200/liborr.so`, sm_100a; `python tools/sass_opcodes.py`)\n")
print("Static instruction counts in the shipped library (built by `__graft_entry__.build()`): `UTCHMMA` = tcgen05.mma (`.2CTA` forms: cta_group::2),\n"
      "`UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` = TMA tensor load (`cp.async.bulk.tensor`), `UBLKCP` = `cp.async.bulk`\n"
      "(TMA 1-D bulk copy), `SYNCS` = mbarrier operations, `REDUX` = warp reductions, `DADD/DFMA/DMUL` + `F2F.F64.F32` = the fp64 re-score arithmetic\n"
      "(fp32 products widened to fp64).  No `HMMA`/`IMMA` (mma.sync) anywhere: the tensor work is tcgen05 only.  Template arguments:\n"
      "`orr_scan_kernel<NV, TR, PIPE>`, `orr_batch_gemm_kernel<PASSES, MODE, CLUSTER>` (MODE 0 = main pass, 1 = sampling pass; CLUSTER 2 = one CTA pair,\n"
      "4 = two pairs with query multicast), `orr_noemb_scores_kernel<SLOTS/32, WIDE>`.\n")
print("| kernel | instr | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for (mangled, c), nice in sorted(zip(kern.items(), names), key=lambda t: -t[0][1]["instr"]):
    print(f"| `{short(nice)}` | {c['instr']} | " + " | ".join(str(c[name]) for name, _ in COLS) + " |")
print("\nOpcode spellings found (whole library): " + ", ".join(f"`{k}` x{v}" for k, v in spell.most_common()))
