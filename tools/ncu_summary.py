#!/usr/bin/env python
"""Summarise an ncu raw CSV (+ optional SASS source CSV): key metrics, stall reasons, opcode mix.
python tools/ncu_summary.py raw.csv [sass.csv]"""
import csv, sys, collections
def load(p):
    rows=list(csv.reader(open(p))); hdr=rows[0]; u=rows[1]; v=rows[2]
    return {h:(v[i],u[i]) for i,h in enumerate(hdr)}
d=load(sys.argv[1])
keys=['gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','smsp__warps_active.avg.per_cycle_active']
for k in keys:
    for kk in d:
        if kk==k or kk.endswith(k): print(f"| {k} | {d[kk][0]} | {d[kk][1]} |"); break
st={k:float(x[0].replace(',','')) for k,x in d.items() if 'smsp__pcsamp_warps_issue_stalled' in k and not k.endswith('_not_issued') and x[0].replace(',','').replace('.','').isdigit()}
tot=sum(st.values())
print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_','')} {100*x/tot:.1f}%" for k,x in sorted(st.items(), key=lambda t:-t[1])[:10]))
if len(sys.argv)>2:
    rows=list(csv.reader(open(sys.argv[2]))); hdr=rows[1]; data=rows[2:]
    ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
    h=collections.Counter(); hs=collections.Counter(); tot=ts=0
    for r in data:
        if not r[ie].isdigit(): continue
        t=r[ia].strip().split()
        op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        h[op]+=int(r[ie]); hs[op]+=int(r[isamp]); tot+=int(r[ie]); ts+=int(r[isamp])
    print("total warp inst", tot)
    print("opcode mix:", ", ".join(f"{op} {100*c/tot:.1f}% (samples {100*hs[op]/ts:.1f}%)" for op,c in h.most_common(14)))
    # hottest sampled instructions
    top=sorted((r for r in data if r[isamp].isdigit()), key=lambda r:-int(r[isamp]))[:14]
    for r in top: print(f"   {int(r[isamp]):7d}  {r[ia].strip()[:90]}")
