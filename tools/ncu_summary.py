#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + stall-sample totals + top stalled SASS lines.  usage: ncu_summary.py rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 15
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
print("metric,unit,value")
for h, u, v in zip(hdr, units, vals):
    if h in want or h.endswith('TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'): print(f"{h},{u},{v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
cats = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[idx['# Samples']]) for r in data)
print(f"# stall samples total={tot}")
for c in sorted(cats, key=lambda c: -sum(int(r[idx[c]]) for r in data)):
    s = sum(int(r[idx[c]]) for r in data)
    if s: print(f"{c},{s},{100.0*s/tot:.1f}%")
print("# top instructions by samples")
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:ntop]:
    st = {c: int(r[idx[c]]) for c in cats if int(r[idx[c]]) > 0.1 * int(r[idx['# Samples']])}
    print(r[idx['# Samples']], '|', r[idx['Source']].strip()[:70], '|', st)
