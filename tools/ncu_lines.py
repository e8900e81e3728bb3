#!/usr/bin/env python
"""Attribute warp-stall samples of an .ncu-rep to CUDA source lines.  usage: ncu_lines.py rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; out = []
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if hdr is None:
        hdr = r; idx = {}
        for i, h in enumerate(hdr): idx.setdefault(h, i)
        continue
    if r[idx['Address']] != '-': continue          # keep the per-line aggregate rows only
    try: n = int(r[idx['# Samples']])
    except Exception: continue
    if n > 0:
        st = {h: int(r[i]) for h, i in idx.items() if h.startswith('stall_') and 'Not Issued' not in h and r[i].isdigit() and int(r[i]) > 0.15 * n}
        out.append((n, cur, r[idx['Line No']], r[1].strip()[:80], st))
tot = sum(o[0] for o in out)
print("total samples", tot)
for n, f, l, src, st in sorted(out, reverse=True)[:ntop]:
    print(f"{n:8d} {100*n/tot:5.1f}% {f}:{l:>4s}  {src}  {st}")
