#!/usr/bin/env python3
"""Opcode histogram of an `ncu --page source --csv` export, weighted by executed warp instructions
and by stall samples.  Usage: sass_hist.py src.csv [warp_iterations]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], encoding="latin-1")))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ex, smp, cnt = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = r[ci["Source"]].strip()
    toks = s.split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    if not toks: continue
    op = toks[0]
    key = op.split(".")[0]
    if key in ("LDS", "STS", "LDG", "STG", "FFMA2", "LDCU", "LDC"): key = ".".join(op.split(".")[:1])
    n = int(r[ci["Instructions Executed"]]); ex[key] += n; tot += n; cnt[key] += 1
    smp[key] += int(r[ci["# Samples"]])
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
stot = sum(smp.values())
print(f"total warp instructions {tot}" + (f"  per warp-iteration {tot/div:.0f}" if div else ""))
for k, v in ex.most_common(40):
    print(f"{k:12s} static {cnt[k]:5d}  exec {v:13d} {100*v/tot:6.2f}%  samples {100*smp[k]/stot:6.2f}%" + (f"  per-iter {v/div:8.1f}" if div else ""))
