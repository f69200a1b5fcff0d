#!/usr/bin/env python3
"""Per-source-line cost of one kernel from an ncu report: executed warp instructions and stall samples by CUDA line.
Usage: line_hist.py prof.ncu-rep [top_n] [warp_iterations]   (needs a report captured with --import-source on, -lineinfo)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
div = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
ex, smp, txt = collections.Counter(), collections.Counter(), {}
cur_file = ""
hdr = None
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 8:
        continue
    try:
        n = int(r[hdr["Instructions Executed"]]); s = int(r[hdr["# Samples"]])
    except ValueError:
        continue
    if r[0]:   # a CUDA source line row carries the aggregate of its SASS rows
        key = (cur_file, int(r[0]))
        ex[key] += n; smp[key] += s; txt[key] = r[1].strip()[:110]
tot, stot = sum(ex.values()), sum(smp.values())
print(f"total warp instructions {tot}  samples {stot}" + (f"  per warp-iteration {tot / div:.0f}" if div else ""))
for k, v in ex.most_common(top):
    print(f"{k[0]}:{k[1]:5d} exec {100 * v / tot:5.2f}%  samples {100 * smp[k] / stot:5.2f}%" + (f" per-iter {v / div:7.1f}" if div else "") + f"  | {txt[k]}")
