#!/usr/bin/env python3
"""Summarise one kernel launch of an ncu report as JSON: ncu_summary.py prof.ncu-rep [launch_index] > summary.json
Reads `ncu -i <rep> --page raw --csv`; keeps the metrics the round's README cites (time, DRAM bytes, pipe utilisation,
issue activity, occupancy, stall reasons per issue)."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[idx]
keep = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_", "sm__pipe_", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__cycles_active.avg")
d = {}
for h, u, v in zip(hdr, units, r):
    if any(h.startswith(k) for k in keep) and v != "":
        d[h] = f"{v} {u}".strip()
print(json.dumps(d, indent=1, sort_keys=True))
