#!/usr/bin/env python3
"""Host<->device copy bandwidth of the box (pinned memory, cudaMemcpyAsync through torch): H2D alone, D2H alone, both directions
at once on two streams.  The e2e number of bench.py moves 0.70 GB in and 0.66 GB out per 2^20-problem quadrotor step, so these
figures are its ceiling.  Usage: python profiles/microbench/pcie_bench.py"""
import json, time, torch
dev = torch.device("cuda", 0)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 2)
res = {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "bidir_each_GBps": run(True, True), "bytes": n}
print(json.dumps(res))
