#!/usr/bin/env python3
"""End-to-end (host buffers in, host buffers out) throughput of tinympc_cuda_solve_batch for the chunked-launch pipeline and the
streamed single-launch pipeline over chunk counts.  Usage: python profiles/tools/e2e_sweep.py [config] [variant]"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
cfg = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole, rocket=P.rocket, quadrotor_adaptive=lambda: P.quadrotor(adaptive=True))[cfg]()
B = 1 << 20
b = P.make_batch(spec, B, 1.0, seed=1237)
n, m, N = spec.nx, spec.nu, spec.N
pin = lambda a: None if a is None else torch.from_numpy(a).pin_memory().numpy()
x0, Xr, Ur = pin(b.x0), pin(b.Xref), pin(b.Uref)
out = dict(x=torch.empty((B, N, n)).pin_memory().numpy(), u=torch.empty((B, N - 1, m)).pin_memory().numpy(),
           iter=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), status=torch.empty(B, dtype=torch.int32).pin_memory().numpy())
s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("variant", variant)
for streamed, chunks in [(0, 0), (1, 0), (1, 8), (1, 16), (1, 24), (1, 32)]:
    s.cuda.set_option("streamed", streamed); s.cuda.set_option("chunks", chunks)
    s.cuda.solve_batch(x0, Xr, Ur, out=out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s.cuda.solve_batch(x0, Xr, Ur, out=out); ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(json.dumps(dict(config=cfg, variant=variant, streamed=streamed, chunks=chunks, ms=round(t * 1e3, 3), Msolves_s=round(B / t / 1e6, 2),
                          median_ms=round(sorted(ts)[2] * 1e3, 3), pipeline=s.cuda.last_timing(), kernel=s.cuda.last_kernel)))
