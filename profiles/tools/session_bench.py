#!/usr/bin/env python3
"""Closed-loop throughput of device-resident sessions: B warm-started solvers, `steps` times (solve + x0 <- A x0 + B u0).
Wall clock around the C-ABI calls with a device synchronise (the calls are synchronous).  Usage: session_bench.py [config] [B] [steps] [precision]"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
cfg = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 32
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole, rocket=P.rocket, quadrotor_adaptive=lambda: P.quadrotor(adaptive=True))[cfg]()
b = P.make_batch(spec, B, 0.3, seed=11)
s = tm.TinyMPC().setup_from_spec(spec, devices=[0]); s.cuda.set_option("precision", prec)
ses = s.cuda.session(B)
if b.Xref is not None: ses.set_x_ref(b.Xref.astype(np.float64))
ses.set_x0(b.x0.astype(np.float64))
rows = []
for t in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ses.solve(); ses.step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = ses.read("iter")
    rows.append((dt, float(it.mean()), int(it.max())))
tot = sum(r[0] for r in rows[1:]); iters = sum(r[1] for r in rows[1:]) * B
print(json.dumps(dict(config=cfg, batch=B, precision=prec, steps=steps, first_step_ms=round(rows[0][0] * 1e3, 3), first_step_mean_iters=rows[0][1],
                      warm_ms_per_step=round(tot / (steps - 1) * 1e3, 3), warm_mean_iters=round(iters / B / (steps - 1), 2),
                      warm_solves_per_sec=round(B * (steps - 1) / tot), warm_ns_per_admm_iter=round(tot * 1e9 / max(iters, 1), 3))))
ses.close()
