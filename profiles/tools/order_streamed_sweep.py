#!/usr/bin/env python3
"""Compact streamed pipeline with claim order: SMs left to the ordering kernels x share of the shard claimed in index order.
Usage: python profiles/tools/order_streamed_sweep.py [quadrotor|cartpole]"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
cfg = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole)[cfg]()
B = 1 << 20
b = P.make_batch(spec, B, 1.0, seed=1237)
pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
x0, xc = pin(b.x0), (None if b.Xref is None else pin(b.Xref[:, 0, :]))
ipin = lambda: torch.empty(B, dtype=torch.int32).pin_memory().numpy()
out = dict(u0=torch.empty((B, spec.nu)).pin_memory().numpy(), iter=ipin(), status=ipin())
s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("mixed", P.exact_band(spec))
for order, sms, div in [(0, 2, 4), (1, 2, 4), (1, 1, 4), (1, 4, 4), (1, 2, 8), (1, 1, 8), (1, 2, 2), (1, 1, 2), (0, 2, 4)]:
    s.cuda.set_option("order", order); s.cuda.set_option("order_sms", sms); s.cuda.set_option("order_from_div", div)
    s.cuda.solve_batch(x0, xref_const=xc, out=out, compact_out=True)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); s.cuda.solve_batch(x0, xref_const=xc, out=out, compact_out=True); ts.append(time.perf_counter() - t0)
    print(json.dumps(dict(config=cfg, order=order, order_sms=sms, order_from_div=div, ms=round(min(ts) * 1e3, 3), median_ms=round(sorted(ts)[3] * 1e3, 3),
                          Msolves_s=round(B / min(ts) / 1e6, 2), kernel_ms=round(s.cuda.last_timing()["kernel_ms"], 3),
                          checksum=[int(out['iter'].sum()), int(out['status'].sum())])), flush=True)
