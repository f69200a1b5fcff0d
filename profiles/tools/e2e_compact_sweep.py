#!/usr/bin/env python3
"""End-to-end throughput of the compact I/O mode (x0 + one reference state in, u0 + iter + status out) of tinympc_cuda_solve_batch
in the exact-count mode over chunk counts.  Usage: python profiles/tools/e2e_compact_sweep.py [band]"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
spec = P.quadrotor()
band = float(sys.argv[1]) if len(sys.argv) > 1 else P.exact_band(spec)
B = 1 << 20
b = P.make_batch(spec, B, 1.0, seed=1237)
n, m, N = spec.nx, spec.nu, spec.N
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
x0, xc = pin(b.x0), pin(b.Xref[:, 0, :])
ipin = lambda: torch.empty(B, dtype=torch.int32).pin_memory().numpy()
out = dict(u0=torch.empty((B, m)).pin_memory().numpy(), iter=ipin(), status=ipin())
s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("mixed", band)
# (compact_streamed, compact_in_kernel, compact_early_d2h, chunks): one launch chain behind an arrival watermark (result copies under
# the fp64 pass | at the end) | chunked launches, compact I/O inside the kernels | chunked launches, reference replicated on the
# device first and u0 gathered afterwards (the pipeline before the streamed form)
# order: the later three quarters of the shard claimed hardest-first through a list built while the first quarter is being solved
for streamed, in_kernel, early, chunks, order in [(1, 1, 1, 0, 1), (1, 1, 1, 0, 0), (1, 1, 0, 0, 1), (1, 1, 1, 4, 1), (1, 1, 1, 16, 1), (0, 1, 1, 0, 1), (0, 1, 1, 1, 1),
                                                  (0, 0, 1, 0, 0), (0, 0, 1, 1, 0), (0, 0, 1, 4, 0)]:
    s.cuda.set_option("order", order)
    s.cuda.set_option("compact_streamed", streamed)
    s.cuda.set_option("compact_in_kernel", in_kernel)
    s.cuda.set_option("compact_early_d2h", early)
    s.cuda.set_option("chunks", chunks)
    s.cuda.solve_batch(x0, xref_const=xc, out=out, compact_out=True)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s.cuda.solve_batch(x0, xref_const=xc, out=out, compact_out=True); ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(json.dumps(dict(compact_streamed=streamed, compact_in_kernel=in_kernel, compact_early_d2h=early, chunks=chunks, order=order, ms=round(t * 1e3, 3), Msolves_s=round(B / t / 1e6, 2), median_ms=round(sorted(ts)[2] * 1e3, 3),
                          pipeline=s.cuda.last_timing(), kernel=s.cuda.last_kernel,
                          checksum=[int(out['iter'].sum()), int(out['status'].sum()), float(np.abs(out['u0'].astype(np.float64)).sum())])))
