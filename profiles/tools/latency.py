#!/usr/bin/env python3
"""Latency path: (1) device-resident batches of 1 .. 65 536 quadrotor problems in plain fp32 (thread per problem), fp64 (lane group per
problem) and the default exact-count mode, CUDA events, best of 20; (2) one tiny_solve (the reference's single-solver entry) on the
lane-group kernel and on the warp-per-problem kernel, wall clock per call and per ADMM iteration, next to the reference C++ on one
host core.  Test infrastructure for (2) (it calls the oracle).  Usage: python profiles/tools/latency.py"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
dev = torch.device("cuda:0")
spec = P.quadrotor()
n, m, N = spec.nx, spec.nu, spec.N
base = P.make_batch(spec, 1 << 16, 1.0, seed=99)
for B in (1, 32, 256, 1024, 4096, 16384, 65536):
    t = lambda a: None if a is None else torch.from_numpy(a[:B]).to(dev)
    x0, Xr, Ur = t(base.x0), t(base.Xref), t(base.Uref)
    x = torch.empty((B, N, n), device=dev); u = torch.empty((B, N - 1, m), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    ptr = lambda a: None if a is None else a.data_ptr()
    row = dict(config="quadrotor", batch=B)
    for name, prec, band in (("fp32", 32, 0.0), ("fp64", 64, 0.0), ("exact", 32, P.exact_band(spec))):
        s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("precision", prec); s.cuda.set_option("mixed", band)
        stream = torch.cuda.current_stream().cuda_stream
        best = 1e9
        for r in range(22):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.cuda.solve_batch_device(B, ptr(x0), ptr(Xr), ptr(Ur), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream)
            e1.record(); torch.cuda.synchronize()
            if r > 1: best = min(best, e0.elapsed_time(e1))
        row[name + "_ms"] = round(best, 4); row[name + "_kernel"] = s.cuda.last_kernel
        row["max_iter_" + name] = int(it.max().item())
    print(json.dumps(row), flush=True)

# (2) one tiny_solve: cartpole G2 (51 iterations) and quadrotor G3 (100 iterations)
import oracle as O
impl = "ref" if O.available("ref") else "port"
def one(spec, x0, xref, tag):
    for force in (0, 1):
        s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("force_wpp", force)
        ts = []
        for r in range(30):
            s.reset_workspace() if hasattr(s, "reset_workspace") else None
            s.set_x0(x0)
            if xref is not None: s.set_x_ref(xref)
            torch.cuda.synchronize(); t0 = time.perf_counter(); s.solve(); ts.append(time.perf_counter() - t0)
        iters = s.get_stats()["iter"]
        tbest = min(ts[3:])
        print(json.dumps(dict(case=tag, kernel=s.cuda.last_kernel, iters=int(iters), call_us=round(tbest * 1e6, 1), us_per_iter=round(tbest * 1e6 / max(iters, 1), 2))), flush=True)
    ts = []
    for r in range(12):
        ses = O.Session(spec, impl)          # a fresh (cold) reference solver per repetition; only solve() is timed
        if xref is not None: ses.set_x_ref(xref)
        ses.set_x0(x0)
        t0 = time.perf_counter(); rr = ses.solve(); ts.append(time.perf_counter() - t0)
        if r < 11: ses.close()
    print(json.dumps(dict(case=tag, kernel="reference C++ (1 host core, incl. ctypes call)", iters=int(rr["iter"]), call_us=round(min(ts[3:]) * 1e6, 1),
                          us_per_iter=round(min(ts[3:]) * 1e6 / max(int(rr["iter"]), 1), 2))), flush=True)
    ses.close()
pc = P.cartpole(matlab_defaults=True)
one(pc, np.array([0.5, 0, 0, 0.0]), None, "cartpole G2 (u in +-0.5, tol 1e-4)")
pq = P.quadrotor()
xr = np.zeros((pq.N, 12)); xr[:, 2] = 2.0
one(pq, np.array([0, 1, 0, 0.2, 0, 0, 0.1, 0, 0, 0, 0, 0.0]), xr, "quadrotor G3 (hover, 100 iterations)")
