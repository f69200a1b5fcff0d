#!/usr/bin/env python3
"""How much of the persistent kernel's time is its tail?  The same 2^20 quadrotor problems (the bench batch) solved device-resident
in different ORDERS: as generated, sorted by the expected difficulty |Kinf (x0 - xref)| / u_bound (hardest first: longest-processing-
time scheduling), and with the sure-easy ones (key < 0.6) moved to the end.  No kernel change: the order of the input arrays is the
claim order of the lanes.  Usage: python profiles/tools/order_study.py [config]"""
import importlib, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
cfg = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole)[cfg]()
B = 1 << 20
b = P.make_batch(spec, B, 1.0, seed=1237)
solver = tm.TinyMPC(); solver.setup_from_spec(spec, devices=[0])
A, Bm = np.asarray(spec.A, np.float64), np.asarray(spec.B, np.float64).reshape(spec.nx, spec.nu)
Q1, R1 = np.diag(spec.Qdiag) + 2 * spec.rho * np.eye(spec.nx), np.diag(spec.Rdiag) + 2 * spec.rho * np.eye(spec.nu)
Pm = Q1.copy()
for _ in range(1000):      # infinite-horizon gain of the rho-augmented problem (tiny_api.cpp:244-318; rho enters twice)
    K = np.linalg.solve(R1 + Bm.T @ Pm @ Bm, Bm.T @ Pm @ A)
    Pm = Q1 + A.T @ Pm @ (A - Bm @ K)
d = b.x0.astype(np.float64) - (b.Xref[:, 0, :] if b.Xref is not None else 0)
key = (np.abs(d @ K.T) / np.minimum(-spec.u_min[0], spec.u_max[0])).max(1)
orders = {"as generated": np.arange(B), "hardest first (sorted by key)": np.argsort(-key, kind="stable"),
          "sure-easy (key < 0.6) last": np.concatenate([np.where(key >= 0.6)[0], np.where(key < 0.6)[0]]),
          "easiest first (sorted, worst case)": np.argsort(key, kind="stable")}
dev = torch.device("cuda", 0)
n, m, N = spec.nx, spec.nu, spec.N
for mode, band in (("exact-count", P.exact_band(spec)), ("plain fp32", 0.0)):
    solver.cuda.set_option("mixed", band)
    ref_it = None
    for name, o in orders.items():
        take = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a[o])).to(dev)
        x0, Xr, Ur = take(b.x0), take(b.Xref), take(b.Uref)
        x = torch.empty((B, N, n), device=dev); u = torch.empty((B, N - 1, m), device=dev)
        it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream()
        step = lambda: solver.cuda.solve_batch_device(B, ptr(x0), ptr(Xr), ptr(Ur), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            step()
        e1.record(stream); torch.cuda.synchronize()
        inv = np.empty(B, np.int64); inv[o] = np.arange(B)
        its = it.cpu().numpy()[inv]
        if ref_it is None:
            ref_it = its
        print(json.dumps(dict(config=cfg, mode=mode, order=name, ms_per_step=round(e0.elapsed_time(e1) / 10, 3), same_counts=bool((its == ref_it).all()),
                              mean_iters=float(its.mean()), kernel=solver.cuda.last_kernel)), flush=True)
