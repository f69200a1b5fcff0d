#!/usr/bin/env python3
"""Latency of small batches: thread-per-problem instances against the warp-per-problem kernel (option force_wpp) over batch sizes.
Device-resident data, CUDA events, best of 20.  Usage: python profiles/tools/small_batch.py [config ...]"""
import importlib, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
specs = dict(quadrotor=P.quadrotor, cartpole=P.cartpole, rocket=P.rocket, quadrotor_adaptive=lambda: P.quadrotor(adaptive=True))
dev = torch.device("cuda:0")
for cfg in (sys.argv[1:] or ["quadrotor", "quadrotor_adaptive"]):
    spec = specs[cfg]()
    n, m, N = spec.nx, spec.nu, spec.N
    base = P.make_batch(spec, 1 << 17, 1.0, seed=99)
    for B in (1, 32, 256, 1024, 4096, 16384, 65536, 131072):
        t = lambda a: None if a is None else torch.from_numpy(a[:B]).to(dev)
        x0, Xr, Ur = t(base.x0), t(base.Xref), t(base.Uref)
        x = torch.empty((B, N, n), device=dev); u = torch.empty((B, N - 1, m), device=dev)
        it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
        ptr = lambda a: None if a is None else a.data_ptr()
        row = dict(config=cfg, batch=B)
        for wpp in (0, 1):
            s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0]); s.cuda.set_option("force_wpp", wpp)
            stream = torch.cuda.current_stream().cuda_stream
            best = 1e9
            for r in range(22):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s.cuda.solve_batch_device(B, ptr(x0), ptr(Xr), ptr(Ur), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream)
                e1.record(); torch.cuda.synchronize()
                if r > 1: best = min(best, e0.elapsed_time(e1))
            row["wpp_ms" if wpp else "tpp_ms"] = round(best, 4)
            row["wpp_kernel" if wpp else "tpp_kernel"] = s.cuda.last_kernel
            row["iters_wpp" if wpp else "iters_tpp"] = int(it.sum().item())
        print(json.dumps(row), flush=True)
