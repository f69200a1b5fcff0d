#!/usr/bin/env python3
"""Time every compiled kernel variant of a config on the GPU (device-resident data, CUDA events).
Tuning tool, not part of the product; uses the CPU oracle only to obtain the family cache and a
parity spot check."""
import importlib, sys, json
from pathlib import Path
import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
P = importlib.import_module("tinympc-matlab_b200.problems")
capi = importlib.import_module("tinympc-matlab_b200.capi")
import oracle as O, cases


def run(p, scale, B, variants, precision=32, reps=3, options=None):
    b = P.make_batch(p, B, scale)
    dev = torch.device("cuda:0")
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    x0, Xref, Uref = t(b.x0), t(b.Xref), t(b.Uref)
    x = torch.empty((B, p.N, p.nx), device=dev); u = torch.empty((B, p.N - 1, p.nu), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    ptr = lambda a: None if a is None else a.data_ptr()
    fam = cases.family_from_spec(p, O.get_cache(p, "port"))
    nchk = 2000
    g = O.solve_batch(p, b.slice(0, nchk), "ref" if O.available("ref") else "port")
    for v in variants:
        s = capi.CudaSolver(); s.set_option("precision", precision); s.set_option("variant", v); s.set_family(fam)
        for ok, ov in (options or {}).items(): s.set_option(ok, ov)
        stream = torch.cuda.current_stream().cuda_stream
        best = 1e9
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.solve_batch_device(B, ptr(x0), ptr(Xref), ptr(Uref), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream)
            e1.record(); torch.cuda.synchronize()
            if r > 0: best = min(best, e0.elapsed_time(e1))
        iters = int(it.sum().item())
        same = (it[:nchk].cpu().numpy() == g["iter"])
        dx = np.abs(x[:nchk].cpu().numpy()[same] - g["x"][same]).max()
        print(json.dumps(dict(options=options, cfg=p.name, scale=scale, B=B, variant=v, kernel=s.last_kernel, ms=round(best, 3),
                              solves_per_s=round(B / best * 1e3), ns_per_iter=round(best * 1e6 / iters, 4),
                              mean_iters=round(iters / B, 2), count_mismatch=int((~same).sum()), dx=float(dx))), flush=True)
        s.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "occupancy":
        # issue efficiency against resident warps: the small cartpole shape fits up to 4 CTAs of 7 warps per SM
        for c in (1, 2, 3, 4):
            run(P.cartpole(N=10), 1.0, 1 << 20, [0], options={"ctas_per_sm": c})
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    run(P.quadrotor(), 0.3, B, [0, 3, 2, 1])
    run(P.quadrotor(), 1.0, B, [0, 3, 2, 1])
    run(P.cartpole(), 0.3, B, [0, 4, 3, 2, 1])
    run(P.cartpole(), 1.0, B, [0, 4, 3, 2, 1])
    run(P.rocket(), 1.0, B // 4, [0])
    run(P.quadrotor(adaptive=True), 1.0, B // 4, [0])
    run(P.quadrotor(), 1.0, B // 8, [0], precision=64)
