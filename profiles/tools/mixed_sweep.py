#!/usr/bin/env python3
"""Mixed-precision exact-count mode: for every BASELINE config, sweep the relative band of option "mixed" and report
(a) how many problems the fp32 pass hands to the fp64 pass, (b) how many iteration-count / status mismatches remain against
the fp64 kernels (which reproduce the reference's counts exactly, tests/test_gpu_parity.py), (c) the time per batch.
Usage (GPU box): python profiles/tools/mixed_sweep.py [--batch 262144] [--out gpurun_out/mixed_sweep.jsonl]"""
import argparse
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 18)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "mixed_sweep.jsonl"))
    ap.add_argument("--bands", default="0,0.001,0.003,0.01,0.03,0.1,0.3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="0", help="kernel variants to sweep (0 = default, 5 = direct-form tpp2, 6 = incremental tpp3)")
    ap.add_argument("--configs", default="", help="comma list of config names to keep (default all)")
    a = ap.parse_args()
    tm = importlib.import_module("tinympc-matlab_b200")
    P = importlib.import_module("tinympc-matlab_b200.problems")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    out = open(a.out, "w")
    cfgs = [("quadrotor", P.quadrotor, 1.0), ("quadrotor", P.quadrotor, 0.3), ("cartpole", P.cartpole, 1.0), ("cartpole", P.cartpole, 0.3),
            ("rocket", P.rocket, 1.0), ("quadrotor_adaptive", lambda: P.quadrotor(adaptive=True), 1.0)]
    keep = [c for c in a.configs.split(",") if c]
    for name, mk, scale in cfgs:
        if keep and name not in keep:
            continue
        spec = mk()
        n, m, N = spec.nx, spec.nu, spec.N
        B = a.batch
        b = P.make_batch(spec, B, scale, seed=4321)
        s = tm.TinyMPC().setup_from_spec(spec, devices=[0])
        td = lambda v: None if v is None else torch.from_numpy(v).to(dev)
        x0, Xr, Ur = td(b.x0), td(b.Xref), td(b.Uref)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream()

        def run(precision, band, variant=0):
            s.cuda.set_option("variant", variant)
            s.cuda.set_option("precision", precision)
            s.cuda.set_option("mixed", band)
            x = torch.zeros((B, N, n), device=dev); u = torch.zeros((B, N - 1, m), device=dev)
            it = torch.zeros(B, dtype=torch.int32, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
            call = lambda: s.cuda.solve_batch_device(B, ptr(x0), ptr(Xr), ptr(Ur), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream.cuda_stream)
            call(); call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(a.reps):
                call()
            e1.record(stream)
            torch.cuda.synchronize()
            marked = s.cuda.last_marked if band > 0 else 0
            return dict(x=x.cpu().numpy(), u=u.cpu().numpy(), iter=it.cpu().numpy(), status=st.cpu().numpy(), ms=e0.elapsed_time(e1) / a.reps,
                        marked=marked, kernel=s.cuda.last_kernel)

        g = run(64, 0.0)
        for variant, band in [(int(v), float(bd)) for v in a.variants.split(",") for bd in a.bands.split(",")]:
            r = run(32, band, variant)
            bad = (r["iter"] != g["iter"]) | (r["status"] != g["status"])
            same = ~bad
            rec = dict(config=name, scale=scale, batch=B, variant=variant, band=band, marked=int(r["marked"]), marked_frac=r["marked"] / B,
                       mismatches=int(bad.sum()), mismatch_frac=float(bad.mean()), ms=r["ms"], ms_fp64=g["ms"],
                       solves_per_s=B / (r["ms"] * 1e-3), max_dx_on_equal=float(np.abs(r["x"][same] - g["x"][same]).max()),
                       max_du_on_equal=float(np.abs(r["u"][same] - g["u"][same]).max()), kernel=r["kernel"])
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
