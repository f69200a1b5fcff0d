#!/usr/bin/env python3
"""Small batches of every dispatch path, meant to run under `compute-sanitizer --tool memcheck` (and racecheck / initcheck):
hybrid and plain tpp3 instances (box, per-problem bounds, reference-free, cones), tpp2 fp64 / adaptive, the general kernel, sessions."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
def run(name, spec, B, **opt):
    b = P.make_batch(spec, B, 1.0, seed=3)
    s = tm.TinyMPC().setup_from_spec(spec, devices=[0])
    for k, v in opt.items(): s.cuda.set_option(k, v)
    r = s.cuda.solve_batch(b.x0, b.Xref, b.Uref)
    print(f"{name:28s} B={B:6d} kernel={s.cuda.last_kernel} mean_iters={r['iter'].mean():.2f}", flush=True)
q, c, ro, qa = P.quadrotor(), P.cartpole(), P.rocket(), P.quadrotor(adaptive=True)
run("quadrotor plain (small)", q, 2048)
run("quadrotor hybrid", q, 60000)
run("quadrotor direct", q, 4096, variant=5)
run("quadrotor fp64", q, 4096, precision=64)
run("quadrotor mixed", q, 60000, mixed=0.003)
run("cartpole plain (small)", c, 2048)
run("cartpole hybrid", c, 80000)
run("rocket cones", ro, 8192)
run("adaptive", qa, 4096)
run("general kernel f32", q, 512, force_wpp=1)
run("general kernel f64", ro, 512, force_wpp=1, precision=64)
s = tm.TinyMPC().setup_from_spec(q, devices=[0]); ses = s.cuda.session(256)
b = P.make_batch(q, 256, 0.3, seed=4); ses.set_x_ref(b.Xref.astype(np.float64)); ses.set_x0(b.x0.astype(np.float64))
for _ in range(3): ses.solve(); ses.step()
print("session ok", ses.read("iter").mean())
