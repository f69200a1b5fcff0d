#!/usr/bin/env python3
"""Small invocations of every mode of the lane-group kernel (tmpc_gpp.cuh) and of the impulse-response fp32 kernel, meant to run
under `compute-sanitizer --tool memcheck / racecheck / initcheck`: fp64 batches of the three shapes (ragged sizes), adaptive rho,
the exact-count mode in its sequential and concurrent forms, sessions, tiny_solve."""
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
    print(f"{name:34s} B={B:6d} kernel={s.cuda.last_kernel} mean_iters={r['iter'].mean():.2f}", flush=True)
q, c, qa = P.quadrotor(), P.cartpole(), P.quadrotor(adaptive=True)
rb = P.rocket(linear=False); rb.en_state_soc = rb.en_input_soc = 0
rb.Acx = rb.qcx = rb.Acu = rb.qcu = np.zeros(0, np.int32); rb.cx = rb.cu = np.zeros(0); rb.name = "rocket_box"
run("gpp quadrotor fp64", q, 1237, precision=64)
run("gpp cartpole fp64", c, 1001, precision=64)
run("gpp rocket-shaped box fp64", rb, 515, precision=64)
run("gpp adaptive fp64", qa, 777, precision=64)
run("tpp3 impulse-response plain", q, 2048)
run("tpp3 impulse-response hybrid", q, 60000)
run("exact-count sequential", q, 60000, mixed=0.002, fixer_sms=-1)
run("exact-count concurrent", q, 60000, mixed=0.002, fixer_sms=0)
s = tm.TinyMPC().setup_from_spec(q, devices=[0]); s.cuda.set_option("precision", 64); ses = s.cuda.session(300)
b = P.make_batch(q, 300, 0.3, seed=4); ses.set_x_ref(b.Xref.astype(np.float64)); ses.set_x0(b.x0.astype(np.float64))
for _ in range(3): ses.solve(); ses.step()
print("gpp session ok", s.cuda.last_kernel, ses.read("iter").mean(), flush=True)
p1 = P.cartpole(matlab_defaults=True)
t = tm.TinyMPC().setup_from_spec(p1, devices=[0]); t.set_x0([0.5, 0, 0, 0]); t.solve(); t.solve()
print("tiny_solve ok", t.cuda.last_kernel, t.get_stats()["iter"], flush=True)
