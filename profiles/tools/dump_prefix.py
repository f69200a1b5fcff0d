#!/usr/bin/env python3
"""Solve the parity prefix of a bench config on the GPU in a given mode and dump x, u, iter, status (+ the reference's) to
gpurun_out/<tag>.npz for offline error analysis.  Test infrastructure (calls the oracle).
Usage: python profiles/tools/dump_prefix.py <config> <n> <band> <tag> [variant] [precision] [nsolve] [fixer_sms]"""
import importlib, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O
tm = importlib.import_module("tinympc-matlab_b200")
P = importlib.import_module("tinympc-matlab_b200.problems")
cfg, n, band, tag = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
prec = int(sys.argv[6]) if len(sys.argv) > 6 else 32
nsolve = int(sys.argv[7]) if len(sys.argv) > 7 else n   # solve a larger batch (other kernel instance), compare its first n problems
fix = int(sys.argv[8]) if len(sys.argv) > 8 else 0
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole, rocket=P.rocket, quadrotor_adaptive=lambda: P.quadrotor(adaptive=True))[cfg]()
bb = P.make_batch(spec, nsolve, 1.0, seed=1234 + 3)
b = bb.slice(0, n)
s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[0])
s.cuda.set_option("precision", prec); s.cuda.set_option("variant", variant); s.cuda.set_option("mixed", band); s.cuda.set_option("fixer_sms", fix)
r = s.cuda.solve_batch(bb.x0, bb.Xref, bb.Uref)
r = {k: v[:n] for k, v in r.items() if v is not None and hasattr(v, "shape")}
g = O.solve_batch(spec, b, "ref" if O.available("ref") else "port", os.cpu_count() or 1)
same = (r["iter"] == g["iter"]) & (r["status"] == g["status"])
dx = np.abs(r["x"] - g["x"]).reshape(n, -1); du = np.abs(r["u"] - g["u"]).reshape(n, -1)
print(f"{cfg} n={n} band={band} variant={variant} kernel={s.cuda.last_kernel} marked={s.cuda.last_marked}: mismatches {int((~same).sum())}, "
      f"max|dx| {dx.max():.3e} max|du| {du.max():.3e}; matched only: {dx[same].max():.3e} {du[same].max():.3e}")
w = np.argsort(-du.max(1))[:8]
for k in w:
    e = int(du[k].argmax())
    print(f"  problem {k}: iter {r['iter'][k]}/{g['iter'][k]} status {r['status'][k]}/{g['status'][k]} max|du| {du[k].max():.3e} at step {e // spec.nu} elem {e % spec.nu}: "
          f"{r['u'].reshape(n, -1)[k, e]:.7f} vs {g['u'].reshape(n, -1)[k, e]:.7f}; max|dx| {dx[k].max():.3e}")
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
np.savez_compressed(ROOT / "gpurun_out" / f"{tag}.npz", x=r["x"], u=r["u"], iter=r["iter"], status=r["status"], gx=g["x"], gu=g["u"], giter=g["iter"], gstatus=g["status"])
