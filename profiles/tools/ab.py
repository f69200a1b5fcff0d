#!/usr/bin/env python3
"""A/B of kernel variants: python profiles/tools/ab.py quadrotor:0,7 cartpole:0,7 rocket:0,7 [--B 1048576]"""
import sys
sys.argv[0:1] = [sys.argv[0]]
from variant_sweep import run, P  # noqa
B = 1 << 20
args = [a for a in sys.argv[1:]]
if "--B" in args:
    i = args.index("--B"); B = int(args[i + 1]); del args[i:i + 2]
specs = dict(quadrotor=lambda: P.quadrotor(), cartpole=lambda: P.cartpole(), rocket=lambda: P.rocket(), quadrotor_adaptive=lambda: P.quadrotor(adaptive=True),
             cartpole10=lambda: P.cartpole(N=10))
for a in args:
    name, vs = a.split(":")
    scale = 1.0
    if "@" in name:
        name, sc = name.split("@"); scale = float(sc)
    run(specs[name](), scale, B, [int(v) for v in vs.split(",")], reps=5)
