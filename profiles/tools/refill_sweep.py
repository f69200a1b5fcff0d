#!/usr/bin/env python3
"""Sweep of the batched-refill threshold (option refill_min) over the BASELINE configs."""
import sys
from variant_sweep import run, P  # noqa
B = 1 << 20
for name, spec, scale in (("quadrotor", P.quadrotor(), 1.0), ("quadrotor", P.quadrotor(), 0.3), ("cartpole", P.cartpole(), 1.0), ("rocket", P.rocket(), 1.0)):
    for m in (1, 3, 8, 0):
        run(spec, scale, B, [0], reps=4, options={"refill_min": m})
