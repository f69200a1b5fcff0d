#!/usr/bin/env python3
"""Where does the fp32 error of one Riccati backward+forward pass (admm.cpp:13-32) come from?  numpy experiment."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle as O
P = importlib.import_module("tinympc-matlab_b200.problems")

def run(p, cache, b, dt, mode="std", D=None):
    n, m, N = p.nx, p.nu, p.N
    f = lambda a: np.asarray(a, dt)
    A, Bm, K, Pinf, Qi, AK = p.A, p.B, cache["Kinf"], cache["Pinf"], cache["Quu_inv"], cache["AmBKt"]
    if D is not None:   # x~ = D^-1 x
        Di = 1.0 / D
        A = Di[:, None] * A * D[None, :]; Bm = Di[:, None] * Bm; K = K * D[None, :]; Pinf = D[:, None] * Pinf * D[None, :]
        AK = D[:, None] * AK * Di[None, :]
    A, Bm, K, Pinf, Qi, AK = map(f, (A, Bm, K, Pinf, Qi, AK))
    Bn = b.size
    g = O.solve_batch(p, b, "ref")
    # linear cost of the converged point: q = -Xref.*Q - rho*(v - g) ~ use v = x*, g = 0 as a representative w
    Qd, Rd = p.Qdiag + p.rho, p.Rdiag + p.rho
    xs, us = g["x"], g["u"]
    q = -(b.Xref.astype(np.float64) * Qd) - p.rho * xs
    r = -p.rho * us
    pN = -(b.Xref[:, N - 1].astype(np.float64) @ cache["Pinf"]) - p.rho * xs[:, N - 1]
    x0 = b.x0.astype(np.float64)
    if D is not None:
        q = q * D; pN = pN * D; x0 = x0 / D
    q, r, pN, x0 = map(f, (q, r, pN, x0))
    d = np.zeros((Bn, N - 1, m), dt); pv = pN.copy()
    for i in range(N - 2, -1, -1):
        d[:, i] = (pv @ Bm + r[:, i]) @ Qi.T
        pv = q[:, i] + pv @ AK.T - r[:, i] @ K
    x = np.zeros((Bn, N, n), dt); u = np.zeros((Bn, N - 1, m), dt); x[:, 0] = x0
    for i in range(N - 1):
        u[:, i] = -(x[:, i] @ K.T) - d[:, i]
        if mode == "cl":
            x[:, i + 1] = x[:, i] @ AK - d[:, i] @ Bm.T
        else:
            x[:, i + 1] = x[:, i] @ A.T + u[:, i] @ Bm.T
    if D is not None:
        x = x * D
    return d.astype(np.float64), x.astype(np.float64), u.astype(np.float64)

p = P.quadrotor(); b = P.make_batch(p, 2000, 1.0, seed=5)
cache = O.get_cache(p, "ref")
d64, x64, u64 = run(p, cache, b, np.float64)
for name, kw in (("std", {}), ("closed-loop fwd", dict(mode="cl"))):
    d32, x32, u32 = run(p, cache, b, np.float32, **kw)
    print(f"{name:18s} max|dd|={np.abs(d32-d64).max():.2e} max|dx|={np.abs(x32-x64).max():.2e} max|du|={np.abs(u32-u64).max():.2e}   per-step du: {np.abs(u32-u64).max(axis=(0,2))}")
# balancing: D from sqrt(diag(Pinf))^-1
Dg = 1.0 / np.sqrt(np.diag(cache["Pinf"])); Dg = Dg / Dg.max()
d32, x32, u32 = run(p, cache, b, np.float32, D=Dg)
d64b, x64b, u64b = run(p, cache, b, np.float64, D=Dg)
print("balanced check fp64:", np.abs(x64b - x64).max(), np.abs(u64b-u64).max())
print(f"balanced           max|dd|={np.abs(d32-d64).max():.2e} max|dx|={np.abs(x32-x64).max():.2e} max|du|={np.abs(u32-u64).max():.2e}")
print("Pinf diag", np.diag(cache["Pinf"]).round(1)); print("AK absmax rows", np.abs(cache["AmBKt"]).max(axis=1).round(2))
print("|K|", np.abs(cache["Kinf"]).max(axis=0).round(3)); print("B max", np.abs(p.B).max(axis=1).round(4))
