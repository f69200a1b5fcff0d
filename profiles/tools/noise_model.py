#!/usr/bin/env python3
"""CPU experiment behind DESIGN.md section "fp32 noise": a numpy restatement of the box-constrained ADMM iteration
(admm.cpp:274-389) in float32, in two algebraically identical forms, against the fp64 reference's iteration counts.

  direct : x(k+1) = LQR(w(k)) evaluated from scratch every iteration (what admm.cpp does; costates p ~ Pinf x ~ 1e3..1e4
           carry fp32 rounding of ~5e-4 that lands in u and x at every iteration as FRESH noise)
  delta  : x(k+1) = x(k) + LQR_homogeneous(w(k) - w(k-1)); the Riccati sweeps run on increments whose size shrinks with
           the residuals, so the fresh noise is proportional to the residual itself

Test infrastructure (it calls the oracle); not part of the product.  Usage: python profiles/tools/noise_model.py [B]"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle as O  # noqa: E402

P = importlib.import_module("tinympc-matlab_b200.problems")


def admm(p, cache, b, dt, form):
    n, m, N = p.nx, p.nu, p.N
    f = lambda a: np.asarray(a, dt)
    A, Bm, K, Pinf, Qi, AK = f(p.A), f(p.B), f(cache["Kinf"]), f(cache["Pinf"]), f(cache["Quu_inv"]), f(cache["AmBKt"])
    rho = dt(p.rho)
    Qd, Rd = f(p.Qdiag + p.rho), f(p.Rdiag + p.rho)
    xmin, xmax, umin, umax = f(p.x_min), f(p.x_max), f(p.u_min), f(p.u_max)
    Bn = b.size
    x0 = f(b.x0)
    Xref = f(b.Xref) if b.Xref is not None else np.zeros((Bn, N, n), dt)
    Uref = f(b.Uref) if b.Uref is not None else np.zeros((Bn, N - 1, m), dt)
    x = np.zeros((Bn, N, n), dt); u = np.zeros((Bn, N - 1, m), dt)
    g = np.zeros_like(x); y = np.zeros_like(u); v = np.zeros_like(x); z = np.zeros_like(u)
    x[:, 0] = x0
    q = np.zeros_like(x); r = np.zeros_like(u); pN = np.zeros((Bn, n), dt)
    wq_prev = np.zeros_like(x); wr_prev = np.zeros_like(u)   # previous (q, r, pN) for the delta form
    pN_prev = np.zeros((Bn, n), dt)
    it = np.zeros(Bn, np.int32); st = np.full(Bn, 11, np.int32); done = np.zeros(Bn, bool)

    def sweeps(q, r, pN, x_init):
        d = np.zeros((Bn, N - 1, m), dt)
        pv = pN.copy()
        for i in range(N - 2, -1, -1):
            d[:, i] = (pv @ Bm + r[:, i]) @ Qi.T
            pv = q[:, i] + pv @ AK.T - r[:, i] @ K
        xs = np.zeros((Bn, N, n), dt); us = np.zeros((Bn, N - 1, m), dt)
        xs[:, 0] = x_init
        for i in range(N - 1):
            us[:, i] = -(xs[:, i] @ K.T) - d[:, i]
            xs[:, i + 1] = xs[:, i] @ A.T + us[:, i] @ Bm.T
        return xs, us

    for k in range(p.max_iter):
        if form == "direct" or k == 0:
            xs, us = sweeps(q, r, pN, x0)
            x, u = xs, us
        else:
            dx, du = sweeps(q - wq_prev, r - wr_prev, pN - pN_prev, np.zeros((Bn, n), dt))
            x = x + dx; u = u + du
        wq_prev, wr_prev, pN_prev = q, r, pN
        vn = np.minimum(xmax, np.maximum(xmin, x + g)); zn = np.minimum(umax, np.maximum(umin, u + y))
        g = g + x - vn; y = y + u - zn
        q = -(Xref * Qd) - rho * (vn - g); r = -(Uref * Rd) - rho * (zn - y)
        pN = -(Xref[:, N - 1] @ Pinf) - rho * (vn[:, N - 1] - g[:, N - 1])
        q[:, N - 1] = 0
        px = np.abs(x - vn).max(axis=(1, 2)); dx_ = rho * np.abs(v - vn).max(axis=(1, 2))
        pu = np.abs(u - zn).max(axis=(1, 2)); du_ = rho * np.abs(z - zn).max(axis=(1, 2))
        ok = (px < p.abs_pri_tol) & (pu < p.abs_pri_tol) & (dx_ < p.abs_dua_tol) & (du_ < p.abs_dua_tol)
        newly = ok & ~done
        it[~done] = k + 1
        st[newly] = 1
        done |= ok
        v, z = vn, zn
        if done.all():
            break
    return it, st


def main():
    Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    for name, mk, scale in (("quadrotor", P.quadrotor, 1.0), ("quadrotor", P.quadrotor, 0.3), ("cartpole", P.cartpole, 1.0)):
        p = mk()
        b = P.make_batch(p, Bn, scale, seed=99)
        cache = O.get_cache(p, "ref" if O.available("ref") else "port")
        gold = O.solve_batch(p, b, "ref" if O.available("ref") else "port")
        for dt, form in ((np.float64, "direct"), (np.float64, "delta"), (np.float32, "direct"), (np.float32, "delta")):
            it, st = admm(p, cache, b, dt, form)
            bad = (it != gold["iter"]) | (st != gold["status"])
            print(f"{name} s={scale} {np.dtype(dt).name:8s} {form:7s}: {int(bad.sum())}/{Bn} count/status mismatches vs reference")


if __name__ == "__main__":
    main()
