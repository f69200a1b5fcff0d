#!/usr/bin/env python3
"""CPU experiment behind the incremental-form kernel with cones and half-spaces (tmpc_tpp3.cuh, CONSTR): a numpy restatement
of the ADMM iteration with box + second-order-cone + linear-inequality slacks (admm.cpp:81-247) in float32, in the direct
form (every sweep from scratch) and in the increment form the kernel uses (stored pre-projection values t per family, the
increment of (slack - dual) taken as a difference of stored values), against the fp64 reference's iteration counts.

Test infrastructure (it calls the oracle); not part of the product.  Usage: python profiles/tools/noise_model_con.py [B]"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle as O  # noqa: E402

P = importlib.import_module("tinympc-matlab_b200.problems")


def proj_soc(t, start, dim, mu, dt):
    """admm.cpp:39-60 on t[..., start:start+dim] (float mu, float norm)."""
    t = t.copy()
    blk = t[..., start:start + dim]
    last = blk[..., dim - 1]
    u0 = last * dt(np.float32(mu))
    a = np.sqrt((blk[..., :dim - 1] ** 2).sum(-1)).astype(np.float32).astype(dt)
    zero = a <= -u0
    inside = (~zero) & (a <= u0)
    with np.errstate(divide="ignore", invalid="ignore"):
        fct = dt(0.5) * (dt(1) + u0 / a)
    out = blk.copy()
    scale = np.where(zero, dt(0), np.where(inside, dt(1), fct))
    out[..., :dim - 1] = blk[..., :dim - 1] * scale[..., None]
    lastn = np.where(zero, dt(0), np.where(inside, last, fct * (a.astype(np.float32) / np.float32(mu)).astype(dt)))
    out[..., dim - 1] = lastn
    t[..., start:start + dim] = out
    return t


def proj_lin(t, Al, bl, dt):
    """admm.cpp:70-73 applied row after row in place (:148-159)."""
    t = t.copy()
    for c in range(Al.shape[0]):
        a = Al[c].astype(dt)
        val = (t * a).sum(-1)
        nrm = dt((a * a).sum())
        dist = np.where(val > dt(bl[c]), (val - dt(bl[c])) / nrm, dt(0))
        t = t - dist[..., None] * a
    return t


def admm(p, cache, b, dt, form):
    n, m, N = p.nx, p.nu, p.N
    f = lambda a: np.asarray(a, dt)
    A, Bm, K, Pinf, Qi, AK = f(p.A), f(np.asarray(p.B).reshape(n, m)), f(cache["Kinf"]), f(cache["Pinf"]), f(cache["Quu_inv"]), f(cache["AmBKt"])
    APf, BPf, fd = f(cache["APf"]).ravel(), f(cache["BPf"]).ravel(), f(p.f).ravel()
    rho = dt(p.rho)
    Qd, Rd = f(p.Qdiag + p.rho), f(p.Rdiag + p.rho)
    xmin, xmax, umin, umax = f(p.x_min), f(p.x_max), f(p.u_min), f(p.u_max)
    Bn = b.size
    x0 = f(b.x0)
    Xref = f(b.Xref); Uref = f(b.Uref)
    it = np.zeros(Bn, np.int32); st = np.full(Bn, 11, np.int32); done = np.zeros(Bn, bool)

    def sweeps(q, r, pN, x_init, affine):
        d = np.zeros((Bn, N - 1, m), dt)
        pv = pN.copy()
        for i in range(N - 2, -1, -1):
            d[:, i] = (pv @ Bm + r[:, i] + (BPf if affine else 0)) @ Qi.T
            pv = q[:, i] + pv @ AK.T - r[:, i] @ K + (APf if affine else 0)
        xs = np.zeros((Bn, N, n), dt); us = np.zeros((Bn, N - 1, m), dt)
        xs[:, 0] = x_init
        for i in range(N - 1):
            us[:, i] = -(xs[:, i] @ K.T) - d[:, i]
            xs[:, i + 1] = xs[:, i] @ A.T + us[:, i] @ Bm.T + (fd if affine else 0)
        return xs, us

    def project(fam, t, state):
        if fam == "box":
            return np.minimum(xmax, np.maximum(xmin, t)) if state else np.minimum(umax, np.maximum(umin, t))
        if fam == "soc":
            A_, q_, c_ = (p.Acx, p.qcx, p.cx) if state else (p.Acu, p.qcu, p.cu)
            for c in range(len(A_)):
                t = proj_soc(t, int(A_[c]), int(q_[c]), float(c_[c]), dt)
            return t
        Al, bl = (np.asarray(p.Alin_x), np.asarray(p.blin_x)) if state else (np.asarray(p.Alin_u), np.asarray(p.blin_u))
        return proj_lin(t, Al, bl, dt)

    fams_x = ["box"] + (["soc"] if p.en_state_soc and len(p.Acx) else []) + (["lin"] if p.en_state_linear else [])
    fams_u = ["box"] + (["soc"] if p.en_input_soc and len(p.Acu) else []) + (["lin"] if p.en_input_linear else [])
    # stored pre-projection values t = x + dual_prev per family (cold start: 0)
    tx = {fm: np.zeros((Bn, N, n), dt) for fm in fams_x}
    tu = {fm: np.zeros((Bn, N - 1, m), dt) for fm in fams_u}
    q = np.zeros((Bn, N, n), dt); r = np.zeros((Bn, N - 1, m), dt); pN = np.zeros((Bn, n), dt)
    x = u = None
    for k in range(p.max_iter):
        if k == 0:
            x, u = sweeps(q, r, pN, x0, True)
        elif form == "direct":
            x, u = sweeps(q, r, pN, x0, True)
        else:
            dxs, dus = sweeps(dq, dr, dpN, np.zeros((Bn, n), dt), False)
            x = x + dxs; u = u + dus
        first = k == 0
        dwx = np.zeros((Bn, N, n), dt); dwu = np.zeros((Bn, N - 1, m), dt)
        wx = np.zeros((Bn, N, n), dt); wu = np.zeros((Bn, N - 1, m), dt)
        for fm in fams_x:
            to = tx[fm]
            vo = np.zeros_like(to) if first else project(fm, to, True)
            tn = x + (to - vo)
            vn = project(fm, tn, True)
            if fm == "box":
                px = np.abs(x - vn).max(axis=(1, 2)); dxr = rho * np.abs(vn - vo).max(axis=(1, 2))
            dwx += dt(2) * (vn - vo) - (tn - to)
            wx += dt(2) * vn - tn
            tx[fm] = tn
        for fm in fams_u:
            to = tu[fm]
            vo = np.zeros_like(to) if first else project(fm, to, False)
            tn = u + (to - vo)
            vn = project(fm, tn, False)
            if fm == "box":
                pu = np.abs(u - vn).max(axis=(1, 2)); dur = rho * np.abs(vn - vo).max(axis=(1, 2))
            dwu += dt(2) * (vn - vo) - (tn - to)
            wu += dt(2) * vn - tn
            tu[fm] = tn
        if form == "direct":
            q = -(Xref * Qd) - rho * wx; r = -(Uref * Rd) - rho * wu
            pN = -(Xref[:, N - 1] @ Pinf) - rho * wx[:, N - 1]
            q[:, N - 1] = 0
        else:
            dq = -rho * dwx; dr = -rho * dwu
            if first:
                dq = dq - Xref * Qd; dr = dr - Uref * Rd
                dpN = -(Xref[:, N - 1] @ Pinf) - rho * dwx[:, N - 1]
            else:
                dpN = dq[:, N - 1].copy()
            dq[:, N - 1] = 0
            if first:   # the direct quantities, for the k = 1 full sweep of the direct form only
                pass
        ok = (px < p.abs_pri_tol) & (pu < p.abs_pri_tol) & (dxr < p.abs_dua_tol) & (dur < p.abs_dua_tol)
        newly = ok & ~done
        it[~done] = k + 1
        st[newly] = 1
        done |= ok
        if done.all():
            break
    return it, st


def main():
    Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    for name, mk in (("rocket", P.rocket), ("rocket_nolinear", lambda: P.rocket(linear=False))):
        p = mk()
        b = P.make_batch(p, Bn, 1.0, seed=99)
        impl = "ref" if O.available("ref") else "port"
        cache = O.get_cache(p, impl)
        gold = O.solve_batch(p, b, impl)
        for dt, form in ((np.float64, "direct"), (np.float64, "delta"), (np.float32, "direct"), (np.float32, "delta")):
            it, st = admm(p, cache, b, dt, form)
            bad = (it != gold["iter"]) | (st != gold["status"])
            print(f"{name} {np.dtype(dt).name:8s} {form:7s}: {int(bad.sum())}/{Bn} count/status mismatches vs reference "
                  f"(mean iters {it.mean():.2f} vs {gold['iter'].mean():.2f})", flush=True)


if __name__ == "__main__":
    main()
