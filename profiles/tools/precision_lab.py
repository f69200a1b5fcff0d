#!/usr/bin/env python3
"""CPU experiment for the round-2 precision work on the incremental-form kernel (tmpc_tpp3.cuh): which representation of the
per-problem state makes the fp32 termination decisions agree with the fp64 reference often enough that the exact-count mode
("mixed": fp64 re-solve of the problems within a relative band of a tolerance) needs only a narrow band.

A numpy restatement of the incremental ADMM iteration (box + optional cone + half-space families, admm.cpp:81-271) with the
arithmetic of the increments in float32 and a choice of how the ACCUMULATED state (x, u and the pre-projection slacks t) is
held and how the dual residual is formed:

  base   : x, t in float32; dual residual = difference of the stored projections (what round 1 shipped)
  ince   : as base, but the dual residual element is the increment  delta = dx + (x_old - v_old)  wherever the box is inactive
           before and after (no representation noise of the stored values in the test)
  comp   : x, u, t held to double-float accuracy (emulated with float64 storage); increments still float32
  comp+ince
  +conv  : the backward pass in impulse-response form, d_i = Quu_inv r_i + sum_{j>i} G_{j-i-1} s_j (what tmpc_tpp3.cuh ships for the
           quadrotor shapes), instead of the costate recursion; +p64 / +p64q: the recursion (and Quu_inv B' p) carried in float64;
           +f64w: the rollout in float64; <variant>:K : the first K iterations' sweeps entirely in float64

For every variant: iteration-count mismatches against the fp64 run of the same model, and -- the quantity that sizes the band --
for every problem the largest relative disagreement |r32 - r64| / tol of a termination residual over the iterations both runs
executed (a band wider than that, applied around the tolerance, catches every possible flip).

Test infrastructure (it calls the oracle); not part of the product.  Usage: python profiles/tools/precision_lab.py [config] [B]"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle as O  # noqa: E402
from noise_model_con import proj_soc, proj_lin  # noqa: E402

P = importlib.import_module("tinympc-matlab_b200.problems")
f32 = np.float32


def admm(p, cache, b, dt, variant="base", kmax=None):
    """dt: arithmetic type of the increments (float32 / float64).  Returns it, st, residual trace (K, B, 4), x, u."""
    KF = int(variant.split(":")[1]) if ":" in variant else 0
    variant = variant.split(":")[0]
    P64 = 2 if "p64q" in variant else (1 if "p64" in variant else 0)
    F64 = "f64w" in variant
    CONV = "conv" in variant
    comp = "comp" in variant
    ince = "ince" in variant
    xdf = variant.startswith("xdf")            # x, u double-float; per-family DUALS stored in float32 (not the pre-projection t)
    p32 = variant in ("xdf32",)                 # cone / half-space projections evaluated in float32 on rounded inputs
    sdt = np.float64 if (comp or xdf) else dt          # storage type of the accumulated state
    # which families keep their dual in float32: xdf all, xdfb box only, xdfc box + cone, xdfl box + half-space
    f32dual = lambda fm: {"xdf": True, "xdf32": True, "xdfb": fm == "box", "xdfc": fm in ("box", "soc"), "xdfl": fm in ("box", "lin")}[variant]
    n, m, N = p.nx, p.nu, p.N
    f = lambda a: np.asarray(a, dt)
    A, Bm, K, Pinf, Qi, AK = f(p.A), f(np.asarray(p.B).reshape(n, m)), f(cache["Kinf"]), f(cache["Pinf"]), f(cache["Quu_inv"]), f(cache["AmBKt"])
    APf, BPf, fd = f(cache["APf"]).ravel(), f(cache["BPf"]).ravel(), f(p.f).ravel()
    rho = dt(p.rho)
    Qd, Rd = f(p.Qdiag + p.rho), f(p.Rdiag + p.rho)
    xmin, xmax, umin, umax = (np.asarray(a, sdt) for a in (p.x_min, p.x_max, p.u_min, p.u_max))
    Bn = b.size
    x0 = f(b.x0)
    Xref = f(b.Xref) if b.Xref is not None else np.zeros((Bn, N, n), dt)
    Uref = f(b.Uref) if b.Uref is not None else np.zeros((Bn, N - 1, m), dt)
    it = np.zeros(Bn, np.int32); st = np.full(Bn, 11, np.int32); done = np.zeros(Bn, bool)
    kmax = kmax or p.max_iter
    trace = np.full((kmax, Bn, 4), np.nan)

    D_ = np.float64
    B64, Qi64, AK64, K64 = (np.asarray(a, D_) for a in (np.asarray(p.B).reshape(n, m), cache["Quu_inv"], cache["AmBKt"], cache["Kinf"]))
    APf64, BPf64 = np.asarray(cache["APf"], D_).ravel(), np.asarray(cache["BPf"], D_).ravel()
    A64 = np.asarray(p.A, D_); fd64 = np.asarray(p.f, D_).ravel()
    G64 = [Qi64 @ B64.T]
    for _k in range(N - 2): G64.append(G64[-1] @ AK64)
    G32 = [np.asarray(g_, dt) for g_ in G64]
    def sweeps(q, r, pN, x_init, affine):
        d = np.zeros((Bn, N - 1, m), dt)
        if CONV:
            # impulse-response form of the backward pass, float32: dd_i = Quu_inv r_i + sum_{j>i} G_{j-i-1} s_j, s_j = q_j - K' r_j, s_{N-1} = p_N
            s_ = np.zeros((Bn, N, n), dt)
            for j in range(N - 1): s_[:, j] = q[:, j] - r[:, j] @ K
            s_[:, N - 1] = pN
            acc = np.zeros((Bn, N - 1, m), dt)
            for j in range(N - 1, 0, -1):
                for k in range(j):
                    acc[:, j - 1 - k] = acc[:, j - 1 - k] + s_[:, j] @ G32[k].T
            for i in range(N - 1):
                d[:, i] = acc[:, i] + (r[:, i] + (BPf if affine else 0)) @ Qi.T
            if affine and np.abs(APf64).max() > 0: raise SystemExit('conv: affine APf not modelled')
        elif P64:
            pv = pN.astype(D_)
            for i in range(N - 2, -1, -1):
                r64 = r[:, i].astype(D_)
                bp = pv @ B64                                  # B' p in double
                if P64 == 1:
                    d[:, i] = ((bp + r64 + (BPf64 if affine else 0)).astype(dt)) @ Qi.T       # Quu_inv product in float32
                else:
                    d[:, i] = ((bp + r64 + (BPf64 if affine else 0)) @ Qi64.T).astype(dt)
                pv = q[:, i].astype(D_) + pv @ AK64.T - r64 @ K64 + (APf64 if affine else 0)
        else:
          pv = pN.copy()
          for i in range(N - 2, -1, -1):
            d[:, i] = (pv @ Bm + r[:, i] + (BPf if affine else 0)) @ Qi.T
            pv = q[:, i] + pv @ AK.T - r[:, i] @ K + (APf if affine else 0)
        if F64:
            xs = np.zeros((Bn, N, n), D_); us = np.zeros((Bn, N - 1, m), D_)
            xs[:, 0] = x_init
            for i in range(N - 1):
                us[:, i] = -(xs[:, i] @ K64.T) - d[:, i]
                xs[:, i + 1] = xs[:, i] @ A64.T + us[:, i] @ B64.T + (fd64 if affine else 0)
            return xs.astype(dt), us.astype(dt)
        xs = np.zeros((Bn, N, n), dt); us = np.zeros((Bn, N - 1, m), dt)
        xs[:, 0] = x_init
        for i in range(N - 1):
            us[:, i] = -(xs[:, i] @ K.T) - d[:, i]
            xs[:, i + 1] = xs[:, i] @ A.T + us[:, i] @ Bm.T + (fd if affine else 0)
        return xs, us

    def sweeps64(q, r, pN, x_init, affine):
        D=np.float64
        d = np.zeros((Bn, N - 1, m), D)
        pv = pN.copy()
        for i in range(N - 2, -1, -1):
            d[:, i] = (pv @ Bm + r[:, i] + (BPf if affine else 0)) @ Qi.T
            pv = q[:, i] + pv @ AK.T - r[:, i] @ K + (APf if affine else 0)
        xs = np.zeros((Bn, N, n), D); us = np.zeros((Bn, N - 1, m), D)
        xs[:, 0] = x_init
        for i in range(N - 1):
            us[:, i] = -(xs[:, i] @ K.T) - d[:, i]
            xs[:, i + 1] = xs[:, i] @ A.T + us[:, i] @ Bm.T + (fd if affine else 0)
        return xs, us

    def project(fam, t, state):
        tdt = t.dtype.type
        if fam == "box":
            return np.minimum(xmax, np.maximum(xmin, t)) if state else np.minimum(umax, np.maximum(umin, t))
        if fam == "soc":
            A_, q_, c_ = (p.Acx, p.qcx, p.cx) if state else (p.Acu, p.qcu, p.cu)
            for c in range(len(A_)):
                t = proj_soc(t, int(A_[c]), int(q_[c]), float(c_[c]), tdt)
            return t
        Al, bl = (np.asarray(p.Alin_x), np.asarray(p.blin_x)) if state else (np.asarray(p.Alin_u), np.asarray(p.blin_u))
        return proj_lin(t, Al, bl, tdt)

    fams_x = ["box"] + (["soc"] if p.en_state_soc and len(p.Acx) else []) + (["lin"] if p.en_state_linear else [])
    fams_u = ["box"] + (["soc"] if p.en_input_soc and len(p.Acu) else []) + (["lin"] if p.en_input_linear else [])
    tx = {fm: np.zeros((Bn, N, n), sdt) for fm in fams_x}
    tu = {fm: np.zeros((Bn, N - 1, m), sdt) for fm in fams_u}
    q = np.zeros((Bn, N, n), dt); r = np.zeros((Bn, N - 1, m), dt); pN = np.zeros((Bn, n), dt)
    x = np.zeros((Bn, N, n), sdt); u = np.zeros((Bn, N - 1, m), sdt)
    for k in range(kmax):
        if k < KF:
            A_, Bm_, K_, Qi_, AK_, APf_, BPf_, fd_ = A, Bm, K, Qi, AK, APf, BPf, fd
            g=lambda a: np.asarray(a, np.float64)
            A, Bm, K, Pinf, Qi, AK = g(p.A), g(np.asarray(p.B).reshape(n, m)), g(cache["Kinf"]), g(cache["Pinf"]), g(cache["Quu_inv"]), g(cache["AmBKt"])
            APf, BPf, fd = g(cache["APf"]).ravel(), g(cache["BPf"]).ravel(), g(p.f).ravel()
            dt_save=dt
            if k == 0:
                dxs, dus = sweeps64(g(q), g(r), g(pN), g(x0), True)
            else:
                dxs, dus = sweeps64(g(dq), g(dr), g(dpN), np.zeros((Bn, n)), False)
            dxs=dxs.astype(dt); dus=dus.astype(dt)
            A, Bm, K, Qi, AK, APf, BPf, fd = A_, Bm_, K_, Qi_, AK_, APf_, BPf_, fd_
            Pinf=f(cache["Pinf"])
        elif k == 0:
            dxs, dus = sweeps(q, r, pN, x0, True)
        else:
            dxs, dus = sweeps(dq, dr, dpN, np.zeros((Bn, n), dt), False)
        x_old, u_old = x, u
        x = x + dxs.astype(sdt); u = u + dus.astype(sdt)          # base: rounds to float32; comp: double-float accumulate
        first = k == 0
        dwx = np.zeros((Bn, N, n), dt); dwu = np.zeros((Bn, N - 1, m), dt)
        res = {}
        for state, fams, tt, xx, xx_old, dxx in ((True, fams_x, tx, x, x_old, dxs), (False, fams_u, tu, u, u_old, dus)):
            dw = dwx if state else dwu
            for fm in fams:
                if xdf:
                    # state: x (double-float) and the family's dual G (float32): t_old = x_old + G, v_old = proj(t_old)
                    G = tt[fm].astype(dt) if f32dual(fm) else tt[fm]   # stored as float32 (xdfb: only the box duals)
                    to = xx_old + G.astype(sdt)
                    if p32 and fm != "box":
                        to_r = to.astype(dt)
                        vo_r = np.zeros_like(to_r) if first else project(fm, to_r, state)
                        go = (to_r - vo_r).astype(dt)
                        tn_r = (xx + go.astype(sdt)).astype(dt)
                        vn_r = project(fm, tn_r, state)
                        dw += (dt(2) * (vn_r - vo_r) - (tn_r - to_r)).astype(dt)
                        tt[fm] = go.astype(sdt)
                        continue
                    vo = np.zeros_like(to) if first else project(fm, to, state)
                    go = (to - vo).astype(dt) if f32dual(fm) else (to - vo)
                    tn = xx + go.astype(sdt)
                    vn = project(fm, tn, state)
                    a = (xx - vn).astype(dt)
                    e = (vn - vo).astype(dt)
                    if fm == "box":
                        res["p" + ("x" if state else "u")] = np.abs(a).max(axis=(1, 2))
                        res["d" + ("x" if state else "u")] = rho * np.abs(e).max(axis=(1, 2))
                        dw += e - a
                    else:
                        # t_new - t_old = x_new - v_old exactly (t_new = x_new + (t_old - v_old)): formed from x and v, the rounding of the
                        # stored dual does not enter the increment of the linear cost
                        dw += (dt(2) * e - ((xx - vo) if "id" in variant or True else (tn - to)).astype(dt)).astype(dt)
                    tt[fm] = go.astype(sdt)
                    continue
                to = tt[fm]
                vo = np.zeros_like(to) if first else project(fm, to, state)
                tn = xx + (to - vo)
                vn = project(fm, tn, state)
                a = (xx - vn).astype(dt)
                e = (vn - vo).astype(dt)
                if ince and fm == "box" and not first:
                    delta = (dxx + (xx_old - vo).astype(dt)).astype(dt)
                    inside = (vo == to) & (vn == tn)
                    e = np.where(inside, delta, e)
                if fm == "box":
                    res["p" + ("x" if state else "u")] = np.abs(a).max(axis=(1, 2))
                    res["d" + ("x" if state else "u")] = rho * np.abs(e).max(axis=(1, 2))
                    dw += e - a
                else:
                    dw += (dt(2) * (vn - vo) - (tn - to)).astype(dt)
                tt[fm] = tn
        dq = -rho * dwx; dr = -rho * dwu
        if first:
            dq = dq - Xref * Qd; dr = dr - Uref * Rd
            dpN = -(Xref[:, N - 1] @ Pinf) - rho * dwx[:, N - 1]
        else:
            dpN = dq[:, N - 1].copy()
        dq[:, N - 1] = 0
        trace[k, :, 0], trace[k, :, 1], trace[k, :, 2], trace[k, :, 3] = res["px"], res["dx"], res["pu"], res["du"]
        ok = (res["px"] < p.abs_pri_tol) & (res["pu"] < p.abs_pri_tol) & (res["dx"] < p.abs_dua_tol) & (res["du"] < p.abs_dua_tol)
        newly = ok & ~done
        it[~done] = k + 1
        st[newly] = 1
        done |= ok
        if done.all():
            break
    admm.last_x, admm.last_u = np.asarray(x, np.float64), np.asarray(u, np.float64)
    return it, st, trace


def band_needed(p, it32, tr32, it64, tr64):
    """per problem: max over the iterations both runs executed of the relative residual disagreement, counted only where it can
    matter (the fp64 residual within a factor 2 of its tolerance)"""
    tol = np.array([p.abs_pri_tol, p.abs_dua_tol, p.abs_pri_tol, p.abs_dua_tol])
    K = tr64.shape[0]
    kk = np.arange(K)[:, None]
    live = kk < np.minimum(it32, it64)[None, :]
    rel = np.abs(tr32 - tr64) / tol
    near = (tr64 > 0.5 * tol) & (tr64 < 2 * tol)
    rel = np.where(live[:, :, None] & near, rel, 0.0)
    return np.nanmax(rel, axis=(0, 2))


def at_risk(p, it64, tr64, band):
    """fraction of the problems the fp32 pass would hand to fp64 with this band: some executed check where all four residuals are
    below (1 + band) tol but not all below (1 - band) tol"""
    tol = np.array([p.abs_pri_tol, p.abs_dua_tol, p.abs_pri_tol, p.abs_dua_tol])
    K = tr64.shape[0]
    live = np.arange(K)[:, None] < it64[None, :]
    up = (tr64 < tol * (1 + band)).all(-1)
    dn = (tr64 < tol * (1 - band)).all(-1)
    return float(((up & ~dn) & live).any(0).mean())


def mixed_outcome(p, it32, tr32, it64, band):
    """what the exact-count mode would do with this fp32 run: a problem is handed to fp64 at its first executed check whose four
    residuals are all below (1 + band) tol but not all below (1 - band) tol (tmpc_tpp3.cuh, amb_band); returns (fraction
    handed over, count mismatches left among the others)"""
    tol = np.array([p.abs_pri_tol, p.abs_dua_tol, p.abs_pri_tol, p.abs_dua_tol])
    K = tr32.shape[0]
    live = np.arange(K)[:, None] < it32[None, :]
    up = (tr32 < tol * (1 + band)).all(-1)
    dn = (tr32 < tol * (1 - band)).all(-1)
    marked = ((up & ~dn) & live).any(0)
    return float(marked.mean()), int(((it32 != it64) & ~marked).sum())


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "rocket"
    Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    p = dict(rocket=P.rocket, quadrotor=P.quadrotor, cartpole=P.cartpole, rocket_nolinear=lambda: P.rocket(linear=False))[cfg]()
    b = P.make_batch(p, Bn, scale, seed=99)
    impl = "ref" if O.available("ref") else "port"
    cache = O.get_cache(p, impl)
    gold = O.solve_batch(p, b, impl)
    it64, st64, tr64 = admm(p, cache, b, np.float64, "base")
    x64, u64 = admm.last_x, admm.last_u
    print(f"{cfg} B={Bn}: fp64 model vs reference: {int((it64 != gold['iter']).sum())} count mismatches; mean iters {it64.mean():.2f}")
    for band in (0.001, 0.003, 0.01, 0.03, 0.1, 0.3):
        print(f"   band {band:5.3f}: {100 * at_risk(p, it64, tr64, band):6.2f} % of the problems at risk")
    for variant in (sys.argv[4].split(",") if len(sys.argv) > 4 else ("base", "ince", "comp", "comp+ince")):
        it, st, tr = admm(p, cache, b, np.float32, variant)
        bad = (it != gold["iter"]) | (st != gold["status"])
        bn = band_needed(p, it, tr, it64, tr64)
        qs = np.quantile(bn, [0.5, 0.9, 0.99, 0.999, 1.0])
        same = it == it64
        print(f"      drift of the iterate where counts agree: max|x - x64| = {np.abs(admm.last_x - x64)[same].max():.2e}, "
              f"max|u - u64| = {np.abs(admm.last_u - u64)[same].max():.2e}")
        print(f"{cfg} f32 {variant:10s}: {int(bad.sum()):5d}/{Bn} count mismatches; relative residual disagreement near tol: "
              f"median {qs[0]:.2e} p90 {qs[1]:.2e} p99 {qs[2]:.2e} p99.9 {qs[3]:.2e} max {qs[4]:.2e}", flush=True)
        print("      exact-count mode (band: % handed to fp64 / mismatches left): " +
              "  ".join("%g: %.2f%% / %d" % ((bd,) + (lambda r: (100 * r[0], r[1]))(mixed_outcome(p, it, tr, gold["iter"], bd)))
                        for bd in (0.0003, 0.001, 0.003, 0.01, 0.03)), flush=True)


if __name__ == "__main__":
    main()
