#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference C++ (oracle/_ref).

TEST INFRASTRUCTURE.  Runs only in the build container (needs oracle/_ref/libtinympc_ref.so, built
by `make -C oracle ref` from /root/reference).  The reference's own tests pin no solve result
(SURVEY.md section 4), so these vectors -- outputs of the reference itself -- are the parity pin.
The .npz files are committed; tests never need /root/reference.

Cases
  G1..G5     the single-problem known-answer cases of SURVEY.md section 8c
  batch_*    64-problem seeded batches of every BASELINE.json config (inputs + reference outputs)
  cache_*    the cache tiny_setup computes for each problem family
  mpc_*      warm-started closed-loop sequences (the pattern of quadrotor_hovering.cpp:73-93)
"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
P = importlib.import_module("tinympc-matlab_b200.problems")
import oracle as O  # noqa: E402

OUT = ROOT / "tests" / "golden"


def save(name, batch, res, **extra):
    d = dict(x0=batch.x0, x=res["x"], u=res["u"], iter=res["iter"], status=res["status"],
             residuals=res["residuals"], rho=res["rho"], **extra)
    for k in ("Xref", "Uref", "x_min", "x_max", "u_min", "u_max"):
        v = getattr(batch, k)
        if v is not None:
            d[k] = v
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(f"{name}: B={batch.size} iters mean {res['iter'].mean():.2f} unsolved {(res['status'] == 11).mean():.3f}")


def main():
    f32 = np.float32
    # ---- G1, G2: examples/cartpole_example_one_solve.m:13-31 (MATLAB settings) ----
    b = P.Batch(np.array([[0.5, 0, 0, 0]], f32))
    save("G1_cartpole_unconstrained", b, O.solve_batch(P.cartpole(u_bound=None, matlab_defaults=True), b))
    save("G2_cartpole_ubound", b, O.solve_batch(P.cartpole(u_bound=0.5, matlab_defaults=True), b))
    # ---- G3, G5: quadrotor_hovering.cpp:35-66 ----
    x0 = np.zeros((1, 12), f32); x0[0, 1] = 1; x0[0, 3] = 0.2; x0[0, 6] = 0.1
    Xref = np.zeros((1, 10, 12), f32); Xref[:, :, 2] = 2
    b = P.Batch(x0, Xref, None)
    save("G3_quadrotor_hover", b, O.solve_batch(P.quadrotor(), b))
    save("G5_quadrotor_adaptive", b, O.solve_batch(P.quadrotor(adaptive=True), b))
    # ---- G4: examples/rocket_landing_constraints.m:17-78, first solve ----
    Xr, Ur = P.rocket_refs()
    b = P.Batch((1.1 * P.ROCKET_XINIT)[None].astype(f32), Xr[None].astype(f32), Ur[None].astype(f32))
    save("G4_rocket_soc", b, O.solve_batch(P.rocket(linear=False), b))
    save("G4_rocket_soc_linear", b, O.solve_batch(P.rocket(linear=True), b))

    # ---- seeded batches of every config ----
    for name, p, scale in [("batch_cartpole", P.cartpole(), 1.0), ("batch_cartpole_easy", P.cartpole(), 0.3),
                           ("batch_quadrotor", P.quadrotor(), 1.0), ("batch_quadrotor_easy", P.quadrotor(), 0.3),
                           ("batch_rocket", P.rocket(), 1.0), ("batch_rocket_nolinear", P.rocket(linear=False), 1.0),
                           ("batch_quadrotor_adaptive", P.quadrotor(adaptive=True), 1.0)]:
        b = P.make_batch(p, 64, scale)
        save(name, b, O.solve_batch(p, b))
    # per-problem bounds (BASELINE.json: problems "differ in x0, references and bounds")
    p = P.quadrotor()
    b = P.make_batch(p, 64, 1.0, seed=77)
    rng = np.random.default_rng(78)
    ub = rng.uniform(0.2, 0.6, size=(64, 1, 4)) * np.ones((1, p.N - 1, 1))
    xb = rng.uniform(0.5, 5.0, size=(64, 1, 12)) * np.ones((1, p.N, 1))
    b.u_min, b.u_max = (-ub).astype(f32), ub.astype(f32)
    b.x_min, b.x_max = (-xb).astype(f32), xb.astype(f32)
    save("batch_quadrotor_perproblem_bounds", b, O.solve_batch(p, b))
    # settings edge cases: check_termination=3, max_iter small
    p = P.quadrotor().with_(check_termination=3, max_iter=20)
    b = P.make_batch(p, 64, 0.5, seed=91)
    save("batch_quadrotor_check3_max20", b, O.solve_batch(p, b))

    # ---- caches ----
    for name, p in [("cartpole", P.cartpole()), ("quadrotor", P.quadrotor(adaptive=True)), ("rocket", P.rocket())]:
        np.savez_compressed(OUT / f"cache_{name}.npz", **O.get_cache(p))

    # ---- warm-started closed loop (quadrotor_hovering.cpp:73-93): x0 <- A x0 + B work->u.col(0) ----
    for name, p in [("mpc_quadrotor", P.quadrotor()), ("mpc_cartpole", P.cartpole(N=10))]:
        s = O.Session(p)
        if p.nx == 12:
            x0 = np.zeros(12); x0[1] = 1; x0[3] = 0.2; x0[6] = 0.1
            xr = np.zeros((p.N, 12)); xr[:, 2] = 2
        else:
            x0 = np.array([0.5, 0, 0, 0.0]); xr = np.zeros((p.N, 4))
        s.set_x_ref(xr)
        rec = dict(x0=[], x=[], u=[], iter=[], status=[], work_u0=[])
        for _ in range(30):
            s.set_x0(x0)
            r = s.solve()
            rec["x0"].append(x0.copy())
            for k in ("x", "u", "iter", "status", "work_u0"):
                rec[k].append(r[k])
            x0 = p.A @ x0 + np.asarray(p.B).reshape(p.nx, p.nu) @ r["work_u0"] + p.f
        s.close()
        np.savez_compressed(OUT / f"{name}.npz", Xref=xr, **{k: np.array(v) for k, v in rec.items()})
        print(name, "iters", np.array(rec["iter"])[:10])


if __name__ == "__main__":
    main()
