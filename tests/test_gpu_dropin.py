"""GPU tests of the drop-in surface: the MATLAB-class mirror (TinyMPC) and the host C++ API mirror driving the
GPU exactly the way the reference's examples drive the CPU solver -- single solves with warm starts
(tiny_solve on the full workspace), closed-loop MPC, the new solve_batch, and the general
warp-per-problem kernel on shapes / feature mixes that have no specialised kernel."""
import importlib

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
P = cases.P


@pytest.fixture(scope="module")
def tm():
    import torch
    assert torch.cuda.is_available()
    return importlib.import_module("tinympc-matlab_b200")


def test_cartpole_one_solve_like_the_matlab_example(tm):
    """examples/cartpole_example_one_solve.m:13-31 -> SURVEY G1 (9 iterations, u[0] = 1.178262262)."""
    p = P.cartpole(u_bound=None, matlab_defaults=True)
    _, _, g = cases.load("G1_cartpole_unconstrained")
    s = tm.TinyMPC()
    s.setup(p.A, p.B, np.diag(p.Qdiag), np.diag(p.Rdiag), 20, rho=1.0)
    s.set_x0([0.5, 0, 0, 0])
    assert s.solve() == 0
    sol, st = s.get_solution(), s.get_stats()
    assert st["iter"] == 9 and st["status"] == 1 and s.last_status == 0
    assert np.abs(sol["controls"].T - g["u"][0]).max() < 1e-10 and np.abs(sol["states"].T - g["x"][0]).max() < 1e-10
    # G2: + set_bound_constraints([], [], -0.5, 0.5) on a fresh solver
    _, _, g = cases.load("G2_cartpole_ubound")
    s = tm.TinyMPC()
    s.setup(p.A, p.B, np.diag(p.Qdiag), np.diag(p.Rdiag), 20, rho=1.0)
    s.set_bound_constraints([], [], -0.5, 0.5)
    s.set_x0([0.5, 0, 0, 0])
    s.solve()
    sol, st = s.get_solution(), s.get_stats()
    assert st["iter"] == 51 and st["status"] == 1
    assert np.abs(sol["controls"].T - g["u"][0]).max() < 1e-10


def test_rocket_through_the_matlab_surface_reproduces_the_cone_swap(tm):
    """examples/rocket_landing_constraints.m:17-78 first solve -> SURVEY G4 (37 iterations), incl. quirk Q3."""
    d = P.rocket(linear=False)
    _, b, g = cases.load("G4_rocket_soc")
    s = tm.TinyMPC()
    s.setup(d.A, d.B, np.diag(d.Qdiag), np.diag(d.Rdiag), 10, rho=1.0, fdyn=d.f, max_iter=100, abs_pri_tol=2e-3)
    s.set_bound_constraints([-5, -5, -0.5, -10, -10, -20], [5, 5, 100, 10, 10, 20], [-10, -10, -10], [105, 105, 105])
    s.set_cone_constraints([0], [3], [0.5], [0], [3], [0.25])      # MATLAB order: state cone mu 0.5, input cone mu 0.25
    s.set_x_ref(b.Xref[0].T.astype(np.float64)); s.set_u_ref(b.Uref[0].T.astype(np.float64))   # the golden run saw float32-rounded refs
    s.set_x0(b.x0[0].astype(np.float64))
    s.solve()
    sol, st = s.get_solution(), s.get_stats()
    assert st["iter"] == g["iter"][0] == 37 and st["status"] == 1
    assert np.abs(sol["controls"].T - g["u"][0]).max() < 1e-9 and np.abs(sol["states"].T - g["x"][0]).max() < 1e-9
    # + linear rows -> G4lin (43 iterations)
    _, _, g = cases.load("G4_rocket_soc_linear")
    s.reset_workspace()
    s.set_linear_constraints([[0, 0, -1.0, 0, 0, 0]], [0.0], [[0, 0, 1.0]], [50.0])
    s.set_x0(b.x0[0].astype(np.float64))
    s.solve()
    st = s.get_stats()
    assert st["iter"] == g["iter"][0] == 43
    assert np.abs(s.get_solution()["controls"].T - g["u"][0]).max() < 1e-9


@pytest.mark.parametrize("name", ["mpc_quadrotor", "mpc_cartpole"])
def test_closed_loop_warm_start_matches_reference(name, tm):
    """The pattern of T/examples/quadrotor_hovering.cpp:73-93: the workspace persists between solves; every one of the 30
    steps must take the reference's iteration count and produce its trajectory."""
    g = np.load(cases.GOLDEN / f"{name}.npz")
    p = P.quadrotor() if "quad" in name else P.cartpole(N=10)
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_x_ref(g["Xref"].T)
    for k in range(len(g["iter"])):
        s.set_x0(g["x0"][k])
        s.solve()
        st = s.get_stats()
        assert st["iter"] == g["iter"][k] and st["status"] == g["status"][k], (k, st["iter"], g["iter"][k])
        assert np.abs(s.get_solution()["states"].T - g["x"][k]).max() < 1e-9
        assert np.abs(s.work_u0() - g["work_u0"][k]).max() < 1e-9


def test_adaptive_rho_single_solve_mutates_cache_like_reference(tm):
    p, b, g = cases.load("G5_quadrotor_adaptive")
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_x_ref(b.Xref[0].T.astype(np.float64))
    s.set_x0(b.x0[0].astype(np.float64))
    s.solve()
    st = s.get_stats()
    assert st["iter"] == 100 and st["status"] == 11 and abs(st["rho"] - 2.44014511) < 1e-7
    assert np.abs(s.get_solution()["states"].T - g["x"][0]).max() < 1e-8


def test_solve_batch_through_the_matlab_surface(tm):
    p, b, g = cases.load("batch_quadrotor_perproblem_bounds")
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_option("precision", 64)
    t = lambda a: np.transpose(a, (2, 1, 0))            # (B, steps, dim) -> MATLAB dim x steps x B
    r = s.solve_batch(b.x0.T, t(b.Xref), t(b.Uref), t(b.x_min), t(b.x_max), t(b.u_min), t(b.u_max))
    assert np.array_equal(r["iter"], g["iter"]) and np.array_equal(r["status"], g["status"])
    assert np.abs(t(r["states"]) - g["x"]).max() < 2e-6 * max(1, np.abs(g["x"]).max())
    # shared reference given as a single column, like set_x_ref accepts (src/TinyMPC.m:393-405)
    p2, b2, g2 = cases.load("batch_cartpole")
    s2 = tm.TinyMPC().setup_from_spec(p2)
    s2.set_option("precision", 64)
    r2 = s2.solve_batch(b2.x0.T, np.zeros(4), 0.0)
    assert np.array_equal(r2["iter"], g2["iter"])


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("name", ["batch_quadrotor", "batch_rocket", "batch_quadrotor_adaptive", "batch_cartpole", "batch_quadrotor_perproblem_bounds"])
def test_general_warp_per_problem_kernel_matches_golden(name, precision, tm):
    p, b, g = cases.load(name)
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_option("precision", precision)
    s.set_option("force_wpp", 1)
    r = s.cuda.solve_batch(b.x0, b.Xref, b.Uref, b.x_min, b.x_max, b.u_min, b.u_max)
    assert s.cuda.last_kernel.startswith("wpp_")
    same = (r["iter"] == g["iter"]) & (r["status"] == g["status"])
    if precision == 64:
        assert same.all()
        assert np.abs(r["rho"] - g["rho"]).max() < 1e-5
    else:
        assert (~same).mean() <= (0.35 if "rocket" in name else 0.1)
    tol = (2e-6 if precision == 64 else 1e-4) * max(1.0, float(np.abs(g["x"]).max()))
    assert np.abs(r["x"][same] - g["x"][same]).max() <= tol and np.abs(r["u"][same] - g["u"][same]).max() <= tol


def test_unspecialised_shapes_and_feature_mixes_fall_back_to_the_general_kernel(tm, oracle_mod):
    # a horizon with no compiled thread-per-problem instance
    p = P.quadrotor(N=7)
    b = P.make_batch(p, 500, 1.0, seed=5)
    g = oracle_mod.solve_batch(p, b, "port")
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_option("precision", 64)
    r = s.cuda.solve_batch(b.x0, b.Xref, b.Uref)
    assert s.cuda.last_kernel == "wpp_f64_generic"
    assert np.array_equal(r["iter"], g["iter"]) and np.abs(r["x"] - g["x"]).max() < 1e-5
    # adaptive rho together with linear constraints (no specialised kernel either)
    p = P.quadrotor(adaptive=True).with_(en_input_linear=1, Alin_u=np.array([[1.0, 1.0, 1.0, 1.0]]), blin_u=np.array([0.8]))
    b = P.make_batch(p, 300, 1.0, seed=6)
    g = oracle_mod.solve_batch(p, b, "port")
    s = tm.TinyMPC().setup_from_spec(p)
    s.set_option("precision", 64)
    r = s.cuda.solve_batch(b.x0, b.Xref, b.Uref)
    assert s.cuda.last_kernel == "wpp_f64_generic"
    assert np.array_equal(r["iter"], g["iter"]) and np.abs(r["x"] - g["x"]).max() < 1e-5 and np.abs(r["rho"] - g["rho"]).max() < 1e-4


def test_edge_cases(tm):
    p = P.cartpole()
    s = tm.TinyMPC().setup_from_spec(p)
    # empty batch
    r = s.cuda.solve_batch(np.zeros((0, 4), np.float32))
    assert r["iter"].shape == (0,)
    # ragged batch sizes around the CTA / chunk granularity, results independent of the batch they sit in
    b = P.make_batch(p, 1000, 1.0, seed=3)
    full = s.cuda.solve_batch(b.x0)
    for n in (1, 31, 33, 257, 999):
        part = s.cuda.solve_batch(b.x0[:n])
        assert np.array_equal(part["iter"], full["iter"][:n]) and np.array_equal(part["x"], full["x"][:n])
    # max_iter = 0: the reference loop never runs -> zero solution, iter 0, status 11
    s.update_settings(max_iter=0)
    r = s.cuda.solve_batch(b.x0[:5])
    assert (r["iter"] == 0).all() and (r["status"] == 11).all() and not r["x"].any()
    # bounds enabled but never supplied -> a clear error, not garbage (the batched entry validates)
    q = P.cartpole(u_bound=None).with_(en_input_bound=1)
    s2 = tm.CudaSolver()
    fam = cases.family_from_spec(q, tm.TinyMPC().setup_from_spec(P.cartpole(u_bound=None)).get_cache() | dict(dKinf=None, dPinf=None))
    s2.set_family(fam)
    with pytest.raises(tm.TinympcCudaError):
        s2.solve_batch(b.x0[:4])
    # check_termination = 0 is rejected (the reference divides by it, admm.cpp:255)
    fam["check_termination"] = 0
    with pytest.raises(tm.TinympcCudaError):
        s2.set_family(fam)


# ------------------------------------------------------------------------------------------------------------
# The reference's own host code on top of the GPU hot path: oracle/_ref/libtinympc_refhost_b200.so is the UNMODIFIED
# reference tiny_api.cpp (tiny_setup, Riccati cache precompute, setters; Eigen) linked WITHOUT admm.cpp /
# rho_benchmark.cpp -- its solve() is oracle/ref_b200_binding.cpp, the binding of INTEGRATION.md option B, which
# forwards the live TinySolver to tinympc_cuda_solve_workspace.  Same driver, same inputs as the golden files.
# ------------------------------------------------------------------------------------------------------------
REFHOST_CASES = ["G1_cartpole_unconstrained", "G2_cartpole_ubound", "G3_quadrotor_hover", "G4_rocket_soc",
                 "G4_rocket_soc_linear", "G5_quadrotor_adaptive", "batch_cartpole", "batch_quadrotor", "batch_rocket",
                 "batch_quadrotor_adaptive", "batch_quadrotor_perproblem_bounds", "batch_quadrotor_check3_max20"]


@pytest.mark.parametrize("name", REFHOST_CASES)
def test_reference_host_code_with_gpu_hot_path(name, oracle_mod):
    O = oracle_mod
    if not O.available("refhost_b200"):
        pytest.skip("oracle/_ref/libtinympc_refhost_b200.so not built (needs /root/reference at build time)")
    p, b, g = cases.load(name)
    n = min(b.size, 16)
    r = O.solve_batch(p, b.slice(0, n), "refhost_b200", threads=1)
    assert np.array_equal(r["iter"], g["iter"][:n]), (r["iter"], g["iter"][:n])
    assert np.array_equal(r["status"], g["status"][:n])
    sx, su = max(1.0, np.abs(g["x"][:n]).max()), max(1.0, np.abs(g["u"][:n]).max())
    assert np.abs(r["x"] - g["x"][:n]).max() < 1e-9 * sx
    assert np.abs(r["u"] - g["u"][:n]).max() < 1e-9 * su


@pytest.mark.parametrize("name", ["mpc_quadrotor", "mpc_cartpole"])
def test_reference_host_closed_loop_with_gpu_hot_path(name, oracle_mod):
    """Warm-started closed loop (quadrotor_hovering.cpp:73-93) through the reference's API, hot path on the GPU."""
    O = oracle_mod
    if not O.available("refhost_b200"):
        pytest.skip("oracle/_ref/libtinympc_refhost_b200.so not built")
    g = np.load(cases.GOLDEN / f"{name}.npz")
    p = P.quadrotor() if "quad" in name else P.cartpole(N=10)
    s = O.Session(p, "refhost_b200")
    s.set_x_ref(g["Xref"])
    for k in range(len(g["iter"])):
        s.set_x0(g["x0"][k])
        r = s.solve()
        assert r["iter"] == g["iter"][k] and r["status"] == g["status"][k], (k, r["iter"], g["iter"][k])
        assert np.abs(r["x"] - g["x"][k]).max() < 1e-9
        assert np.abs(r["work_u0"] - g["work_u0"][k]).max() < 1e-9
    s.close()


# ------------------------------------------------------------------------------------------------------------
# Sessions: a batch of warm-started solvers resident on the device
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["quadrotor", "cartpole", "rocket", "quadrotor_adaptive"])
def test_session_closed_loop_matches_reference_per_problem(cfg, tm, oracle_mod):
    """{set_x0; solve; x0 = A x0 + B u0 + f} for 12 steps on 24 independent systems at once (fp64): every solver of the
    session must take exactly the iterations, and produce the trajectory, of a reference solver driven the same way."""
    O = oracle_mod
    p = {"quadrotor": P.quadrotor, "cartpole": lambda: P.cartpole(N=10), "rocket": P.rocket,
         "quadrotor_adaptive": lambda: P.quadrotor(adaptive=True)}[cfg]()
    impl = "ref" if O.available("ref") else "port"
    B, steps = 24, 12
    b = P.make_batch(p, B, 0.6, seed=11)
    s = tm.TinyMPC().setup_from_spec(p, devices=[0])
    s.cuda.set_option("precision", 64)
    ses = s.cuda.session(B)
    Xref = b.Xref.astype(np.float64) if b.Xref is not None else np.zeros((B, p.N, p.nx))
    Uref = b.Uref.astype(np.float64) if b.Uref is not None else np.zeros((B, p.N - 1, p.nu))
    ses.set_x_ref(Xref); ses.set_u_ref(Uref)
    x0 = b.x0.astype(np.float64)
    ses.set_x0(x0)
    refs = [O.Session(p, impl) for _ in range(B)]
    for k in range(B):
        refs[k].set_x_ref(Xref[k]); refs[k].set_u_ref(Uref[k])
    Bm = np.asarray(p.B).reshape(p.nx, p.nu)
    xr = x0.copy()
    for t in range(steps):
        ses.solve()
        it, st, sx, wu = ses.read("iter"), ses.read("status"), ses.read("sol_x"), ses.read("u")
        for k in range(B):
            refs[k].set_x0(xr[k])
            r = refs[k].solve()
            assert it[k] == r["iter"] and st[k] == r["status"], (cfg, t, k, it[k], r["iter"])
            assert np.abs(sx[k] - r["x"]).max() < 1e-8 * max(1.0, np.abs(r["x"]).max())
            assert np.abs(wu[k, 0] - r["work_u0"]).max() < 1e-8 * max(1.0, np.abs(r["work_u0"]).max())
            xr[k] = p.A @ xr[k] + Bm @ r["work_u0"] + p.f
        ses.step()                                        # on the device
        assert np.abs(ses.read("x0") - xr).max() < 1e-8 * max(1.0, np.abs(xr).max())
    for r in refs:
        r.close()
    ses.close()


def test_session_fp32_tracks_the_reference(tm, oracle_mod):
    """fp32 session: same closed loop within the fp32 tolerance of the north star (1e-4 on states / controls)."""
    O = oracle_mod
    p = P.quadrotor()
    impl = "ref" if O.available("ref") else "port"
    B, steps = 16, 8
    b = P.make_batch(p, B, 0.3, seed=5)
    s = tm.TinyMPC().setup_from_spec(p, devices=[0])
    ses = s.cuda.session(B)                      # precision option 32 (default)
    ses.set_x_ref(b.Xref.astype(np.float64)); ses.set_x0(b.x0.astype(np.float64))
    refs = [O.Session(p, impl) for _ in range(B)]
    xr = b.x0.astype(np.float64)
    Bm = np.asarray(p.B).reshape(p.nx, p.nu)
    for k in range(B):
        refs[k].set_x_ref(b.Xref[k].astype(np.float64))
    for t in range(steps):
        ses.solve()
        sx = ses.read("sol_x")
        for k in range(B):
            refs[k].set_x0(xr[k]); r = refs[k].solve()
            assert np.abs(sx[k] - r["x"]).max() < 2e-3        # a flipped iteration count moves the iterate by O(tol)
            xr[k] = p.A @ xr[k] + Bm @ r["work_u0"] + p.f
        ses.step()
        assert np.abs(ses.read("x0") - xr).max() < 2e-3
    ses.close()
