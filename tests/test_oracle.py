"""CPU tests of the checker itself: the C port (oracle/tinympc_oracle.c) against the golden vectors
produced by the unmodified reference, and against the reference library when it is present."""
import numpy as np
import pytest

import cases

TOL = 1e-9  # double vs double, different summation order only


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_port_matches_reference_golden(name, oracle_mod):
    p, b, g = cases.load(name)
    r = oracle_mod.solve_batch(p, b, "port")
    assert np.array_equal(r["iter"], g["iter"])
    assert np.array_equal(r["status"], g["status"])
    assert np.abs(r["x"] - g["x"]).max() < TOL
    assert np.abs(r["u"] - g["u"]).max() < TOL
    assert np.abs(r["rho"] - g["rho"]).max() < 1e-9
    assert np.abs(r["residuals"] - g["residuals"]).max() < 1e-9


def test_survey_known_answers():
    """The literal numbers of SURVEY.md section 8c (G1, G2, G3) are what the golden files hold."""
    _, _, g = cases.load("G1_cartpole_unconstrained")
    assert g["iter"][0] == 9 and g["status"][0] == 1
    assert abs(g["u"][0, 0, 0] - 1.178262262) < 1e-8 and abs(g["u"][0, 18, 0] + 0.6795788863) < 1e-8
    assert np.abs(g["x"][0, 19] - [0.511176418, 0.006026307871, 0.04467673962, 0.1480966804]).max() < 1e-8
    _, _, g = cases.load("G2_cartpole_ubound")
    assert g["iter"][0] == 51 and abs(g["u"][0, 6, 0] - 0.453693677) < 1e-8
    _, _, g = cases.load("G3_quadrotor_hover")
    assert g["iter"][0] == 100 and g["status"][0] == 11 and abs(g["x"][0, 1, 8] - 0.436747853) < 1e-8
    _, _, g = cases.load("G5_quadrotor_adaptive")
    assert abs(g["rho"][0] - 2.44014511) < 1e-7
    _, _, g = cases.load("G4_rocket_soc")
    assert g["iter"][0] == 37 and abs(g["u"][0, 0, 2] - 65.24071) < 1e-4
    _, _, g = cases.load("G4_rocket_soc_linear")
    assert g["iter"][0] == 43 and abs(g["u"][0, 0, 2] - 50.000262) < 1e-5


@pytest.mark.parametrize("family", ["cartpole", "quadrotor", "rocket"])
def test_port_cache_matches_reference(family, oracle_mod, problems):
    p = dict(cartpole=problems.cartpole(), quadrotor=problems.quadrotor(adaptive=True), rocket=problems.rocket())[family]
    g = np.load(cases.GOLDEN / f"cache_{family}.npz")
    c = oracle_mod.get_cache(p, "port")
    for k in g.files:
        assert np.abs(c[k] - g[k]).max() < 1e-8 * max(1.0, np.abs(g[k]).max()), k


@pytest.mark.parametrize("name", ["mpc_quadrotor", "mpc_cartpole"])
def test_port_warm_started_session(name, oracle_mod, problems):
    g = np.load(cases.GOLDEN / f"{name}.npz")
    p = problems.quadrotor() if "quad" in name else problems.cartpole(N=10)
    s = oracle_mod.Session(p, "port")
    s.set_x_ref(g["Xref"])
    for k in range(len(g["iter"])):
        s.set_x0(g["x0"][k])
        r = s.solve()
        assert r["iter"] == g["iter"][k] and r["status"] == g["status"][k]
        assert np.abs(r["x"] - g["x"][k]).max() < 1e-8 and np.abs(r["work_u0"] - g["work_u0"][k]).max() < 1e-8
    s.close()


def test_port_matches_reference_library_random(oracle_mod, problems):
    if not oracle_mod.available("ref"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for p in (problems.cartpole(), problems.quadrotor(), problems.rocket(), problems.quadrotor(adaptive=True)):
        b = problems.make_batch(p, 300, 1.0, seed=4242)
        a, c = oracle_mod.solve_batch(p, b, "ref"), oracle_mod.solve_batch(p, b, "port")
        assert np.array_equal(a["iter"], c["iter"]) and np.array_equal(a["status"], c["status"])
        assert np.abs(a["x"] - c["x"]).max() < TOL and np.abs(a["u"] - c["u"]).max() < TOL
