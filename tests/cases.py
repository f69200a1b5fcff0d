"""Golden case table: file name under tests/golden -> the ProblemSpec that produced it
(the same table oracle/gen_golden.py used with the unmodified reference)."""
import importlib
from pathlib import Path

import numpy as np

P = importlib.import_module("tinympc-matlab_b200.problems")
GOLDEN = Path(__file__).resolve().parent / "golden"

CASES = {
    "G1_cartpole_unconstrained": lambda: P.cartpole(u_bound=None, matlab_defaults=True),
    "G2_cartpole_ubound": lambda: P.cartpole(u_bound=0.5, matlab_defaults=True),
    "G3_quadrotor_hover": lambda: P.quadrotor(),
    "G4_rocket_soc": lambda: P.rocket(linear=False),
    "G4_rocket_soc_linear": lambda: P.rocket(linear=True),
    "G5_quadrotor_adaptive": lambda: P.quadrotor(adaptive=True),
    "batch_cartpole": lambda: P.cartpole(),
    "batch_cartpole_easy": lambda: P.cartpole(),
    "batch_quadrotor": lambda: P.quadrotor(),
    "batch_quadrotor_easy": lambda: P.quadrotor(),
    "batch_rocket": lambda: P.rocket(),
    "batch_rocket_nolinear": lambda: P.rocket(linear=False),
    "batch_quadrotor_adaptive": lambda: P.quadrotor(adaptive=True),
    "batch_quadrotor_perproblem_bounds": lambda: P.quadrotor(),
    "batch_quadrotor_check3_max20": lambda: P.quadrotor().with_(check_termination=3, max_iter=20),
}


def load(name):
    g = dict(np.load(GOLDEN / f"{name}.npz"))
    b = P.Batch(g["x0"], g.get("Xref"), g.get("Uref"), g.get("x_min"), g.get("x_max"), g.get("u_min"), g.get("u_max"))
    return CASES[name](), b, g


def family_from_spec(p, cache):
    """ProblemSpec + cache dict (Kinf, Pinf, Quu_inv, AmBKt, APf, BPf, dKinf, dPinf) -> the dict
    tinympc-matlab_b200/capi.py:family_struct expects (= contents of a TinySolver)."""
    fam = dict(nx=p.nx, nu=p.nu, N=p.N, Adyn=p.A, Bdyn=p.B, fdyn=p.f, Q=p.Qdiag + p.rho, R=p.Rdiag + p.rho, rho=p.rho,
               Kinf=cache["Kinf"], Pinf=cache["Pinf"], Quu_inv=cache["Quu_inv"], AmBKt=cache["AmBKt"],
               APf=cache["APf"], BPf=cache["BPf"])
    for k in ("abs_pri_tol", "abs_dua_tol", "max_iter", "check_termination", "en_state_bound", "en_input_bound",
              "en_state_soc", "en_input_soc", "en_state_linear", "en_input_linear", "adaptive_rho", "adaptive_rho_min",
              "adaptive_rho_max", "adaptive_rho_enable_clipping", "x_min", "x_max", "u_min", "u_max",
              "Acx", "qcx", "cx", "Acu", "qcu", "cu", "Alin_x", "blin_x", "Alin_u", "blin_u"):
        fam[k] = getattr(p, k)
    if p.adaptive_rho:
        fam["dKinf_drho"], fam["dPinf_drho"] = cache["dKinf"], cache["dPinf"]
    return fam
