"""Problem descriptions and synthetic batch generators for the BASELINE.json configs.

Pure data + numpy; no solver arithmetic lives here.  The numbers come from the reference's own
examples (extracted once by oracle/extract_problem_data.py into tinympc-matlab_b200/problem_data.json):

  cartpole  : examples/cartpole_example_one_solve.m:13-20
  quadrotor : tinympc/TinyMPC/examples/problem_data/quadrotor_20hz_params.hpp:5-37,
              bounds of tinympc/TinyMPC/examples/quadrotor_hovering.cpp:41-44
  rocket    : examples/rocket_landing_constraints.m:17-78

Memory conventions: a trajectory the reference stores as a column-major ``nx x N`` Eigen matrix is a
C-ordered numpy array of shape ``(N, nx)`` here (identical bytes); a batch is ``(B, N, nx)``.
Small matrices (A, B, Kinf, ...) use the mathematical ``(rows, cols)`` shape.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, replace
from pathlib import Path
from typing import Optional

import numpy as np

_DATA = Path(__file__).resolve().parent / "problem_data.json"


@dataclass
class ProblemSpec:
    """Everything tiny_setup + the constraint setters + TinySettings hold (types.hpp:43-187)."""
    name: str
    nx: int
    nu: int
    N: int
    A: np.ndarray
    B: np.ndarray
    f: np.ndarray
    Qdiag: np.ndarray          # the user's diag(Q), before "+rho" (tiny_api.cpp:107)
    Rdiag: np.ndarray
    rho: float
    # TinySettings; defaults are the C++ ones (tiny_api_constants.hpp:5-14, tiny_api.cpp:347-373)
    abs_pri_tol: float = 1e-3
    abs_dua_tol: float = 1e-3
    max_iter: int = 1000
    check_termination: int = 1
    en_state_bound: int = 0
    en_input_bound: int = 0
    en_state_soc: int = 0
    en_input_soc: int = 0
    en_state_linear: int = 0
    en_input_linear: int = 0
    adaptive_rho: int = 0
    adaptive_rho_min: float = 1.0
    adaptive_rho_max: float = 100.0
    adaptive_rho_enable_clipping: int = 1
    # shared bounds, shapes (N, nx) and (N-1, nu)
    x_min: Optional[np.ndarray] = None
    x_max: Optional[np.ndarray] = None
    u_min: Optional[np.ndarray] = None
    u_max: Optional[np.ndarray] = None
    # cones as they land in the workspace (state = work->Acx/qcx/cx, input = work->Acu/qcu/cu)
    Acx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    qcx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    cx: np.ndarray = field(default_factory=lambda: np.zeros(0))
    Acu: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    qcu: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    cu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # linear rows
    Alin_x: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    blin_x: np.ndarray = field(default_factory=lambda: np.zeros(0))
    Alin_u: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    blin_u: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # adaptive-rho sensitivities: 0 none, 1 hard-coded quadrotor tables, 2 explicit
    sens_mode: int = 0
    dKinf: Optional[np.ndarray] = None
    dPinf: Optional[np.ndarray] = None
    dC1: Optional[np.ndarray] = None
    dC2: Optional[np.ndarray] = None

    def with_(self, **kw) -> "ProblemSpec":
        return replace(self, **kw)


def _load():
    return json.loads(_DATA.read_text())


def _bounds(lo, hi, steps):
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    return np.tile(lo, (steps, 1)), np.tile(hi, (steps, 1))


def cartpole(N: int = 20, u_bound: Optional[float] = 0.5, matlab_defaults: bool = False) -> ProblemSpec:
    """Config 1/2.  ``matlab_defaults`` = the TinyMPC.m settings (tol 1e-4, max_iter 100,
    src/TinyMPC.m:26-39); otherwise the C2 benchmark settings (tol 1e-3, max_iter 100)."""
    d = _load()["cartpole"]
    p = ProblemSpec("cartpole", 4, 1, N, np.array(d["A"]), np.array(d["B"]), np.array(d["f"]),
                    np.array(d["Q"]), np.array(d["R"]), d["rho"], max_iter=100)
    if matlab_defaults:
        p.abs_pri_tol = p.abs_dua_tol = 1e-4
    if u_bound is not None:
        # set_bound_constraints([], [], -u, u): x defaults to +-1e17 (src/TinyMPC.m:261-264)
        p.x_min, p.x_max = _bounds([-1e17] * 4, [1e17] * 4, N)
        p.u_min, p.u_max = _bounds([-u_bound], [u_bound], N - 1)
        p.en_state_bound = p.en_input_bound = 1
    return p


def quadrotor(N: int = 10, adaptive: bool = False) -> ProblemSpec:
    """Config 3 (and 5 with ``adaptive``): quadrotor_hovering.cpp data, x in +-5, u in +-0.5."""
    d = _load()["quadrotor"]
    p = ProblemSpec("quadrotor", 12, 4, N, np.array(d["A"]), np.array(d["B"]), np.array(d["f"]),
                    np.array(d["Q"]), np.array(d["R"]), d["rho"], max_iter=100)
    p.x_min, p.x_max = _bounds([-5.0] * 12, [5.0] * 12, N)
    p.u_min, p.u_max = _bounds([-0.5] * 4, [0.5] * 4, N - 1)
    p.en_state_bound = p.en_input_bound = 1
    if adaptive:
        p.adaptive_rho = 1
        p.sens_mode = 1
        p.name = "quadrotor_adaptive"
    return p


def rocket(N: int = 10, linear: bool = True) -> ProblemSpec:
    """Config 4: rocket_landing_constraints.m, bounds + one state cone + one input cone as the MEX
    layer lands them (MATLAB cx=0.5, cu=0.25 -> work->cx=0.25, work->cu=0.5),
    plus the two linear rows of SURVEY section 8c/G4 when ``linear``."""
    d = _load()["rocket"]
    p = ProblemSpec("rocket", 6, 3, N, np.array(d["A"]), np.array(d["B"]), np.array(d["f"]),
                    np.array(d["Q"]), np.array(d["R"]), d["rho"], max_iter=100,
                    abs_pri_tol=2e-3, abs_dua_tol=1e-4)
    p.x_min, p.x_max = _bounds([-5, -5, -0.5, -10, -10, -20], [5, 5, 100, 10, 10, 20], N)
    p.u_min, p.u_max = _bounds([-10] * 3, [105] * 3, N - 1)
    p.en_state_bound = p.en_input_bound = 1
    p.Acx, p.qcx, p.cx = np.array([0], np.int32), np.array([3], np.int32), np.array([0.25])
    p.Acu, p.qcu, p.cu = np.array([0], np.int32), np.array([3], np.int32), np.array([0.5])
    p.en_state_soc = p.en_input_soc = 1
    if linear:
        p.Alin_x, p.blin_x = np.array([[0, 0, -1.0, 0, 0, 0]]), np.array([0.0])
        p.Alin_u, p.blin_u = np.array([[0, 0, 1.0]]), np.array([50.0])
        p.en_state_linear = p.en_input_linear = 1
    else:
        p.name = "rocket_nolinear"
    return p


# Relative band of the exact-count ("mixed") mode per problem family: the fp32 pass hands every problem whose termination
# decision falls within this band of a tolerance to the fp64 kernel (tinympc_cuda_set_option "mixed").  Measured so that 2^18
# random problems of the family reproduce the reference's iteration counts and status codes without exception.
EXACT_BAND = {"cartpole": 0.003, "quadrotor": 0.002, "rocket": 0.003, "rocket_nolinear": 0.003, "quadrotor_adaptive": 0.3}


def exact_band(p: "ProblemSpec") -> float:
    return EXACT_BAND.get(p.name, 0.3)


# Families whose parity-exact mode is plain fp64: under adaptive rho the fp32 direct-form kernel needs a 30 % band (two thirds of the
# batch re-solved) and its un-marked problems still drift 5e-4 away at max_iter, so the lane-group fp64 kernel (tmpc_gpp.cuh) solves
# the whole batch -- faster than the mixed mode, and exact.
EXACT_PRECISION = {"quadrotor_adaptive": 64}


def exact_precision(p: "ProblemSpec") -> int:
    return EXACT_PRECISION.get(p.name, 32)


ROCKET_XINIT = np.array([4.0, 2.0, 20.0, -3.0, 2.0, -4.5])


def rocket_refs(N: int = 10, k: int = 0, ntotal: int = 100):
    """Xref/Uref of rocket_landing_constraints.m:72-74 (straight line to the origin, Uref[2]=10)."""
    xg = np.zeros(6)
    Xref = np.stack([ROCKET_XINIT + (xg - ROCKET_XINIT) * (i + k) / (ntotal - 1) for i in range(N)])
    Uref = np.zeros((N - 1, 3))
    Uref[:, 2] = 10.0
    return Xref, Uref


@dataclass
class Batch:
    x0: np.ndarray                       # (B, nx) float32
    Xref: Optional[np.ndarray] = None    # (B, N, nx) float32 or None (= zeros)
    Uref: Optional[np.ndarray] = None    # (B, N-1, nu) float32 or None
    x_min: Optional[np.ndarray] = None   # optional per-problem bounds (B, N, nx) float32
    x_max: Optional[np.ndarray] = None
    u_min: Optional[np.ndarray] = None   # (B, N-1, nu)
    u_max: Optional[np.ndarray] = None

    @property
    def size(self) -> int:
        return int(self.x0.shape[0])

    def slice(self, lo: int, hi: int) -> "Batch":
        s = lambda a: None if a is None else a[lo:hi]
        return Batch(self.x0[lo:hi], s(self.Xref), s(self.Uref), s(self.x_min), s(self.x_max), s(self.u_min), s(self.u_max))


def make_batch(p: ProblemSpec, B: int, scale: float = 1.0, seed: Optional[int] = None) -> Batch:
    """Synthetic batches of SURVEY section 8d: numpy default_rng in float64, rounded to float32;
    the SAME float32 values go to the CPU oracle and to the GPU."""
    base = p.name.split("_")[0]
    if seed is None:
        seed = 1234 + {"cartpole": 2, "quadrotor": 3, "rocket": 4}[base] + (2 if p.adaptive_rho else 0)
    rng = np.random.default_rng(seed)
    U = lambda *shape: rng.uniform(-1.0, 1.0, size=shape)
    N = p.N
    if base == "cartpole":
        x0 = scale * U(B, 4) * np.array([0.5, 0.2, 0.1, 0.2])
        return Batch(x0.astype(np.float32))
    if base == "quadrotor":
        amp = np.array([0.5] * 3 + [0.1] * 3 + [0.3] * 3 + [0.2] * 3)
        x0 = scale * U(B, 12) * amp
        xr = np.zeros((B, 12))
        xr[:, :3] = scale * 0.5 * U(B, 3)
        Xref = np.repeat(xr[:, None, :], N, axis=1)          # replicated over the horizon, passed in full
        Uref = np.zeros((B, N - 1, 4))
        return Batch(x0.astype(np.float32), Xref.astype(np.float32), Uref.astype(np.float32))
    if base == "rocket":
        x0 = 1.1 * ROCKET_XINIT * (1.0 + 0.1 * scale * U(B, 6))
        Xr, Ur = rocket_refs(N)
        Xref = np.broadcast_to(Xr, (B, N, 6)).copy()
        Uref = np.broadcast_to(Ur, (B, N - 1, 3)).copy()
        return Batch(x0.astype(np.float32), Xref.astype(np.float32), Uref.astype(np.float32))
    raise ValueError(p.name)
