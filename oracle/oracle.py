"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (tinympc-matlab_b200/) never does.

  impl="ref"  -> oracle/_ref/libtinympc_ref.so : the unmodified reference C++ (built from
                 /root/reference by oracle/Makefile) behind oracle/ref_driver.cpp
  impl="port" -> oracle/liboracle_port.so      : the C restatement oracle/tinympc_oracle.c
  impl="refhost_b200" -> oracle/_ref/libtinympc_refhost_b200.so : the drop-in proof -- the unmodified reference
                 tiny_api.cpp behind the same driver, with ONLY solve() (admm.cpp) replaced by the B200 C ABI
                 (oracle/ref_b200_binding.cpp).  This one needs a GPU; it is the thing under test, not a checker.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIBS = {"ref": HERE / "_ref" / "libtinympc_ref.so", "ref_O2": HERE / "_ref" / "libtinympc_ref_O2.so", "port": HERE / "liboracle_port.so",
         "refhost_b200": HERE / "_ref" / "libtinympc_refhost_b200.so"}
_PREFIX = {"ref": "ref", "ref_O2": "ref", "port": "port", "refhost_b200": "ref"}
_loaded = {}

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


class CProblem(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("nu", C.c_int), ("N", C.c_int),
        ("A", c_dp), ("B", c_dp), ("f", c_dp), ("Qdiag", c_dp), ("Rdiag", c_dp), ("rho", C.c_double),
        ("abs_pri_tol", C.c_double), ("abs_dua_tol", C.c_double),
        ("max_iter", C.c_int), ("check_termination", C.c_int),
        ("en_state_bound", C.c_int), ("en_input_bound", C.c_int),
        ("en_state_soc", C.c_int), ("en_input_soc", C.c_int),
        ("en_state_linear", C.c_int), ("en_input_linear", C.c_int),
        ("adaptive_rho", C.c_int), ("adaptive_rho_min", C.c_double), ("adaptive_rho_max", C.c_double),
        ("adaptive_rho_enable_clipping", C.c_int),
        ("x_min", c_dp), ("x_max", c_dp), ("u_min", c_dp), ("u_max", c_dp),
        ("n_state_cones", C.c_int), ("Acx", c_ip), ("qcx", c_ip), ("cx", c_dp),
        ("n_input_cones", C.c_int), ("Acu", c_ip), ("qcu", c_ip), ("cu", c_dp),
        ("n_state_lin", C.c_int), ("Alin_x", c_dp), ("blin_x", c_dp),
        ("n_input_lin", C.c_int), ("Alin_u", c_dp), ("blin_u", c_dp),
        ("sens_mode", C.c_int), ("dKinf", c_dp), ("dPinf", c_dp), ("dC1", c_dp), ("dC2", c_dp),
    ]


class CBatchIn(C.Structure):
    _fields_ = [("batch", C.c_int), ("x0", c_fp), ("Xref", c_fp), ("Uref", c_fp),
                ("x_min", c_fp), ("x_max", c_fp), ("u_min", c_fp), ("u_max", c_fp)]


class CBatchOut(C.Structure):
    _fields_ = [("x", c_dp), ("u", c_dp), ("iter", c_ip), ("status", c_ip), ("residuals", c_dp), ("rho", c_dp)]


class CCacheOut(C.Structure):
    _fields_ = [(n, c_dp) for n in ("Kinf", "Pinf", "Quu_inv", "AmBKt", "APf", "BPf", "dKinf", "dPinf", "dC1", "dC2")]


def available(impl: str) -> bool:
    return _LIBS[impl].exists()


def lib(impl: str):
    if impl not in _loaded:
        path = _LIBS[impl]
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (impl={impl})")
        _loaded[impl] = C.CDLL(str(path))
    return _loaded[impl]


class _Keep:
    """Holds the numpy arrays a ctypes struct points into."""

    def __init__(self):
        self.refs = []

    def d(self, a, fortran=False):
        if a is None:
            return None
        a = np.asarray(a, np.float64)
        a = np.asfortranarray(a).ravel(order="F") if fortran else np.ascontiguousarray(a).ravel()
        a = np.ascontiguousarray(a)
        self.refs.append(a)
        return a.ctypes.data_as(c_dp)

    def i(self, a):
        a = np.ascontiguousarray(np.asarray(a, np.int32)).ravel()
        self.refs.append(a)
        return a.ctypes.data_as(c_ip)

    def f(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, np.float32))
        self.refs.append(a)
        return a.ctypes.data_as(c_fp)


def c_problem(p, keep: _Keep) -> CProblem:
    """ProblemSpec (tinympc-matlab_b200/problems.py) -> oracle_problem (oracle_abi.h)."""
    cp = CProblem()
    cp.nx, cp.nu, cp.N = p.nx, p.nu, p.N
    cp.A, cp.B = keep.d(p.A, True), keep.d(np.asarray(p.B).reshape(p.nx, p.nu), True)
    cp.f, cp.Qdiag, cp.Rdiag, cp.rho = keep.d(p.f), keep.d(p.Qdiag), keep.d(p.Rdiag), float(p.rho)
    for k in ("abs_pri_tol", "abs_dua_tol", "max_iter", "check_termination", "en_state_bound", "en_input_bound",
              "en_state_soc", "en_input_soc", "en_state_linear", "en_input_linear", "adaptive_rho",
              "adaptive_rho_min", "adaptive_rho_max", "adaptive_rho_enable_clipping", "sens_mode"):
        setattr(cp, k, getattr(p, k))
    # trajectories are (steps, dim) C-order == column-major dim x steps
    cp.x_min, cp.x_max, cp.u_min, cp.u_max = keep.d(p.x_min), keep.d(p.x_max), keep.d(p.u_min), keep.d(p.u_max)
    cp.n_state_cones, cp.Acx, cp.qcx, cp.cx = len(p.Acx), keep.i(p.Acx), keep.i(p.qcx), keep.d(p.cx)
    cp.n_input_cones, cp.Acu, cp.qcu, cp.cu = len(p.Acu), keep.i(p.Acu), keep.i(p.qcu), keep.d(p.cu)
    cp.n_state_lin = int(np.asarray(p.Alin_x).shape[0]) if np.asarray(p.Alin_x).size else 0
    cp.n_input_lin = int(np.asarray(p.Alin_u).shape[0]) if np.asarray(p.Alin_u).size else 0
    cp.Alin_x = keep.d(p.Alin_x, True) if cp.n_state_lin else None
    cp.blin_x = keep.d(p.blin_x) if cp.n_state_lin else None
    cp.Alin_u = keep.d(p.Alin_u, True) if cp.n_input_lin else None
    cp.blin_u = keep.d(p.blin_u) if cp.n_input_lin else None
    if p.sens_mode == 2:
        cp.dKinf, cp.dPinf = keep.d(p.dKinf, True), keep.d(p.dPinf, True)
        cp.dC1, cp.dC2 = keep.d(p.dC1, True), keep.d(p.dC2, True)
    return cp


def solve_batch(p, batch, impl: str = "ref", threads: int = 0) -> dict:
    """Cold-start solve of every problem in ``batch`` (problems.Batch).  Returns a dict with
    x (B,N,nx) f64, u (B,N-1,nu) f64, iter, status (int32), residuals (B,4), rho (B,)."""
    L = lib(impl)
    keep = _Keep()
    cp = c_problem(p, keep)
    B = batch.size
    cin = CBatchIn()
    cin.batch = B
    cin.x0, cin.Xref, cin.Uref = keep.f(batch.x0), keep.f(batch.Xref), keep.f(batch.Uref)
    cin.x_min, cin.x_max = keep.f(batch.x_min), keep.f(batch.x_max)
    cin.u_min, cin.u_max = keep.f(batch.u_min), keep.f(batch.u_max)
    out = dict(x=np.zeros((B, p.N, p.nx)), u=np.zeros((B, p.N - 1, p.nu)), iter=np.zeros(B, np.int32),
               status=np.zeros(B, np.int32), residuals=np.zeros((B, 4)), rho=np.zeros(B))
    co = CBatchOut()
    co.x, co.u = out["x"].ctypes.data_as(c_dp), out["u"].ctypes.data_as(c_dp)
    co.iter, co.status = out["iter"].ctypes.data_as(c_ip), out["status"].ctypes.data_as(c_ip)
    co.residuals, co.rho = out["residuals"].ctypes.data_as(c_dp), out["rho"].ctypes.data_as(c_dp)
    fn = getattr(L, _PREFIX[impl] + "_solve_batch")
    fn.restype = C.c_int
    rc = fn(C.byref(cp), C.byref(cin), C.byref(co), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"{impl}_solve_batch failed rc={rc}")
    return out


def get_cache(p, impl: str = "ref") -> dict:
    L = lib(impl)
    keep = _Keep()
    cp = c_problem(p, keep)
    nx, nu = p.nx, p.nu
    shapes = dict(Kinf=(nu, nx), Pinf=(nx, nx), Quu_inv=(nu, nu), AmBKt=(nx, nx), APf=(nx,), BPf=(nu,),
                  dKinf=(nu, nx), dPinf=(nx, nx), dC1=(nu, nu), dC2=(nx, nx))
    flat = {k: np.zeros(int(np.prod(s))) for k, s in shapes.items()}
    cc = CCacheOut()
    for k in shapes:
        setattr(cc, k, flat[k].ctypes.data_as(c_dp))
    fn = getattr(L, _PREFIX[impl] + "_get_cache")
    fn.restype = C.c_int
    rc = fn(C.byref(cp), C.byref(cc))
    if rc != 0:
        raise RuntimeError(f"{impl}_get_cache failed rc={rc}")
    return {k: flat[k].reshape(shapes[k], order="F") for k in shapes}


def codegen(p, out_dir, impl: str = "ref") -> None:
    """Run the REFERENCE code generator (tiny_codegen, codegen.cpp:68-80) on the solver this spec sets up (reference build only)."""
    L = lib(impl)
    keep = _Keep()
    cp = c_problem(p, keep)
    fn = getattr(L, _PREFIX[impl] + "_codegen")
    fn.restype = C.c_int
    rc = fn(C.byref(cp), str(out_dir).encode())
    if rc != 0:
        raise RuntimeError(f"{impl}_codegen failed rc={rc}")


class Session:
    """Warm-started single solver: the closed-loop pattern of quadrotor_hovering.cpp:73-93."""

    def __init__(self, p, impl: str = "ref"):
        self.L, self.p, self.pre = lib(impl), p, _PREFIX[impl]
        self._keep = _Keep()
        self._cp = c_problem(p, self._keep)
        fn = getattr(self.L, self.pre + "_session_create")
        fn.restype = C.c_void_p
        self.h = C.c_void_p(fn(C.byref(self._cp)))
        if not self.h:
            raise RuntimeError("session_create failed")

    def _call(self, name, *args):
        fn = getattr(self.L, f"{self.pre}_session_{name}")
        fn.restype = C.c_int
        return fn(self.h, *args)

    def set_x0(self, x0):
        a = np.ascontiguousarray(x0, np.float64)
        return self._call("set_x0", a.ctypes.data_as(c_dp))

    def set_x_ref(self, xr):
        a = np.ascontiguousarray(xr, np.float64)
        return self._call("set_x_ref", a.ctypes.data_as(c_dp))

    def set_u_ref(self, ur):
        a = np.ascontiguousarray(ur, np.float64)
        return self._call("set_u_ref", a.ctypes.data_as(c_dp))

    def solve(self) -> dict:
        p = self.p
        x, u, u0 = np.zeros((p.N, p.nx)), np.zeros((p.N - 1, p.nu)), np.zeros(p.nu)
        it, st = C.c_int(0), C.c_int(0)
        rc = self._call("solve", x.ctypes.data_as(c_dp), u.ctypes.data_as(c_dp), C.byref(it), C.byref(st),
                        u0.ctypes.data_as(c_dp))
        return dict(rc=rc, x=x, u=u, iter=it.value, status=st.value, work_u0=u0)

    def close(self):
        if self.h:
            fn = getattr(self.L, self.pre + "_session_destroy")
            fn.restype = None
            fn(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hardware_threads() -> int:
    return os.cpu_count() or 1
