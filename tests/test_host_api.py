"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the host C++
mirror of tiny_setup reproduces the reference cache, the MATLAB-class mirror expands arguments like
src/TinyMPC.m, and -- without a GPU -- every solve fails loudly instead of falling back to a CPU path."""
import ctypes
import importlib
import re
from pathlib import Path

import numpy as np
import pytest

import cases

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def tm():
    lib = ROOT / "tinympc-matlab_b200" / "libtinympc_b200.so"
    if not lib.exists():
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("tinympc-matlab_b200")


def test_library_exports_every_declared_symbol(tm):
    L = tm.capi.load()
    header = (ROOT / "include" / "tinympc_b200.h").read_text()
    declared = set(re.findall(r"\b(tinympc_cuda_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(tm.capi.EXPORTS), declared ^ set(tm.capi.EXPORTS)
    for name in sorted(declared) + tm.capi.HOST_EXPORTS:
        assert hasattr(L, name), f"{name} not exported by libtinympc_b200.so"
    assert b"sm_100a" in L.tinympc_cuda_version()


def test_no_cpu_fallback_without_gpu(tm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert tm.capi.load().tinympc_cuda_device_count() == 0
    with pytest.raises(tm.TinympcCudaError):
        tm.CudaSolver()
    s = tm.TinyMPC()
    p = cases.P.cartpole()
    s.setup_from_spec(p)
    s.set_x0([0.5, 0, 0, 0])
    with pytest.raises(tm.TinympcCudaError):
        s.solve()
    with pytest.raises(tm.TinympcCudaError):
        s.solve_batch(np.zeros((4, 3)))


@pytest.mark.parametrize("family", ["cartpole", "quadrotor", "rocket"])
def test_host_setup_reproduces_reference_cache(family, tm):
    """tiny_setup + tiny_precompute_and_set_cache of the host mirror vs the cache of the unmodified reference
    (tests/golden/cache_*.npz), including the double-rho quirk Q1."""
    p = dict(cartpole=cases.P.cartpole(), quadrotor=cases.P.quadrotor(adaptive=True), rocket=cases.P.rocket())[family]
    g = np.load(cases.GOLDEN / f"cache_{family}.npz")
    s = tm.TinyMPC().setup_from_spec(p)
    c = s.get_cache()
    for k in c:
        assert np.abs(c[k] - g[k]).max() <= 1e-8 * max(1.0, np.abs(g[k]).max()), k


def test_matlab_argument_expansion(tm):
    s = tm.TinyMPC()
    assert s._expand_bounds([], 3, 4, -1e17).shape == (3, 4) and s._expand_bounds([], 3, 4, -1e17)[0, 0] == -1e17
    assert np.all(s._expand_bounds(0.5, 2, 3, 0) == 0.5)
    assert np.array_equal(s._expand_bounds(np.array([[1.0], [2.0]]), 2, 3, 0), [[1, 1, 1], [2, 2, 2]])
    assert np.array_equal(s._expand_bounds(np.array([[1.0, 2.0]]), 2, 3, 0), [[1, 1, 1], [2, 2, 2]])
    assert np.array_equal(s._expand_matrix(np.array([1.0, 2.0]), 2, 3), [[1, 1, 1], [2, 2, 2]])
    full = np.arange(6.0).reshape(2, 3)
    assert s._expand_matrix(full, 2, 3) is not None and np.array_equal(s._expand_matrix(full, 2, 3), full)
    # defaults of the MATLAB class (src/TinyMPC.m:26-39)
    assert s.settings["abs_pri_tol"] == 1e-4 and s.settings["max_iter"] == 100 and not s.settings["en_input_bound"]


def test_setup_validation_and_settings_roundtrip(tm):
    s = tm.TinyMPC()
    with pytest.raises(AssertionError):
        s.setup(np.eye(3), np.ones((2, 1)), np.eye(3), np.eye(1), 5)
    with pytest.raises(RuntimeError):
        tm.TinyMPC().set_x0([0])
    p = cases.P.cartpole()
    s.setup(p.A, p.B, np.diag(p.Qdiag), np.diag(p.Rdiag), 20, rho=1.0, max_iter=77, bogus_option=3)
    assert s.settings["max_iter"] == 77
    s.set_bound_constraints([], [], -0.5, 0.5)
    assert s.settings["en_state_bound"] and s.settings["en_input_bound"] and s.u_max.shape == (1, 19) and s.x_min[0, 0] == -1e17
    s.update_settings(abs_pri_tol=1e-3, not_a_setting=1)
    d = np.zeros(4); i = (ctypes.c_int * 11)()
    s._L.tinympc_host_get_settings.argtypes = [ctypes.c_void_p, tm.capi.c_dp, ctypes.POINTER(ctypes.c_int * 11)]
    s._L.tinympc_host_get_settings(s._h, d.ctypes.data_as(tm.capi.c_dp), ctypes.byref(i))
    assert d[0] == 1e-3 and i[0] == 77 and i[2] == 1 and i[3] == 1
    with pytest.raises(NotImplementedError):
        s.codegen("out")
