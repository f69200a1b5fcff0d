"""The MEX gateway (tinympc-matlab_b200/matlab/bindings.cpp) built against a stub mex.h: command dispatch, error ids and
argument checks on the CPU; the cartpole one-solve example and solve_batch through the gateway on the GPU."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "tinympc-matlab_b200"
EXE = ROOT / "tests" / "stub_mex" / "mex_driver"


@pytest.fixture(scope="module")
def driver():
    lib = PKG / "libtinympc_b200.so"
    if not lib.exists():
        import __graft_entry__
        __graft_entry__.build()
    srcs = [ROOT / "tests" / "stub_mex" / "mex_driver.cpp", PKG / "matlab" / "bindings.cpp"]
    if not EXE.exists() or EXE.stat().st_mtime < max(p.stat().st_mtime for p in srcs + [lib]):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "tests" / "stub_mex"), "-I", str(PKG / "csrc" / "host"), "-I", str(ROOT / "include"),
                               *map(str, srcs), "-o", str(EXE), f"-L{PKG}", "-ltinympc_b200", f"-Wl,-rpath,{PKG}"])
    return EXE


def run(exe, mode):
    out = subprocess.run([str(exe), mode], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout.splitlines()


def test_gateway_dispatch_and_errors_on_host(driver):
    lines = run(driver, "host")
    assert "setup_status 0" in lines
    assert "error_id TinyMPC:InvalidFunction" in lines and "error_id TinyMPC:InvalidInput" in lines
    assert lines[-1] == "host_only_done"


@pytest.mark.gpu
def test_gateway_cartpole_example_and_solve_batch_on_gpu(driver):
    lines = run(driver, "gpu")
    _, _, g = cases.load("G2_cartpole_ubound")
    assert "solve_ret 0" in lines and "iter 51 status 1" in lines
    u = np.array([float(t) for t in next(l for l in lines if l.startswith("u ")).split()[1:]])
    assert np.abs(u - g["u"][0, :, 0]).max() < 1e-9
    b = next(l for l in lines if l.startswith("batch_iter")).split()
    assert b[1:4] == ["51", "51", "51"] and b[5] == "3" and abs(float(b[7]) - 0.5) < 1e-6
    # sessions through the gateway: same closed loop as the single solver driven by set_x0 / solve
    it2 = next(l for l in lines if l.startswith("loop_iter2")).split()[1]
    assert "session_iter 51 51 51" in lines
    assert f"session_iter2 {it2} {it2} {it2}" in lines, [l for l in lines if l.startswith("session")]
    assert float(next(l for l in lines if l.startswith("session_x0_err")).split()[1]) < 1e-12
    assert "session_dims 3" in lines
    assert lines.count("session_error_id TinyMPC:NotInitialized") == 1 and "session_error_id TinyMPC:InvalidInput" in lines
    assert lines[-1] == "error_id TinyMPC:NotInitialized"
