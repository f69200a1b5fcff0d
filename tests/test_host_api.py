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
    import tempfile
    from pathlib import Path
    with tempfile.TemporaryDirectory() as td:       # codegen is part of the surface now (tests/test_codegen.py checks the contents)
        s.codegen(Path(td) / "out")
        assert (Path(td) / "out" / "src" / "tiny_data.cpp").exists() and (Path(td) / "out" / "tinympc" / "tiny_b200_family.h").exists()


def test_bench_flop_model_matches_the_survey_table():
    """bench.py's roofline numerator: SURVEY 8d box-only counts, plus the per-family and adaptive-rho terms of the same table."""
    import importlib
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    bench = importlib.import_module("bench")
    P = importlib.import_module("tinympc-matlab_b200.problems")
    q, c, r, qa = P.quadrotor(), P.cartpole(), P.rocket(), P.quadrotor(adaptive=True)
    assert bench.flops_per_iter(q.nx, q.nu, q.N) == 12240 and bench.flops_per_iter(q.nx, q.nu, q.N, q) == 12240
    assert bench.flops_per_iter(c.nx, c.nu, c.N, c) == 3828
    assert bench.flops_per_iter(r.nx, r.nu, r.N) == 4500
    # + 6 per element of the 4 enabled families, 20 per (step, cone), 4 dim + 2 per (step, row)
    assert bench.flops_per_iter(r.nx, r.nu, r.N, r) == 4500 + 2 * 6 * (60 + 27) + 20 * (10 + 9) + 26 * 10 + 14 * 9
    assert bench.flops_per_iter(qa.nx, qa.nu, qa.N, qa) == 12240 + (9 * (576 + 192) + 288 + 12 * 156) // 5
    assert bench.bytes_per_solve(12, 4, 10) == 1304 and bench.bytes_per_solve(4, 1, 20) == 816 and bench.bytes_per_solve(6, 3, 10) == 728


def test_hybrid_layout_plan_fits_the_sm():
    """build.py's planner of the hybrid state layout (tmpc_tpp3.cuh): tensor-memory columns, shared memory and registers of the
    chosen (warps, t columns in tensor memory) fit one SM for every compiled shape."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("tmpc_build", Path(__file__).resolve().parents[1] / "tinympc-matlab_b200" / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.plan_hybrid(12, 4, 10) == (12, 6)
    for nx, nu, N in [(12, 4, 10), (4, 1, 20), (4, 1, 10), (6, 3, 10)]:
        warps, ttm = b.plan_hybrid(nx, nu, N)
        assert warps % 4 == 0 and 0 <= ttm <= N - 1
        assert (warps // 4) * ((N - 2) + ttm) * nx <= 512
        smem = (3 * nu * (N - 1) + (N - 1 - ttm) * nx) * warps * 32 * 4 + 1024 + b.pack_elems(nx, nu, N) * 4
        assert smem <= 227 * 1024
    names = [b.name_of(i) for i in b.default_instances()]
    assert len(names) == len(set(names)), "two kernel instances with the same name would be one symbol at link time"


def test_generated_kernel_instances_are_distinct_template_instantiations():
    """Every generated translation unit must instantiate a DIFFERENT kernel template: two variants with the same template
    arguments would be one weak symbol at link time, and the `variant` option would silently run the other one's code."""
    import importlib.util
    import re
    from pathlib import Path
    pkg = Path(__file__).resolve().parents[1] / "tinympc-matlab_b200"
    spec = importlib.util.spec_from_file_location("tmpc_build", pkg / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.gen_sources(b.default_instances())          # idempotent: writes csrc/gen/*.cu only where the text changed
    gen = pkg / "csrc" / "gen"
    seen = {}
    for f in sorted(list(gen.glob("tpp*.cu")) + list(gen.glob("gpp*.cu"))):
        m = re.search(r"using \w+ = ((?:Tpp\d?|Gpp)Cfg<[^;]*>);", f.read_text())
        assert m, f.name
        assert m.group(1) not in seen, f"{f.name} and {seen[m.group(1)]} instantiate the same kernel"
        seen[m.group(1)] = f.name
    assert len(seen) > 40


def test_instance_list_covers_the_dispatch_rules():
    """build.py's instance list is what tmpc_capi.cu::find_kernel dispatches over: every BASELINE shape must have (a) an fp32
    incremental-form box instance as variant 0, the quadrotor ones with the impulse-response backward pass (CONV = auto) and the
    costate recursion kept as variant 7, (b) a lane-group fp64 instance (box; adaptive rho for the quadrotor) ahead of the
    thread-per-problem fp64 instances, which move to variant 6 except for per-problem bounds and the cone family."""
    import importlib.util
    from pathlib import Path
    pkg = Path(__file__).resolve().parents[1] / "tinympc-matlab_b200"
    spec = importlib.util.spec_from_file_location("tmpc_build", pkg / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    inst = b.default_instances()
    names = [b.name_of(i) for i in inst]
    shapes = [(12, 4, 10), (4, 1, 20), (4, 1, 10), (6, 3, 10)]
    for (nx, nu, N) in shapes:
        assert any(i["gen"] == 3 and (i["nx"], i["nu"], i["N"]) == (nx, nu, N) and i["feat"] == b.BOX and i["variant"] == 0 for i in inst)
        g = [k for k, i in enumerate(inst) if i["gen"] == 5 and (i["nx"], i["nu"], i["N"]) == (nx, nu, N) and i["feat"] == b.BOX]
        t = [k for k, i in enumerate(inst) if i["gen"] == 2 and i["bits"] == 64 and (i["nx"], i["nu"], i["N"]) == (nx, nu, N)
             and i["feat"] == b.BOX and not i["ppb"]]
        assert len(g) == 1 and inst[g[0]]["variant"] == 0 and inst[g[0]]["gs"] >= nx + nu
        assert t and all(inst[k]["variant"] == 6 for k in t), "shared-bounds fp64 box batches belong to the lane-group kernel"
        assert any(i["gen"] == 2 and i["bits"] == 64 and i["ppb"] and i["variant"] == 0 and (i["nx"], i["nu"], i["N"]) == (nx, nu, N) for i in inst)
    assert any(i["gen"] == 5 and i["feat"] == b.ADP and (i["nx"], i["nu"]) == (12, 4) and i["variant"] == 0 for i in inst)
    assert all(i["variant"] == 6 for i in inst if i["gen"] == 2 and i["bits"] == 64 and i["feat"] == b.ADP)
    q = [i for i in inst if i["gen"] == 3 and (i["nx"], i["nu"]) == (12, 4) and i["feat"] == b.BOX]
    assert all(i["conv"] == -1 for i in q if i["variant"] in (0, 9)) and any(i["conv"] == 0 and i["variant"] == 7 for i in q)
    assert len(names) == len(set(names))


def test_compact_pipeline_upload_plan():
    """tinympc_cuda_plan_compact_chunks (tmpc_capi.cu plan_compact_chunks): the upload plan of the compact streamed host pipeline.
    Chunks tile the shard, every inner boundary falls on a multiple of 32 problems (no 128-byte line holds inputs of two chunks), the
    auto plan starts with 1/64 of the shard and doubles, and the chunks before the first ordered one cover at most n / div problems."""
    import ctypes as C
    tm = importlib.import_module("tinympc-matlab_b200")
    L = tm.capi.load()
    buf = (C.c_int * 80)()
    fo = C.c_int()
    for n in (65536, 70001, 300007, 1 << 20, (1 << 20) + 17, 1 << 24):
        for chunks in (0, 2, 3, 7, 16, 64, 1000):
            for div in (2, 4, 8):
                nch = L.tinympc_cuda_plan_compact_chunks(n, chunks, div, buf, 80, C.byref(fo))
                b = list(buf[:nch + 1])
                assert nch >= 1 and b[0] == 0 and b[-1] == n and all(x < y for x, y in zip(b, b[1:])), (n, chunks, b)
                assert all(x % 32 == 0 for x in b[:-1]), (n, chunks, b)
                assert 1 <= fo.value <= max(1, nch - 1) and (fo.value == 1 or b[fo.value] <= n // div), (n, chunks, div, fo.value, b)
                if fo.value + 1 < nch:
                    assert b[fo.value + 1] > n // div          # the next boundary would exceed the unordered share
                if chunks == 0:
                    sizes = [y - x for x, y in zip(b, b[1:])]
                    assert nch == 7 and sizes[0] == sizes[1] and sizes[0] <= -(-n // 64) + 31, (n, sizes)
                    assert all(abs(s2 - 2 * s1) <= 0 for s1, s2 in zip(sizes[1:-2], sizes[2:-1])), (n, sizes)   # doubling (the last takes the rest)
                elif chunks >= 2:
                    assert nch <= min(chunks, 64)
    assert L.tinympc_cuda_plan_compact_chunks(1 << 20, 0, 2, buf, 4, C.byref(fo)) == -1      # buffer too small
    n = 1 << 20
    assert L.tinympc_cuda_plan_compact_chunks(n, 0, 2, buf, 80, C.byref(fo)) == 7 and buf[fo.value] == n // 2     # the second half is ordered
    assert L.tinympc_cuda_plan_compact_chunks(n, 0, 4, buf, 80, C.byref(fo)) == 7 and buf[fo.value] == n // 4
