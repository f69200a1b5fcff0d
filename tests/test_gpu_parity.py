"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI,
against the golden vectors produced by the unmodified reference and against the CPU oracle on
seeded random batches.

Tolerances (BASELINE.json north_star): identical status codes and iteration counts, states and
controls within 1e-4 absolute in fp32.  The fp64 "parity mode" of the same kernels must reproduce
the reference's iteration counts exactly and x/u to float32 output rounding (1e-6 relative).  In fp32 a termination test can flip by
one iteration when a residual lands within rounding distance of the tolerance; those
problems are counted and bounded, and x/u is compared on the problems whose counts agree.
"""
import importlib

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

X_TOL_F32 = 1e-4
X_TOL_F64 = 1e-6   # fp64 arithmetic, but the C ABI returns float32 trajectories


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return importlib.import_module("tinympc-matlab_b200.capi")


def solve_gpu(capi, oracle_mod, p, b, precision, variant=0, mixed=0.0, fixer_sms=0):
    s = capi.CudaSolver()
    s.set_option("precision", precision)
    s.set_option("variant", variant)
    s.set_option("mixed", mixed)
    s.set_option("fixer_sms", fixer_sms)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    r = s.solve_batch(b.x0, b.Xref, b.Uref, b.x_min, b.x_max, b.u_min, b.u_max)
    r["kernel"] = s.last_kernel
    r["marked"] = s.last_marked
    s.close()
    return r


def compare(r, g, precision, name, max_flip_frac=0.0, hard_tol=X_TOL_F32):
    """north_star: identical status codes and iteration counts; states / controls within 1e-4 ABSOLUTE in fp32.
    max_flip_frac > 0 only for the plain-fp32 mode (a termination test on rounded values can flip by one check interval,
    SURVEY H1); the exact-count mode and fp64 are held to 0.  The absolute bar is asserted on every problem that converged with
    the reference's iteration count, and -- since the backward pass of the fp32 kernels runs in impulse-response form -- on the
    problems that run into max_iter as well (hard_tol = 1e-4).  Only the direct-form fp32 kernels kept as A/B variants and the fp32
    adaptive-rho kernel (neither is a default of the benchmark) pass hard_tol = 1e-3 for the problems stopped at max_iter, whose
    steps stay large and never damp a rounding difference."""
    B = len(g["iter"])
    same = (r["iter"] == g["iter"]) & (r["status"] == g["status"])
    flips = int((~same).sum())
    assert flips <= max_flip_frac * B, f"{name}: {flips}/{B} iteration/status mismatches (kernel {r['kernel']})"
    if flips:   # a flipped problem stopped a few iterations early/late: its solution is still the same to ~10 x tol
        scale_f = max(1.0, float(np.abs(g["x"]).max()))
        assert np.abs(r["x"][~same] - g["x"][~same]).max() <= 1e-1 * scale_f, name
    if precision == 64:      # fp64 arithmetic behind the float32 batch ABI: output rounding only
        scale = max(1.0, float(np.abs(g["x"]).max()), float(np.abs(g["u"]).max()))
        dx = np.abs(r["x"][same] - g["x"][same]).max() if same.any() else 0.0
        du = np.abs(r["u"][same] - g["u"][same]).max() if same.any() else 0.0
        assert dx <= X_TOL_F64 * scale and du <= X_TOL_F64 * scale, f"{name}: dx={dx:.3e} du={du:.3e} (kernel {r['kernel']})"
        return flips, dx, du
    conv, hard = same & (g["status"] == 1), same & (g["status"] != 1)
    err = np.maximum(np.abs(r["x"] - g["x"]).reshape(B, -1).max(1), np.abs(r["u"] - g["u"]).reshape(B, -1).max(1))
    e_conv = float(err[conv].max()) if conv.any() else 0.0
    e_hard = float(err[hard].max()) if hard.any() else 0.0
    dx = np.abs(r["x"][same] - g["x"][same]).max() if same.any() else 0.0
    du = np.abs(r["u"][same] - g["u"][same]).max() if same.any() else 0.0
    print(f"\n[abs] {name}: max|dxu| converged {e_conv:.2e} ({int(conv.sum())}), at max_iter {e_hard:.2e} ({int(hard.sum())})")
    assert e_conv <= X_TOL_F32, f"{name}: {e_conv:.3e} > 1e-4 absolute on a converged problem (kernel {r['kernel']})"
    assert e_hard <= hard_tol, f"{name}: {e_hard:.3e} on a problem stopped at max_iter (kernel {r['kernel']})"
    return flips, dx, du


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_golden_fp64(name, capi, oracle_mod):
    p, b, g = cases.load(name)
    r = solve_gpu(capi, oracle_mod, p, b, 64)
    compare(r, g, 64, name)
    if "rho" in g:
        assert np.abs(r["rho"] - g["rho"]).max() < 1e-5


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_golden_fp32(name, capi, oracle_mod):
    p, b, g = cases.load(name)
    r = solve_gpu(capi, oracle_mod, p, b, 32)
    # plain fp32: single-problem cases sit exactly on a tolerance by construction (G2: dual residual 9.99986e-5 vs 1e-4) and may
    # flip; the exact-count mode below may not
    B = len(g["iter"])
    compare(r, g, 32, name, max_flip_frac=1.0 if B == 1 else 0.08, hard_tol=1e-3 if p.adaptive_rho else X_TOL_F32)


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_golden_exact_mode(name, capi, oracle_mod, problems):
    """The default mode of the benchmark (option "mixed" = the family's band): every golden case, G2 included, with the
    reference's iteration count and status, x / u within 1e-4 absolute."""
    p, b, g = cases.load(name)
    prec = problems.exact_precision(p)     # adaptive rho: the parity-exact mode is plain fp64 (tmpc_gpp.cuh)
    r = solve_gpu(capi, oracle_mod, p, b, prec, mixed=problems.exact_band(p) if prec == 32 else 0.0)
    compare(r, g, prec, name)


# measured fp32 iteration-count flip rates (one check interval early/late).  Box-constrained batches run the incremental-form
# kernel (tmpc_tpp3.cuh): cartpole 0.01 %, quadrotor 0.01-0.04 % (the direct form, variant 5, has 1-2 %); adaptive quadrotor
# 3 % and rocket 19 % (tol_dua 1e-4 on thrusts of magnitude 100 is at fp32 resolution) still run the direct form.
FLIP_BOUND = {"cartpole": 0.001, "quadrotor": 0.0015, "quadrotor_adaptive": 0.04, "rocket": 0.002}
FLIP_BOUND_DIRECT = {"cartpole": 0.005, "quadrotor": 0.04}


@pytest.mark.parametrize("family,scale", [("cartpole", 0.3), ("cartpole", 1.0), ("quadrotor", 0.3), ("quadrotor", 1.0),
                                          ("rocket", 1.0), ("quadrotor_adaptive", 1.0)])
@pytest.mark.parametrize("precision", [32, 64])
def test_random_batch_vs_oracle(family, scale, precision, capi, oracle_mod, problems):
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket,
             quadrotor_adaptive=lambda: problems.quadrotor(adaptive=True))[family]()
    B = 10000
    b = problems.make_batch(p, B, scale, seed=2024)
    impl = "ref" if oracle_mod.available("ref") else "port"
    g = oracle_mod.solve_batch(p, b, impl)
    r = solve_gpu(capi, oracle_mod, p, b, precision)
    flips, dx, du = compare(r, g, precision, f"{family}@{scale}", max_flip_frac=0.0 if precision == 64 else FLIP_BOUND[family],
                            hard_tol=1e-3 if family == "quadrotor_adaptive" else X_TOL_F32)
    print(f"\n[parity] {family} s={scale} fp{precision}: {flips}/{B} count flips, max|dx|={dx:.2e} max|du|={du:.2e} kernel={r['kernel']}")


@pytest.mark.parametrize("family,scale", [("cartpole", 1.0), ("quadrotor", 0.3), ("quadrotor", 1.0)])
def test_direct_form_variant(family, scale, capi, oracle_mod, problems):
    """variant 5 = the direct-form fp32 kernel (tmpc_tpp2.cuh): same solutions, more count flips than the default"""
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor)[family]()
    B = 10000
    b = problems.make_batch(p, B, scale, seed=2024)
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    r = solve_gpu(capi, oracle_mod, p, b, 32, variant=5)
    assert r["kernel"].startswith("tpp2_f32"), r["kernel"]
    flips, dx, du = compare(r, g, 32, f"{family}@{scale} direct", max_flip_frac=FLIP_BOUND_DIRECT[family], hard_tol=1e-3)
    print(f"\n[parity] {family} s={scale} fp32 direct form: {flips}/{B} count flips, max|dx|={dx:.2e} max|du|={du:.2e} kernel={r['kernel']}")


@pytest.mark.parametrize("fixer_sms", [0, -1])
@pytest.mark.parametrize("family,scale,band,max_flips", [("cartpole", 1.0, 0.003, 0), ("cartpole", 0.3, 0.003, 0), ("quadrotor", 0.3, 0.002, 0),
                                                         ("quadrotor", 1.0, 0.002, 0), ("rocket", 1.0, 0.003, 0), ("quadrotor_adaptive", 1.0, 0.3, 0)])
def test_mixed_mode_exact_counts(family, scale, band, max_flips, fixer_sms, capi, oracle_mod, problems):
    """option "mixed": the fp32 pass stops every problem whose termination decision lies within the relative band of the
    tolerances, an fp64 pass re-solves exactly those -> the reference's iteration counts and status codes.  Measured on 2^18
    problems (profiles/tools/mixed_sweep.py): with the incremental-form kernel a band of 0.3 % leaves 0 (cartpole, easy quadrotor)
    to 2 (hard quadrotor) mismatches at 1-2 % re-solved; the direct-form kernels (rocket, adaptive rho) need a 30 % band."""
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket,
             quadrotor_adaptive=lambda: problems.quadrotor(adaptive=True))[family]()
    B = 10000
    b = problems.make_batch(p, B, scale, seed=2024)
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    # fixer_sms 0: the concurrent producer / consumer pair ("fp32|fp64"); -1: the sequential two-pass form ("fp32+fp64")
    r = solve_gpu(capi, oracle_mod, p, b, 32, mixed=band, fixer_sms=fixer_sms)
    assert ("|" if fixer_sms >= 0 else "+") in r["kernel"] and 0 < r["marked"] < B, (r["kernel"], r["marked"])
    assert not (r["status"] & 0x100).any(), "a marked problem was not re-solved"
    flips, dx, du = compare(r, g, 32, f"{family}@{scale} mixed", max_flip_frac=max_flips / B, hard_tol=1e-3 if family == "quadrotor_adaptive" else X_TOL_F32)
    print(f"\n[parity] {family} s={scale} mixed band={band}: {flips}/{B} count flips, {r['marked']} re-solved in fp64, max|dx|={dx:.2e} "
          f"max|du|={du:.2e} kernel={r['kernel']}")


def test_mixed_mode_host_and_device_paths_agree(capi, oracle_mod, problems):
    """the chunked host pipeline (several fp32 + compaction + fp64 launches on three streams) returns what one launch does"""
    p = problems.quadrotor()
    b = problems.make_batch(p, 40000, 1.0, seed=7)
    s = capi.CudaSolver()
    s.set_option("mixed", 0.01)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    s.set_option("chunks", 1)
    r1 = s.solve_batch(b.x0, b.Xref, b.Uref)
    m1 = s.last_marked
    s.set_option("chunks", 7)
    r7 = s.solve_batch(b.x0, b.Xref, b.Uref)
    m7 = s.last_marked
    s.close()
    assert m1 == m7 and m1 > 0
    assert np.array_equal(r1["iter"], r7["iter"]) and np.array_equal(r1["status"], r7["status"])
    assert np.array_equal(r1["x"], r7["x"]) and np.array_equal(r1["u"], r7["u"])


@pytest.mark.parametrize("family,B,chunks,precision", [("quadrotor", 100003, 0, 32), ("quadrotor", 70000, 7, 32), ("cartpole", 131072, 0, 32),
                                                       ("rocket", 50001, 3, 32), ("quadrotor_adaptive", 40000, 2, 32), ("quadrotor", 40000, 3, 64)])
def test_streamed_pipeline_matches_chunked_launches(family, B, chunks, precision, capi, oracle_mod, problems):
    """option "streamed" (default): one persistent launch consumes the batch while the H2D chunks are still arriving and
    hands results back chunk by chunk (arrival watermark + per-chunk completion counters, tmpc_capi.cu run_shard_streamed).
    Must return bit-for-bit what the one-launch-per-chunk pipeline returns, for ragged sizes and every kernel family."""
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket,
             quadrotor_adaptive=lambda: problems.quadrotor(adaptive=True))[family]()
    b = problems.make_batch(p, B, 1.0, seed=11)
    s = capi.CudaSolver()
    s.set_option("precision", precision)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    s.set_option("streamed", 0)
    s.set_option("chunks", 1)
    r0 = s.solve_batch(b.x0, b.Xref, b.Uref)
    assert s.last_timing()["chunks"] == 1
    s.set_option("streamed", 1)
    s.set_option("chunks", chunks)
    for _ in range(2):   # twice: the control words are reused
        r1 = s.solve_batch(b.x0, b.Xref, b.Uref)
        assert s.last_timing()["chunks"] >= 2, "the streamed path was not taken"
        for k in ("iter", "status", "x", "u", "residuals", "rho"):
            assert np.array_equal(r0[k], r1[k]), f"{family}: streamed pipeline differs in {k}"
    s.close()


@pytest.mark.parametrize("family,B", [("quadrotor", 100003), ("rocket", 50001)])
def test_streamed_exact_mode_matches_the_two_pass_form(family, B, capi, oracle_mod, problems):
    """Exact-count mode through the streamed host pipeline: the fp32 producer consumes the batch while it is arriving and queues
    what it cannot decide, the fp64 consumer runs concurrently on the SMs left free, both count completions per chunk for the
    D2H stream.  Same iteration counts, statuses and solutions as the sequential form (fp32 pass, compaction, fp64 pass) on one
    chunk, bit for bit."""
    p = dict(quadrotor=problems.quadrotor, rocket=problems.rocket)[family]()
    b = problems.make_batch(p, B, 1.0, seed=13)
    s = capi.CudaSolver()
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    s.set_option("mixed", problems.exact_band(p))
    s.set_option("fixer_sms", -1)
    s.set_option("chunks", 1)
    r0 = s.solve_batch(b.x0, b.Xref, b.Uref)
    m0 = s.last_marked
    assert "+" in s.last_kernel
    s.set_option("fixer_sms", 0)
    s.set_option("chunks", 0)
    for _ in range(2):
        r1 = s.solve_batch(b.x0, b.Xref, b.Uref)
        assert "|" in s.last_kernel and s.last_timing()["chunks"] >= 2, (s.last_kernel, s.last_timing())
        assert s.last_marked == m0 > 0
        for k in ("iter", "status", "x", "u", "residuals", "rho"):
            assert np.array_equal(r0[k], r1[k]), f"{family}: streamed exact mode differs in {k}"
    s.close()


def test_streamed_pipeline_per_problem_bounds(capi, oracle_mod, problems):
    name = "batch_quadrotor_perproblem_bounds"
    p, b0, g = cases.load(name)
    reps = 40000 // b0.size + 1
    tile = lambda a: None if a is None else np.ascontiguousarray(np.concatenate([a] * reps, axis=0))
    s = capi.CudaSolver()
    s.set_option("precision", 64)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    s.set_option("chunks", 3)
    r = s.solve_batch(tile(b0.x0), tile(b0.Xref), tile(b0.Uref), tile(b0.x_min), tile(b0.x_max), tile(b0.u_min), tile(b0.u_max))
    assert s.last_timing()["chunks"] == 3
    s.close()
    n = b0.size
    for k in range(0, reps * n, n * 7):
        assert np.array_equal(r["iter"][k:k + n], g["iter"]) and np.array_equal(r["status"][k:k + n], g["status"])
        assert np.abs(r["x"][k:k + n] - g["x"]).max() <= X_TOL_F64 * max(1.0, float(np.abs(g["x"]).max()))


@pytest.mark.parametrize("family", ["quadrotor", "cartpole", "rocket"])
def test_full_size_batch_is_order_independent_and_matches_the_oracle_on_a_sample(family, capi, oracle_mod, problems):
    """BASELINE size (2^20 problems): properties that do not depend on the size the oracle can afford.
    (1) Every problem is solved independently of which lane, warp and refill batch picks it up: solving a permuted batch returns
        the permuted results BIT FOR BIT (this is the check of the lane-refill / batched-refill / work-counter machinery).
    (2) A strided sample of 2048 problems of the big batch matches the reference on the same inputs."""
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket)[family]()
    B = 1 << 20
    b = problems.make_batch(p, B, 1.0, seed=77)
    perm = np.random.default_rng(5).permutation(B)
    take = lambda a: None if a is None else np.ascontiguousarray(a[perm])
    bp = problems.Batch(take(b.x0), take(b.Xref), take(b.Uref))
    r = solve_gpu(capi, oracle_mod, p, b, 32)
    rp = solve_gpu(capi, oracle_mod, p, bp, 32)
    for k in ("iter", "status", "x", "u"):
        assert np.array_equal(r[k][perm], rp[k]), f"{family}: field {k} depends on the position of the problem in the batch"
    idx = np.arange(0, B, B // 2048)
    sub = problems.Batch(b.x0[idx], None if b.Xref is None else b.Xref[idx], None if b.Uref is None else b.Uref[idx])
    g = oracle_mod.solve_batch(p, sub, "ref" if oracle_mod.available("ref") else "port")
    rs = {k: r[k][idx] for k in ("iter", "status", "x", "u")}
    rs["kernel"] = r["kernel"]
    flips, dx, du = compare(rs, g, 32, f"{family} 2^20 sample", max_flip_frac=FLIP_BOUND[family])
    print(f"\n[parity] {family} 2^20 problems: permutation-invariant bit for bit; sample of {len(idx)}: {flips} count flips, "
          f"max|dx|={dx:.2e} max|du|={du:.2e} kernel={r['kernel']}")


@pytest.mark.parametrize("family,streamed", [("quadrotor", 1), ("quadrotor", 0), ("rocket", 1)])
def test_multi_device_split_is_bit_identical(family, streamed, capi, oracle_mod, problems):
    """the library splits a host batch contiguously by problem index over its device contexts, one host thread per
    context, no collective.  The split must not change a single bit of any problem's result.  With one GPU the two contexts sit
    on the same device (the split, the thread fan-out and the remainder handling are exercised all the same); with two or more
    GPUs the batch really spans devices 0 and 1."""
    import torch
    p = dict(quadrotor=problems.quadrotor, rocket=problems.rocket)[family]()
    B = 100003                                    # odd on purpose: the last shard takes the remainder
    b = problems.make_batch(p, B, 1.0, seed=31)
    fam = cases.family_from_spec(p, oracle_mod.get_cache(p, "port"))
    res = []
    for devices in ([0], [0, 1] if torch.cuda.device_count() >= 2 else [0, 0], [0, 0, 0]):
        s = capi.CudaSolver(devices=devices)
        s.set_option("streamed", streamed)
        s.set_family(fam)
        res.append(s.solve_batch(b.x0, b.Xref, b.Uref))
        s.close()
    for r in res[1:]:
        for k in ("iter", "status", "x", "u"):
            assert np.array_equal(res[0][k], r[k]), f"{family}: field {k} changes with the device split"


@pytest.mark.parametrize("streamed", [1, 0])
def test_multi_device_fanout_stress(streamed, capi, oracle_mod, problems):
    """Regression for the round-1 abort (GPUTEST_r01: SIGABRT in test_multi_device_split_is_bit_identical): the per-device worker
    threads of tinympc_cuda_solve_batch wrote the solver's std::string fields concurrently.  50 fan-outs over three contexts,
    alternating between kernels with names of different length (the reallocation that corrupted the heap), every result compared
    bit for bit with the single-context run."""
    q, c = problems.quadrotor(), problems.cartpole()
    bq, bc = problems.make_batch(q, 6007, 1.0, seed=3), problems.make_batch(c, 6007, 1.0, seed=3)
    fq, fc = (cases.family_from_spec(p, oracle_mod.get_cache(p, "port")) for p in (q, c))
    one = capi.CudaSolver(devices=[0])
    ref = {}
    for name, fam, b in (("q", fq, bq), ("c", fc, bc)):
        one.set_family(fam)
        ref[name] = one.solve_batch(b.x0, b.Xref, b.Uref)
    one.close()
    s = capi.CudaSolver(devices=[0, 0, 0])
    s.set_option("streamed", streamed)
    s.set_option("chunks", 2)
    for k in range(50):
        name, fam, b = (("q", fq, bq), ("c", fc, bc))[k % 2]
        s.set_family(fam)
        if k % 10 == 9:
            s.set_option("precision", 64)      # "tpp2_f64_..." / wpp names in between
        r = s.solve_batch(b.x0, b.Xref, b.Uref)
        if k % 10 == 9:
            s.set_option("precision", 32)
            continue
        assert s.last_kernel.startswith("tpp3_"), s.last_kernel
        for f in ("iter", "status", "x", "u"):
            assert np.array_equal(ref[name][f], r[f]), f"pass {k}: field {f} differs from the single-context run"
    assert s.launch_count >= 150
    s.close()


def test_session_outliving_its_solver_fails_cleanly(capi, oracle_mod, problems):
    """a session kept a raw pointer to its solver.  Destroying the solver now orphans its sessions: later calls
    return TINYMPC_CUDA_ENOTREADY (5) instead of touching freed memory, and session_destroy still releases the handle."""
    p = problems.cartpole()
    s = capi.CudaSolver()
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    ss = s.session(8)
    ss.set_x0(np.zeros((8, p.nx)))
    ss.solve()
    L, h = s.L, ss.h
    s.close()
    assert L.tinympc_cuda_session_solve(h) == 5
    assert L.tinympc_cuda_session_step(h, 0) == 5
    out = np.zeros(8)
    assert L.tinympc_cuda_session_read(h, b"iter", out.ctypes.data_as(capi.c_dp)) == 5
    assert L.tinympc_cuda_session_destroy(h) == 0
    ss.h = capi.C.c_void_p()


@pytest.mark.parametrize("pipeline", ["streamed", "chunked"])
@pytest.mark.parametrize("mixed", [0.0, 0.003])
@pytest.mark.parametrize("B", [1000, 70001])
def test_compact_io_matches_full_trajectories(B, mixed, pipeline, capi, oracle_mod, problems):
    """tinympc_cuda_batch_in::xref_const (one reference state per problem, standing for every column of the horizon) and
    tinympc_cuda_batch_out::u0 (first control only): same iteration counts, statuses and first controls as the full-trajectory
    call on the replicated reference, through the host entry and the device entry.  "streamed" (default): the kernels read the
    compact reference in place and a shard of >= 2^16 problems goes through one launch chain behind an arrival watermark
    (run_shard_compact_streamed); "chunked": the reference is replicated on the device first, one launch per chunk."""
    import torch
    p = problems.quadrotor()
    b = problems.make_batch(p, B, 1.0, seed=21)
    assert (b.Xref == b.Xref[:, :1]).all() and not b.Uref.any()
    s = capi.CudaSolver()
    s.set_option("mixed", mixed)
    s.set_option("compact_streamed", 1 if pipeline == "streamed" else 0)
    s.set_option("compact_in_kernel", 1 if pipeline == "streamed" else 0)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    full = s.solve_batch(b.x0, b.Xref, None)
    xc = np.ascontiguousarray(b.Xref[:, 0, :])
    for _ in range(2):   # twice: the control words are reused
        c = s.solve_batch(b.x0, xref_const=xc, compact_out=True)
        assert (s.last_timing()["chunks"] >= 2) == (pipeline == "streamed" and B >= 65536), s.last_timing()
        if mixed > 0:
            assert s.last_marked > 0 and "+" in s.last_kernel, (s.last_marked, s.last_kernel)
        assert set(c) == {"u0", "iter", "status"}
        assert np.array_equal(c["iter"], full["iter"]) and np.array_equal(c["status"], full["status"])
        assert np.array_equal(c["u0"], full["u"][:, 0, :])
    # pinned result arrays: in the exact-count mode of the streamed form the results of the first pass are copied back under the
    # fp64 pass, whose results a kernel then writes over them through the device alias of the host arrays (compact_early_d2h)
    pinned = dict(u0=torch.empty((B, p.nu)).pin_memory().numpy(), iter=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                  status=torch.empty(B, dtype=torch.int32).pin_memory().numpy())
    for early in (1, 0, 1):
        s.set_option("compact_early_d2h", early)
        for a in pinned.values():
            a.fill(-7)
        s.solve_batch(b.x0, xref_const=xc, out=pinned, compact_out=True)
        assert np.array_equal(pinned["iter"], full["iter"]) and np.array_equal(pinned["status"], full["status"]), f"early={early}"
        assert np.array_equal(pinned["u0"], full["u"][:, 0, :]), f"early={early}"
    # mixed forms: compact input with full output, full input with compact output
    a = s.solve_batch(b.x0, xref_const=xc)
    assert np.array_equal(a["x"], full["x"]) and np.array_equal(a["u"], full["u"])
    d = s.solve_batch(b.x0, b.Xref, None, compact_out=True)
    assert np.array_equal(d["u0"], full["u"][:, 0, :])
    # device entry
    dev = torch.device("cuda", 0)
    x0, xcd = torch.from_numpy(b.x0).to(dev), torch.from_numpy(xc).to(dev)
    u0 = torch.empty((B, p.nu), device=dev); it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    s.solve_batch_device(B, x0.data_ptr(), None, None, None, None, it.data_ptr(), st.data_ptr(), xref_const=xcd.data_ptr(), u0=u0.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(it.cpu().numpy(), full["iter"]) and np.array_equal(u0.cpu().numpy(), full["u"][:, 0, :])
    # argument checks
    with pytest.raises(capi.TinympcCudaError):
        s.solve_batch(b.x0, b.Xref, None, xref_const=xc)
    s.close()


@pytest.mark.parametrize("family,mixed,fixer_sms", [("quadrotor", 0.0, -2), ("quadrotor", 0.002, -2), ("quadrotor", 0.002, 0), ("cartpole", 0.003, -2),
                                                    ("quadrotor_noref", 0.0, -2), ("quadrotor_compact", 0.002, -2)])
def test_claim_order_is_scheduling_only(family, mixed, fixer_sms, capi, oracle_mod, problems):
    """Option "order" (default on): a device-resident batch of a box family is bucketed by its expected difficulty
    |Kinf (x0 - xref_0)| / u_bound on the device (counting sort, tmpc_capi.cu order_*_kernel) and the thread-per-problem kernel
    claims the hardest problems first.  Pure scheduling: the results are, bit for bit, those of the index-order run -- in plain fp32,
    in both forms of the exact-count mode, without references, and with compact I/O."""
    import torch
    p = problems.cartpole() if family == "cartpole" else problems.quadrotor()
    B = 300007
    b = problems.make_batch(p, B, 1.0, seed=61)
    if family == "quadrotor_noref":
        b = problems.Batch(b.x0, None, None)
    s = capi.CudaSolver()
    s.set_option("mixed", mixed)
    s.set_option("fixer_sms", fixer_sms)
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    dev = torch.device("cuda", 0)
    tdev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ptr = lambda t: None if t is None else t.data_ptr()
    x0, Xref, Uref = tdev(b.x0), tdev(b.Xref), tdev(b.Uref)
    compact = family == "quadrotor_compact"
    xc = tdev(b.Xref[:, 0, :]) if compact else None
    res, launches = {}, {}
    for order in (0, 1, 1, 0):
        s.set_option("order", order)
        x = torch.full((B, p.N, p.nx), -3.0, device=dev); u = torch.full((B, p.N - 1, p.nu), -3.0, device=dev)
        u0 = torch.full((B, p.nu), -3.0, device=dev)
        it = torch.full((B,), -3, dtype=torch.int32, device=dev); st = torch.full((B,), -3, dtype=torch.int32, device=dev)
        n0 = s.launch_count
        if compact:
            s.solve_batch_device(B, ptr(x0), None, None, None, None, ptr(it), ptr(st), xref_const=ptr(xc), u0=ptr(u0), stream=torch.cuda.current_stream().cuda_stream)
        else:
            s.solve_batch_device(B, ptr(x0), ptr(Xref), ptr(Uref), ptr(x), ptr(u), ptr(it), ptr(st), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        launches[order] = s.launch_count - n0
        assert s.last_kernel.startswith("tpp3_"), s.last_kernel
        r = dict(iter=it.cpu().numpy(), status=st.cpu().numpy(), **(dict(u0=u0.cpu().numpy()) if compact else dict(x=x.cpu().numpy(), u=u.cpu().numpy())))
        if not res:
            res = r
            assert (r["iter"] >= 1).all() and np.isin(r["status"], (1, 11)).all()
        for k in res:
            assert np.array_equal(res[k], r[k]), f"{family}: order={order} changes {k}"
    assert launches[1] == launches[0] + 3, launches      # count, scan, scatter
    # the compact streamed host pipeline orders too: the first quarter of the shard in index order while the rest is still
    # arriving, the rest hardest-first through a list built meanwhile on the SMs the persistent launch leaves free
    if fixer_sms < 0 and (b.Xref is None or (b.Xref == b.Xref[:, :1]).all()):
        xcn = None if b.Xref is None else np.ascontiguousarray(b.Xref[:, 0, :])
        for order in (1, 0, 1):
            s.set_option("order", order)
            n0 = s.launch_count
            c = s.solve_batch(b.x0, xref_const=xcn, compact_out=True)
            launches[order] = s.launch_count - n0
            assert s.last_timing()["chunks"] >= 3, s.last_timing()
            assert np.array_equal(c["iter"], res["iter"]) and np.array_equal(c["status"], res["status"]), f"{family}: streamed, order={order}"
            assert np.array_equal(c["u0"], res["u0"] if compact else res["u"][:, 0, :]), f"{family}: streamed, order={order}"
        if p.nx >= 8:    # (the streamed form orders only shards of problems with >= 8 states)
            assert launches[1] > launches[0] and (launches[1] - launches[0]) % 3 == 0, launches     # count, scan, scatter per ordered chunk
    s.close()


@pytest.mark.timeout(180)
def test_ordered_streamed_shards_of_two_contexts_on_one_device(capi, oracle_mod, problems):
    """Two device contexts on the SAME GPU, each with a compact shard large enough for the ordered streamed form: its ordering kernels
    run on the SM the persistent launch leaves free, which a second persistent launch would take (both would then wait for ordering
    kernels that can never be scheduled) -- host batch calls are therefore serialised per device.  Same bits as one context."""
    p = problems.quadrotor()
    B = 600014
    b = problems.make_batch(p, B, 1.0, seed=71)
    xc = np.ascontiguousarray(b.Xref[:, 0, :])
    fam = cases.family_from_spec(p, oracle_mod.get_cache(p, "port"))
    res = []
    for devices in ([0], [0, 0]):
        s = capi.CudaSolver(devices=devices)
        s.set_option("mixed", problems.exact_band(p))
        s.set_family(fam)
        res.append(s.solve_batch(b.x0, xref_const=xc, compact_out=True))
        assert s.last_timing()["chunks"] >= 3
        s.close()
    for k in ("iter", "status", "u0"):
        assert np.array_equal(res[0][k], res[1][k]), k


@pytest.mark.parametrize("family,precision,B", [("cartpole", 32, 3000), ("cartpole", 64, 3000), ("quadrotor", 64, 3000), ("quadrotor", 32, 60000),
                                                ("rocket", 32, 3000), ("quadrotor_adaptive", 64, 3000), ("cartpole", 32, 140000), ("quadrotor", 64, 70000),
                                                ("quadrotor_adaptive", 64, 66000), ("rocket", 32, 70000)])
def test_compact_reference_read_in_place(family, precision, B, capi, oracle_mod, problems):
    """SolveParams::xref_const: the incremental fp32 kernel and the lane-group fp64 kernel read ONE reference state per problem in
    place of every column of the horizon; every other kernel (here: the rocket's mixed-precision kernel) gets the reference
    replicated on the device first.  Either way the result is, bit for bit, that of the full-trajectory call."""
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket,
             quadrotor_adaptive=lambda: problems.quadrotor(adaptive=True))[family]()
    b = problems.make_batch(p, B, 1.0, seed=41)
    xc = (0.1 * np.random.default_rng(9).standard_normal((B, p.nx))).astype(np.float32)
    Xref = np.ascontiguousarray(np.repeat(xc[:, None, :], p.N, axis=1))
    s = capi.CudaSolver()
    s.set_option("precision", precision)
    if precision == 32:
        s.set_option("mixed", problems.exact_band(p))
    s.set_family(cases.family_from_spec(p, oracle_mod.get_cache(p, "port")))
    s.set_option("chunks", 1)
    full = s.solve_batch(b.x0, Xref, None)
    s.set_option("chunks", 0)
    for in_kernel in (1, 0):
        s.set_option("compact_in_kernel", in_kernel)
        c = s.solve_batch(b.x0, xref_const=xc)
        for k in ("iter", "status", "x", "u"):
            assert np.array_equal(full[k], c[k]), f"{family} fp{precision} in_kernel={in_kernel}: compact reference changes {k} ({s.last_kernel})"
        # ... and SolveParams::u0: the first control as the only solution output (no trajectory scratch, no gather)
        c = s.solve_batch(b.x0, xref_const=xc, compact_out=True)
        assert np.array_equal(full["iter"], c["iter"]) and np.array_equal(full["status"], c["status"])
        assert np.array_equal(full["u"][:, 0, :], c["u0"]), f"{family} fp{precision} in_kernel={in_kernel}: compact output differs ({s.last_kernel})"
    s.close()


@pytest.mark.parametrize("which", ["xref_only", "uref_only", "none"])
def test_partial_references_on_the_hybrid_layout(which, capi, oracle_mod, problems):
    """Xref without Uref, Uref without Xref, neither: the reference-parking code of the refill pass has a branch per case
    (tensor-memory / shared-memory / register columns of the hybrid layout).  60 000 problems, so the hybrid instance runs;
    a 4 000-problem prefix is compared with the oracle, the rest through the plain-layout instance (bit for bit)."""
    p = problems.quadrotor()
    B = 60000
    b = problems.make_batch(p, B, 1.0, seed=57)
    Uref = (0.05 * np.random.default_rng(3).standard_normal((B, p.N - 1, p.nu))).astype(np.float32)
    bb = problems.Batch(b.x0, b.Xref if which == "xref_only" else None, Uref if which == "uref_only" else None)
    r = solve_gpu(capi, oracle_mod, p, bb, 32)
    rp = solve_gpu(capi, oracle_mod, p, bb, 32, variant=9)          # plain layout, same arithmetic
    assert r["kernel"] != rp["kernel"]
    for k in ("iter", "status", "x", "u"):
        assert np.array_equal(r[k], rp[k]), f"{which}: hybrid and plain layouts disagree in {k}"
    n = 4000
    g = oracle_mod.solve_batch(p, bb.slice(0, n), "ref" if oracle_mod.available("ref") else "port")
    rs = {k: r[k][:n] for k in ("iter", "status", "x", "u")}
    rs["kernel"] = r["kernel"]
    flips, dx, du = compare(rs, g, 32, f"quadrotor {which}", max_flip_frac=FLIP_BOUND["quadrotor"])
    print(f"\n[parity] quadrotor {which}: hybrid == plain bit for bit; {flips}/{n} count flips, max|dx|={dx:.2e} kernel={r['kernel']}")


def test_hybrid_layout_with_per_problem_and_time_varying_bounds(capi, oracle_mod, problems):
    """The two other hybrid-layout instances of the quadrotor shape at a batch size that selects them (> one wave of the plain
    layout): per-problem bounds (tiled golden case) and shared bounds that vary over the horizon / do not contain 0 (no fast-box
    shortcut).  Hybrid == plain layout bit for bit, and both match the reference."""
    # (a) per-problem bounds
    p, b0, g = cases.load("batch_quadrotor_perproblem_bounds")
    reps = 60000 // b0.size + 1
    tile = lambda a: None if a is None else np.ascontiguousarray(np.concatenate([a] * reps, axis=0))
    bb = problems.Batch(tile(b0.x0), tile(b0.Xref), tile(b0.Uref), tile(b0.x_min), tile(b0.x_max), tile(b0.u_min), tile(b0.u_max))
    r = solve_gpu(capi, oracle_mod, p, bb, 32)
    rp = solve_gpu(capi, oracle_mod, p, bb, 32, variant=9)
    assert "ppb" in r["kernel"] and r["kernel"] != rp["kernel"]
    for k in ("iter", "status", "x", "u"):
        assert np.array_equal(r[k], rp[k]), f"per-problem bounds: hybrid and plain layouts disagree in {k}"
    n = b0.size
    same = r["iter"][:n] == g["iter"]
    assert same.mean() >= 0.95 and np.abs(r["x"][:n][same] - g["x"][same]).max() <= X_TOL_F32 * max(1.0, float(np.abs(g["x"]).max()))
    assert np.array_equal(r["iter"][:n], r["iter"][7 * n:8 * n]) and np.array_equal(r["x"][:n], r["x"][7 * n:8 * n])
    # (b) shared bounds varying over the horizon, lower input bound above 0 on the last steps
    q = problems.quadrotor()
    u_min, u_max = q.u_min.copy(), q.u_max.copy()
    u_max[:] = np.linspace(0.5, 0.3, q.N - 1)[:, None]
    u_min[-3:] = 0.01
    q2 = q.with_(u_min=u_min, u_max=u_max)
    b = problems.make_batch(q2, 60000, 1.0, seed=91)
    r = solve_gpu(capi, oracle_mod, q2, b, 32)
    rp = solve_gpu(capi, oracle_mod, q2, b, 32, variant=9)
    assert "_fb" not in r["kernel"] and r["kernel"] != rp["kernel"], r["kernel"]
    for k in ("iter", "status", "x", "u"):
        assert np.array_equal(r[k], rp[k]), f"time-varying bounds: hybrid and plain layouts disagree in {k}"
    n = 3000
    gg = oracle_mod.solve_batch(q2, b.slice(0, n), "ref" if oracle_mod.available("ref") else "port")
    rs = {k: r[k][:n] for k in ("iter", "status", "x", "u")}
    rs["kernel"] = r["kernel"]
    flips, dx, du = compare(rs, gg, 32, "quadrotor time-varying bounds", max_flip_frac=0.01)
    print(f"\n[parity] quadrotor hybrid ppb + time-varying bounds: == plain bit for bit; {flips}/{n} count flips, max|dx|={dx:.2e} kernel={r['kernel']}")


# ---- lane-group-per-problem fp64 kernel (tmpc_gpp.cuh) ---------------------------------------------------------------------------
def _gpp_family(problems, name):
    """box families of the three compiled shapes; the rocket without its cones and rows (a box family with the gravity term f)"""
    if name == "rocket_box":
        p = problems.rocket(linear=False)
        p.en_state_soc = p.en_input_soc = 0
        p.Acx = p.qcx = p.Acu = p.qcu = np.zeros(0, np.int32)
        p.cx = p.cu = np.zeros(0)
        p.name = "rocket_box"
        return p
    return dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor)[name]()


@pytest.mark.parametrize("tweak", ["plain", "norefs", "moving_bounds", "no_bounds", "check3", "x0_outside", "few_iters"])
@pytest.mark.parametrize("family", ["quadrotor", "cartpole", "rocket_box"])
def test_lane_group_kernel_matches_the_reference(family, tweak, capi, oracle_mod, problems):
    """fp64 batches of the compiled box shapes run the lane-group kernel: the reference's iteration counts and statuses on every
    problem, x / u to float32 output rounding -- with and without references, bounds that move along the horizon (not the
    fast-box case), bounds switched off, check_termination = 3, x0 outside its box (the column-0 slack never converges), a
    max_iter that cuts every solve short -- and the thread-per-problem fp64 kernel (variant 6) returns the same counts."""
    p = _gpp_family(problems, family)
    B = 3000
    b = problems.make_batch(p, B, 1.0, seed=77)
    if tweak == "norefs":
        b.Xref = None; b.Uref = None
    elif tweak == "moving_bounds":
        ramp = np.linspace(1.0, 0.6, p.N)[:, None]
        p.x_min, p.x_max = p.x_min * ramp, p.x_max * ramp
        p.u_min, p.u_max = p.u_min * ramp[:-1], p.u_max * ramp[:-1]
    elif tweak == "no_bounds":
        p.en_state_bound = p.en_input_bound = 0
    elif tweak == "check3":
        p.check_termination = 3
    elif tweak == "x0_outside":
        b.x0 = (b.x0 + np.float32(1.2) * np.asarray(p.x_max[0], np.float32) * (np.arange(B)[:, None] % 3 == 0)).astype(np.float32)
    elif tweak == "few_iters":
        p.max_iter = 7
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    r = solve_gpu(capi, oracle_mod, p, b, 64)
    assert r["kernel"].startswith("gpp_f64"), r["kernel"]
    compare(r, g, 64, f"{family}/{tweak}")
    r6 = solve_gpu(capi, oracle_mod, p, b, 64, variant=6)
    assert r6["kernel"].startswith("tpp2_f64"), r6["kernel"]
    assert np.array_equal(r["iter"], r6["iter"]) and np.array_equal(r["status"], r6["status"])
    assert np.abs(r["x"] - r6["x"]).max() <= 1e-6 * max(1.0, float(np.abs(g["x"]).max()))


def test_lane_group_kernel_adaptive_rho_matches_the_reference(capi, oracle_mod, problems):
    """adaptive rho on the lane-group kernel: counts, statuses, solutions AND the final rho of every problem (rho_benchmark.cpp:175-212)"""
    p = problems.quadrotor(adaptive=True)
    B = 4000
    b = problems.make_batch(p, B, 1.0, seed=78)
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    r = solve_gpu(capi, oracle_mod, p, b, 64)
    assert r["kernel"].startswith("gpp_f64") and "_adp_" in r["kernel"], r["kernel"]
    compare(r, g, 64, "quadrotor_adaptive/gpp")
    assert np.abs(r["rho"] - g["rho"]).max() < 1e-5 * max(1.0, float(np.abs(g["rho"]).max()))
    assert len(np.unique(np.round(g["rho"], 3))) > 10      # the batch really adapts


@pytest.mark.parametrize("B", [1, 2, 17, 4097])
def test_lane_group_kernel_ragged_batches(B, capi, oracle_mod, problems):
    """batches that do not fill a warp / a CTA / a wave: idle groups must neither write nor stall the others"""
    p = problems.quadrotor()
    b = problems.make_batch(p, B, 1.0, seed=79)
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    r = solve_gpu(capi, oracle_mod, p, b, 64)
    assert r["kernel"].startswith("gpp_f64"), r["kernel"]
    compare(r, g, 64, f"ragged {B}")


@pytest.mark.parametrize("linear", [True, False])
def test_lane_group_kernel_cone_family_matches_the_reference(linear, capi, oracle_mod, problems):
    """rocket family (box + one second-order cone per side, with and without the two linear rows) in fp64 on the lane-group kernel
    (rolled loops, per-slot state in shared memory, cone / half-space projections by the owning lanes): the reference's counts on
    every problem; and as the second pass of the exact-count mode behind the mixed-precision kernel."""
    p = problems.rocket(linear=linear)
    B = 6000
    b = problems.make_batch(p, B, 1.0, seed=81)
    g = oracle_mod.solve_batch(p, b, "ref" if oracle_mod.available("ref") else "port")
    r = solve_gpu(capi, oracle_mod, p, b, 64)
    assert r["kernel"].startswith("gpp_f64") and "_con_" in r["kernel"], r["kernel"]
    compare(r, g, 64, f"rocket linear={linear} gpp")
    rx = solve_gpu(capi, oracle_mod, p, b, 32, mixed=problems.exact_band(p), fixer_sms=-1)
    assert "tpp4" in rx["kernel"] and "+gpp_f64" in rx["kernel"] and 0 < rx["marked"] < B, (rx["kernel"], rx["marked"])
    if linear:      # BASELINE config 4: the north-star bar
        compare(rx, g, 32, f"rocket linear={linear} exact-count")
    else:           # not a BASELINE config: identical counts; the mixed-precision kernel's thrusts (~100) within 2e-6 relative
        assert np.array_equal(rx["iter"], g["iter"]) and np.array_equal(rx["status"], g["status"])
        assert max(np.abs(rx["x"] - g["x"]).max(), np.abs(rx["u"] - g["u"]).max()) <= 2e-4
