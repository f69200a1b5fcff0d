"""GPU test of the batched cache precompute + rho-sensitivities (tinympc_cuda_precompute_batch) against the
reference's own tiny_precompute_and_set_cache (oracle/_ref, tiny_api.cpp:244-318) on per-problem random (A, B, Q, R, rho),
and of the sensitivities against a forward difference of scipy's DARE solution (what TinyMPC.m:223-241 does with idare)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return importlib.import_module("tinympc-matlab_b200.capi")


def perturbed(p, n, seed):
    """n variations of a problem family: dynamics perturbed by 1 %, costs scaled by 0.5 .. 2, rho in [0.5, 2] x the family's"""
    rng = np.random.default_rng(seed)
    A = p.A[None] * (1.0 + 0.01 * rng.standard_normal((n,) + p.A.shape))
    B = np.asarray(p.B).reshape(p.nx, p.nu)[None] * (1.0 + 0.01 * rng.standard_normal((n, p.nx, p.nu)))
    Q = p.Qdiag[None] * rng.uniform(0.5, 2.0, (n, p.nx))
    R = p.Rdiag[None] * rng.uniform(0.5, 2.0, (n, p.nu))
    rho = p.rho * rng.uniform(0.5, 2.0, n)
    f = np.broadcast_to(np.asarray(p.f, np.float64).ravel(), (n, p.nx)).copy()
    return A, B, Q, R, rho, f


@pytest.mark.parametrize("family", ["cartpole", "quadrotor", "rocket"])
def test_precompute_batch_matches_the_reference(family, capi, oracle_mod, problems):
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor, rocket=problems.rocket)[family]()
    n = 24
    A, B, Q, R, rho, f = perturbed(p, n, seed=5)
    # tiny_setup hands Q + rho to the precompute, which adds rho again: mirror that here
    o = capi.precompute_batch(A, B, Q + rho[:, None], R + rho[:, None], rho, f)
    impl = "ref" if oracle_mod.available("ref") else "port"
    for b in range(n):
        g = oracle_mod.get_cache(p.with_(A=A[b], B=B[b], Qdiag=Q[b], Rdiag=R[b], rho=float(rho[b])), impl)
        for k in ("Kinf", "Pinf", "Quu_inv", "AmBKt", "APf", "BPf"):
            ref = np.asarray(g[k]).reshape(o[k][b].shape)
            err = np.abs(o[k][b] - ref).max() / max(1.0, np.abs(ref).max())
            assert err < 1e-9, f"{family} problem {b}: {k} differs by {err:.2e} (relative)"
    assert (o["iters"] > 1).all() and (o["iters"] <= 1000).all()


@pytest.mark.parametrize("family", ["cartpole", "quadrotor"])
def test_sensitivities_match_a_forward_difference_of_the_dare_solution(family, capi, problems):
    from scipy.linalg import solve_discrete_are
    p = dict(cartpole=problems.cartpole, quadrotor=problems.quadrotor)[family]()
    n = 6
    A, B, Q, R, rho, f = perturbed(p, n, seed=9)
    o = capi.precompute_batch(A, B, Q, R, rho, f, sensitivities=True)

    def lqr(b, r):
        Q1, R1 = np.diag(Q[b] + r), np.diag(R[b] + r)
        P = solve_discrete_are(A[b], B[b], Q1, R1)
        K = np.linalg.solve(R1 + B[b].T @ P @ B[b], B[b].T @ P @ A[b])
        return K, P, np.linalg.inv(R1 + B[b].T @ P @ B[b]), (A[b] - B[b] @ K).T

    h = 1e-6
    for b in range(n):
        lo, hi = lqr(b, rho[b]), lqr(b, rho[b] + h)
        for k, i in (("dKinf", 0), ("dPinf", 1), ("dC1", 2), ("dC2", 3)):
            ref = (hi[i] - lo[i]) / h
            err = np.abs(o[k][b] - ref).max() / max(1e-12, np.abs(ref).max())
            assert err < 5e-3, f"{family} problem {b}: {k} differs by {err:.2e} (relative to its largest entry)"
        # and the cache itself against the converged DARE solution (the recursion stops at 1e-5 in Kinf)
        assert np.abs(o["Kinf"][b] - lo[0]).max() < 1e-3 * max(1.0, np.abs(lo[0]).max())
