"""CPU check of the scheduling heuristic behind option "order" (tmpc_capi.cu order_count_kernel): how far the
unconstrained feedback -Kinf (x0 - xref_0) leaves the input bounds predicts how many ADMM iterations the reference needs.
The predictor only decides the ORDER in which the GPU lanes claim the problems -- results never depend on it -- so what is pinned
here is that the order is worth having: rank correlation with the reference's iteration counts, the "sure-easy" threshold, and the
makespan of a lane-level list schedule (one problem per lane at a time, the next one claimed when it finishes)."""
import heapq
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def _key(spec, b, Kinf):
    d = b.x0.astype(np.float64) - (b.Xref[:, 0, :] if b.Xref is not None else 0.0)
    ub = np.minimum(-np.asarray(spec.u_min[0], np.float64), np.asarray(spec.u_max[0], np.float64))
    return (np.abs(d @ Kinf.T) / ub).max(axis=1)


def _makespan(iters, order, lanes):
    h = [0] * lanes
    heapq.heapify(h)
    for j in order:
        heapq.heappush(h, heapq.heappop(h) + int(iters[j]))
    return max(h)


def test_difficulty_key_predicts_the_reference_iteration_counts():
    import oracle as O
    P = importlib.import_module("tinympc-matlab_b200.problems")
    spec = P.quadrotor()
    B = 20000
    b = P.make_batch(spec, B, 1.0, seed=1237)
    g = O.solve_batch(spec, b, "ref" if O.available("ref") else "port")
    it = g["iter"].astype(np.int64)
    Kinf = np.asarray(O.get_cache(spec, "port")["Kinf"], np.float64).reshape(spec.nu, spec.nx)
    key = _key(spec, b, Kinf)
    rank = lambda v: np.argsort(np.argsort(v, kind="stable"), kind="stable")
    rho = float(np.corrcoef(rank(key), rank(it))[0, 1])
    assert rho > 0.85, rho                                   # measured 0.90
    easy = key < 0.6
    assert 0.15 < easy.mean() < 0.45 and it[easy].max() <= 20, (easy.mean(), it[easy].max())      # a quarter of the batch, <= 17 iterations
    assert (it[key >= 2.0] >= 88).mean() > 0.9               # twice the bound and more: (almost) always up to max_iter
    # lane-level list schedule with the batch's ratio of problems to lanes (2^20 problems on 148 x 384 lanes = 18.45 per lane)
    lanes = int(B / 18.45)
    bound = it.sum() / lanes
    as_generated = _makespan(it, range(B), lanes) / bound
    hardest_first = _makespan(it, np.argsort(-key, kind="stable"), lanes) / bound
    bucketed = _makespan(it, np.argsort(-np.minimum(255, np.floor(key * 64)), kind="stable"), lanes) / bound   # the kernel's 256 buckets
    assert as_generated > 1.06 and hardest_first < 1.02 and bucketed < 1.02, (as_generated, hardest_first, bucketed)
