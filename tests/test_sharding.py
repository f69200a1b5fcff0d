"""Multi-rank host logic on CPU (gloo, world_size 2): problem-index shards cover the batch exactly once and the
report reduction is max-of-times / sum-of-work.  The data path itself has no collective."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

S = importlib.import_module("tinympc-matlab_b200.sharding")


@pytest.mark.parametrize("batch,world", [(0, 1), (1, 2), (7, 2), (1 << 20, 8), (1000003, 8), (12, 8), (5, 4)])
def test_shards_partition_the_batch(batch, world):
    seen = np.zeros(batch, dtype=np.int32)
    prev_hi = 0
    for r in range(world):
        lo, hi = S.shard_range(batch, r, world)
        assert lo == prev_hi and lo <= hi <= batch
        assert lo % 4 == 0 or lo == batch
        seen[lo:hi] += 1
        prev_hi = hi
    assert prev_hi == batch and (seen == 1).all()


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        S.shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("tinympc-matlab_b200.problems")
    spec = P.quadrotor()
    B = 1001
    lo, hi = S.shard_range(B, rank, world)
    full = P.make_batch(spec, B, 1.0, seed=5)
    mine = full.slice(lo, hi)
    # stand-in for the per-rank solve: a deterministic per-problem number, so that the gathered result can be checked
    local = mine.x0.astype(np.float64).sum(axis=1)
    ms, sums = S.reduce_report(10.0 + rank, [hi - lo, float(local.sum())], dist)
    q.put((rank, lo, hi, ms, sums, float(full.x0.astype(np.float64).sum())))
    dist.destroy_process_group()


def test_two_rank_gloo_report():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ms0, sums0, tot0), (r1, lo1, hi1, ms1, sums1, _) = out
    assert (lo0, hi1) == (0, 1001) and hi0 == lo1
    assert ms0 == ms1 == 11.0                        # max over ranks
    assert sums0[0] == sums1[0] == 1001              # every problem counted once
    assert abs(sums0[1] - tot0) < 1e-6
