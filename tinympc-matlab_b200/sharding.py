"""Problem-index sharding of a batch over ranks / devices.

Problems are independent, so the path shards with NO collective on the data path: rank r of W owns the
contiguous range shard_range(B, r, W).  The only communication of a multi-GPU run is the reduction of the
per-rank timings and counters for the report (max of the device times, sum of the work), which is what
reduce_report() does over whatever torch.distributed backend is initialised (nccl on the GPU box, gloo in
the CPU tests).  The C library splits host batches the same way (tmpc_capi.cu::tinympc_cuda_solve_batch).
"""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int, align: int = 4):
    """[lo, hi) of rank `rank`: equal contiguous pieces rounded up to `align` problems (keeps every shard's
    float arrays 16-byte aligned for all shapes), the remainder to the last rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    per = ((batch + world - 1) // world + align - 1) // align * align
    lo = min(batch, rank * per)
    hi = batch if rank == world - 1 else min(batch, lo + per)
    return lo, hi


def reduce_report(ms: float, sums, dist=None, device=None):
    """max over ranks of the device time `ms`, sum over ranks of the counters `sums` (a list of floats)."""
    import torch
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    s = torch.tensor([float(v) for v in sums], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(v) for v in s.tolist()]
