#!/usr/bin/env python3
"""BASELINE config 5: batch-size sweep B = 2^10 .. 2^24 of a config (default: the adaptive-rho quadrotor) on the
GPUs of this process group.  One process per GPU (torchrun for N > 1); every rank owns B problems (weak scaling, contiguous
problem-index shards, no collective on the path); device-resident inputs, CUDA events on the launching stream, max over ranks.
One JSON line per batch size.  Usage: python profiles/tools/batch_sweep.py [--config quadrotor_adaptive] [--max-log2 24]"""
import argparse, importlib, json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT))
tm = importlib.import_module("tinympc-matlab_b200"); P = importlib.import_module("tinympc-matlab_b200.problems")
S = importlib.import_module("tinympc-matlab_b200.sharding")

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="quadrotor_adaptive")
ap.add_argument("--min-log2", type=int, default=10)
ap.add_argument("--max-log2", type=int, default=24)
ap.add_argument("--step-log2", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--precision", type=int, default=32)
a = ap.parse_args()
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
    os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)     # keep NCCL's version banner out of the JSON lines
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
spec = dict(quadrotor=P.quadrotor, cartpole=P.cartpole, rocket=P.rocket, quadrotor_adaptive=lambda: P.quadrotor(adaptive=True))[a.config]()
n, m, N = spec.nx, spec.nu, spec.N
s = tm.TinyMPC(); s.setup_from_spec(spec, devices=[local]); s.cuda.set_option("precision", a.precision)
# the largest batch is generated once per rank (2^20 distinct problems tiled up to B: the kernel sees every problem as new)
base = P.make_batch(spec, 1 << min(a.max_log2, 20), a.scale, seed=1234 + 5 + 1000 * rank)
for lg in range(a.min_log2, a.max_log2 + 1, a.step_log2):
    B = 1 << lg
    rep = max(1, B // base.size)
    tile = lambda v: None if v is None else torch.from_numpy(v[:min(B, base.size)]).to(dev).repeat((rep,) + (1,) * (v.ndim - 1)).contiguous()
    x0, Xr, Ur = tile(base.x0), tile(base.Xref), tile(base.Uref)
    x = torch.empty((B, N, n), device=dev); u = torch.empty((B, N - 1, m), device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream()
    step = lambda: s.cuda.solve_batch_device(B, ptr(x0), ptr(Xr), ptr(Ur), ptr(x), ptr(u), ptr(it), ptr(st), stream=stream.cuda_stream)
    steps = max(3, min(200, (1 << 22) // B))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    iters = int(it.sum().item())
    ms_all, (iters_all,) = S.reduce_report(ms, [iters], dist, dev)
    if rank == 0:
        print(json.dumps(dict(config=a.config, n_gpus=world, batch_per_gpu=B, log2_batch=lg, steps=steps, ms_per_step=ms_all / steps,
                              solves_per_sec=world * B * steps / (ms_all * 1e-3), ns_per_admm_iter=ms_all * 1e6 / (iters_all * steps),
                              mean_iters=iters_all / (world * B), kernel=s.cuda.last_kernel, dtype=f"f{a.precision}",
                              l2_note="inputs+outputs %.1f MB per step (L2 = 126 MB)" % ((4 * (n + 2 * n * N + 2 * m * (N - 1)) + 8) * B / 1e6))), flush=True)
    del x0, Xr, Ur, x, u, it, st
    torch.cuda.empty_cache()
if dist is not None:
    dist.destroy_process_group()
