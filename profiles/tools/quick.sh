#!/bin/bash
# Quick GPU check of a kernel change: GPU tests + the kernel-only bench lines of the BASELINE configs.
# Usage (GPU box, repo root): bash profiles/tools/quick.sh <tag> [configs...]
set -u
TAG=$1; shift
CFGS=${@:-quadrotor cartpole rocket quadrotor_adaptive}
O=gpurun_out/$TAG; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q -s) > $O/pytest_gpu.log 2>&1
grep -E "parity\]|passed|failed|rror" $O/pytest_gpu.log | cut -c1-220 | tail -40
for c in $CFGS; do
  timeout 300 python bench.py --config $c --no-cpu-baseline --no-e2e > $O/bench_$c.json 2> $O/bench_$c.err
  python - "$O/bench_$c.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{d['config']['workload'].split(',')[0]:28s} {d['value']/1e6:8.2f} M solves/s  {d['ns_per_admm_iter']:.4f} ns/iter  roofline {d['roofline']['frac']:.3f}  {d['roofline']['kernel']}")
except Exception as e:
    print("bench failed:", sys.argv[1], e)
PY
done
