#!/bin/bash
# Multi-GPU evidence (gpurun --gpus N -- bash profiles/tools/r02_multigpu.sh N): the headline bench in weak and strong scaling, the
# BASELINE config-5 batch sweep (adaptive rho, parity-exact fp64 mode) and the plain-fp32 sweep of the quadrotor.
set -u
N=${1:-2}
O=gpurun_out/r02_mg; mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
run() { if [ "$N" = 1 ]; then timeout 600 python "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; fi; }
run bench.py --gpus $N --no-cpu-baseline > $O/bench_quadrotor_weak_n$N.json 2> $O/err_weak_n$N.log
run bench.py --gpus $N --no-cpu-baseline --scaling strong > $O/bench_quadrotor_strong_n$N.json 2> $O/err_strong_n$N.log
run bench.py --gpus $N --no-cpu-baseline --config quadrotor_adaptive --steps 5 > $O/bench_adaptive_weak_n$N.json 2> $O/err_adp_n$N.log
run profiles/tools/batch_sweep.py --config quadrotor_adaptive --precision 64 --max-log2 24 > $O/batch_sweep_adaptive_fp64_n$N.jsonl 2> $O/err_sweep_n$N.log
run profiles/tools/batch_sweep.py --config quadrotor --precision 32 --max-log2 24 > $O/batch_sweep_quadrotor_fp32_n$N.jsonl 2>> $O/err_sweep_n$N.log
for f in $O/bench_*_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); p=d.get("parity") or {}; e=d.get("e2e") or {}; o=d.get("e2e_other") or {}
    print(sys.argv[1].split("/")[-1], d["n_gpus"], "GPUs", d["scaling"], round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],3), "ms  e2e", round(e.get("value",0)/1e6,2), "/", round(o.get("value",0)/1e6,2), "parity", p.get("pass"), d["config"]["batch_per_gpu"], d["clocks"])
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
done
cut -c1-230 $O/batch_sweep_adaptive_fp64_n$N.jsonl; cut -c1-230 $O/batch_sweep_quadrotor_fp32_n$N.jsonl; tail -3 $O/err_*_n$N.log | cut -c1-300
