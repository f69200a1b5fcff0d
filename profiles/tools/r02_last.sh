#!/bin/bash
# Last evidence run of round 2 on one B200 (gpurun -- bash profiles/tools/r02_last.sh): the whole GPU suite, the default bench line,
# the configs whose host path changed (cartpole, adaptive rho), smoke(), and the launch list of the bench command.
set -u
O=gpurun_out/r02_last; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 400 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_quadrotor_n1.json 2> $O/bench_quadrotor_n1.err
timeout 200 python bench.py --config cartpole --cpu-seconds 3 > $O/bench_cartpole_n1.json 2> $O/bench_cartpole_n1.err
timeout 200 python bench.py --config quadrotor_adaptive --no-cpu-baseline --steps 5 > $O/bench_quadrotor_adaptive_n1.json 2> $O/bench_quadrotor_adaptive_n1.err
python - $O/bench_*_n1.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); p=d.get("parity") or {}; e=d.get("e2e") or {}; rc=d.get("resident_compact") or {}; o=d.get("e2e_other") or {}
        print(f.split("/")[-1], "value", round(d["value"]/1e6,3), "M/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "fp64_ms", d["roofline"].get("fp64_pass_ms"),
              "e2e", round(e.get("value",0)/1e6,2), e.get("pipeline"), "other", round(o.get("value",0)/1e6,2), "resident_compact", round(rc.get("value",0)/1e6,2), "parity", p.get("pass"), p.get("count_mismatch"), p.get("max_abs_du"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
    except Exception as ex: print(f, "failed", ex)
PY
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -4 $O/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_quadrotor.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 > $O/launches_bench.log 2>&1; tail -2 $O/launches_bench.log | cut -c1-300
grep -c "tpp3_kernel\|gpp_kernel\|order_" $O/launches_bench_quadrotor.csv
