#!/bin/bash
# Last evidence run of round 2 on one B200 (gpurun -- bash profiles/tools/r02_last.sh): the whole GPU suite, smoke(), the default
# bench line -- all on the final tree.  (The launch list, the cartpole / adaptive lines and the A/B files in profiles/r02 come from
# the runs of profiles/tools/r02_order.sh and the earlier form of this script, one or two commits before.)
set -u
O=gpurun_out/r02_last; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 300 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
(timeout 150 python -c "import __graft_entry__ as g; g.smoke()") > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -4 $O/smoke.log | cut -c1-400
timeout 200 python bench.py > $O/bench_quadrotor_n1.json 2> $O/bench_quadrotor_n1.err
python - $O/bench_quadrotor_n1.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); p=d.get("parity") or {}; e=d.get("e2e") or {}; rc=d.get("resident_compact") or {}; o=d.get("e2e_other") or {}
        print(f.split("/")[-1], "value", round(d["value"]/1e6,3), "M/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "fp64_ms", d["roofline"].get("fp64_pass_ms"),
              "e2e", round(e.get("value",0)/1e6,2), e.get("pipeline"), "other", round(o.get("value",0)/1e6,2), "resident_compact", round(rc.get("value",0)/1e6,2), "parity", p.get("pass"), p.get("count_mismatch"), p.get("max_abs_du"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
    except Exception as ex: print(f, "failed", ex)
PY
