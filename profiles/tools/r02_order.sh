#!/bin/bash
# Round-2 follow-up run on one B200 (gpurun -- bash profiles/tools/r02_order.sh): claim order (option "order") A/B.
set -u
O=gpurun_out/r02_order; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "claim_order or compact") > $O/pytest_new.log 2>&1; echo "new tests rc=$?" >> $O/pytest_new.log; tail -5 $O/pytest_new.log
show() { python - "$@" <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); p=d.get("parity") or {}; e=d.get("e2e") or {}; rc=d.get("resident_compact") or {}
        print(f.split("/")[-1], "value", round(d["value"]/1e6,3), "M/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "fp64_ms", d["roofline"].get("fp64_pass_ms"),
              "e2e", round(e.get("value",0)/1e6,2), (e.get("pipeline") or {}).get("kernel_ms"), "resident_compact", round(rc.get("value",0)/1e6,2), "parity", p.get("pass"), p.get("count_mismatch"), "launches", d.get("gpu_launches"))
    except Exception as ex: print(f, "failed", ex)
PY
}
timeout 200 python profiles/tools/e2e_compact_sweep.py > $O/e2e_compact_sweep.jsonl 2> $O/e2e_compact_sweep.err; cut -c1-250 $O/e2e_compact_sweep.jsonl; tail -3 $O/e2e_compact_sweep.err
timeout 300 python bench.py --no-cpu-baseline > $O/bench_quadrotor_order.json 2> $O/bench.err
timeout 300 python bench.py --no-cpu-baseline --order 0 > $O/bench_quadrotor_indexorder.json 2>> $O/bench.err
timeout 300 python bench.py --config cartpole --no-cpu-baseline > $O/bench_cartpole_order.json 2>> $O/bench.err
timeout 300 python bench.py --config cartpole --no-cpu-baseline --order 0 > $O/bench_cartpole_indexorder.json 2>> $O/bench.err
tail -3 $O/bench.err
show $O/bench_*.json
(time timeout 600 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
