#!/bin/bash
# Round-2 follow-up run on one B200 (gpurun -- bash profiles/tools/r02_compact.sh): compact host I/O as one launch chain behind an
# arrival watermark, compact reference / first control handled inside the kernels.  The new tests first (short timeout), then the
# whole GPU suite, the default bench line and the end-to-end sweep over the pipelines.  Every step under its own timeout.
set -u
O=gpurun_out/r02_compact; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "compact") > $O/pytest_compact.log 2>&1; echo "compact rc=$?" >> $O/pytest_compact.log; tail -5 $O/pytest_compact.log
timeout 200 python profiles/tools/e2e_compact_sweep.py > $O/e2e_compact_sweep.jsonl 2> $O/e2e_compact_sweep.err; cut -c1-220 $O/e2e_compact_sweep.jsonl; tail -3 $O/e2e_compact_sweep.err
timeout 400 python bench.py > $O/bench_quadrotor_n1.json 2> $O/bench_quadrotor_n1.err; tail -3 $O/bench_quadrotor_n1.err
python - $O/bench_quadrotor_n1.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); p=d.get("parity") or {}; e=d.get("e2e") or {}; o=d.get("e2e_other") or {}
    print("value", round(d["value"]/1e6,3), "M/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), "e2e", round(e.get("value",0)/1e6,2), e.get("pipeline"),
          "other", round(o.get("value",0)/1e6,2), "parity", p.get("pass"), p.get("count_mismatch"), p.get("max_abs_du"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex: print("bench failed", ex)
PY
(time timeout 600 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
