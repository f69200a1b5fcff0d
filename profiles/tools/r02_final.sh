#!/bin/bash
# Round-2 evidence run on one B200 (gpurun -- bash profiles/tools/r02_final.sh): GPU tests, bench lines of every config and mode, the
# launch lists (bench, smoke) and one `ncu --set full` capture per shipped kernel.  Every step under its own timeout.
set -u
O=gpurun_out/r02_final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 900 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_quadrotor_n1.json 2> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_quadrotor_n1.err
for c in cartpole rocket quadrotor_adaptive; do timeout 400 python bench.py --config $c --cpu-seconds 4 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err; done
timeout 300 python bench.py --scale 0.3 --no-cpu-baseline > $O/bench_quadrotor_easy_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --mixed 0 --no-cpu-baseline > $O/bench_quadrotor_plainfp32_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --mixed 0 --variant 7 --no-cpu-baseline > $O/bench_quadrotor_plainfp32_recursion_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --precision 64 --steps 5 --no-cpu-baseline > $O/bench_quadrotor_fp64_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --precision 64 --variant 6 --steps 5 --no-cpu-baseline --no-e2e > $O/bench_quadrotor_fp64_tpp2_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --config cartpole --scale 0.3 --no-cpu-baseline > $O/bench_cartpole_easy_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 300 python bench.py --config rocket --mixed 0 --variant 3 --no-cpu-baseline --no-e2e > $O/bench_rocket_fp32_tpp3_n1.json 2>> $O/bench_quadrotor_n1.err
for f in $O/bench_*_n1.json $O/bench_reference_arm.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); p=d.get("parity") or {}; e=d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], round(d["value"]/1e6,3), "M/s", round(d["ms_per_step"],3), "ms frac", round((d.get("roofline") or {}).get("frac",0),3), "e2e", round(e.get("value",0)/1e6,2), "parity", p.get("pass"), p.get("count_mismatch"), p.get("max_abs_du"), (d.get("roofline") or {}).get("kernel"))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
done
# launch list of the bench command (cold-cache, serialised: the kernel's SHARE of the step is what must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_quadrotor.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 > $O/launches_bench.log 2>&1
# launch list of smoke()
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_smoke.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_under_ncu.log 2>&1; tail -3 $O/smoke_under_ncu.log
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/full_$name \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 "$@" > $O/full_$name.log 2>&1
  python profiles/tools/ncu_summary.py $O/full_$name.ncu-rep > $O/ncu_full_$name.json 2>> $O/full_$name.log
  ncu -i $O/full_$name.ncu-rep --page source --csv > $O/full_${name}_source.csv 2>> $O/full_$name.log
  python profiles/tools/sass_hist.py $O/full_${name}_source.csv > $O/sass_hist_$name.txt 2>> $O/full_$name.log
  rm -f $O/full_${name}_source.csv
  echo "== $name"; grep -E "Kernel Name|time_duration|issue_active|registers_per_thread\"|dram__bytes|warps_active|pipe_fma_cycles|pipe_fp64|pipe_alu" $O/ncu_full_$name.json | cut -c1-160
}
cap quadrotor_tpp3 tpp3 3
cap quadrotor_gpp_fixer gpp 3
cap quadrotor_gpp_fp64 gpp 3 --precision 64 --batch 262144
cap cartpole_tpp3 tpp3 3 --config cartpole
cap rocket_tpp4 tpp4 3 --config rocket
cap adaptive_gpp gpp 3 --config quadrotor_adaptive --batch 262144
rm -f $O/*.ncu-rep
for pr in 64 32; do timeout 200 python profiles/tools/session_bench.py quadrotor 65536 20 $pr; done > $O/session_bench.jsonl 2>&1
timeout 200 python profiles/tools/session_bench.py quadrotor 1048576 10 64 >> $O/session_bench.jsonl 2>&1
timeout 200 python profiles/tools/session_bench.py cartpole 65536 20 64 >> $O/session_bench.jsonl 2>&1
cat $O/session_bench.jsonl
timeout 300 python profiles/tools/latency.py > $O/latency.jsonl 2>&1; cut -c1-200 $O/latency.jsonl
timeout 300 python profiles/tools/e2e_compact_sweep.py > $O/e2e_compact_sweep.jsonl 2>&1
ls $O | head -80
