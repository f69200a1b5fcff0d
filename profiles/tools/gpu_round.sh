#!/bin/bash
# One GPU-box session of the round: GPU tests, bench lines for every config, the ncu launch list of the default bench
# command and one `ncu --set full` capture of the solve kernel.  Usage (from the repo root): gpurun -- bash profiles/tools/gpu_round.sh [tag]
set -u
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_quadrotor_n1.json 2> $O/bench_quadrotor_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_quadrotor_n1.err
for c in cartpole rocket quadrotor_adaptive; do
  timeout 600 python bench.py --config $c --cpu-seconds 4 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
done
timeout 600 python bench.py --scale 0.3 --no-cpu-baseline > $O/bench_quadrotor_easy_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 600 python bench.py --mixed 0.003 --no-cpu-baseline > $O/bench_quadrotor_mixed_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 600 python bench.py --precision 64 --steps 5 --no-cpu-baseline > $O/bench_quadrotor_fp64_n1.json 2>> $O/bench_quadrotor_n1.err
timeout 600 python bench.py --variant 5 --no-cpu-baseline > $O/bench_quadrotor_direct_n1.json 2>> $O/bench_quadrotor_n1.err
cat $O/bench_*_n1.json | cut -c1-400
# launch list of the bench command (cold-cache, serialised: the kernel's SHARE of the step is what must agree)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_quadrotor.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
# one full capture of the solve kernel (4th launch = first timed one)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tpp -s 3 -c 1 -f -o $O/full_quadrotor \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_quadrotor.log 2>&1
python profiles/tools/ncu_summary.py $O/full_quadrotor.ncu-rep > $O/ncu_full_quadrotor.json 2>> $O/full_quadrotor.log
ncu -i $O/full_quadrotor.ncu-rep --page source --csv > $O/full_quadrotor_source.csv 2>> $O/full_quadrotor.log
ls -la $O
