set -u
O=gpurun_out/r01c; mkdir -p $O
python profiles/microbench/pcie_bench.py > $O/pcie.json 2>&1; cat $O/pcie.json
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed" > $O/pytest_streamed.log 2>&1; tail -5 $O/pytest_streamed.log
timeout 600 python profiles/tools/e2e_sweep.py quadrotor 0 > $O/e2e_quadrotor_v0.jsonl 2>&1; cat $O/e2e_quadrotor_v0.jsonl | cut -c1-250
timeout 600 python profiles/tools/e2e_sweep.py quadrotor 5 > $O/e2e_quadrotor_v5.jsonl 2>&1; cat $O/e2e_quadrotor_v5.jsonl | cut -c1-250
timeout 300 python bench.py --variant 6 --no-cpu-baseline --no-e2e > $O/bench_quadrotor_v6.json 2>&1; cut -c1-200 $O/bench_quadrotor_v6.json
timeout 300 python bench.py --config cartpole --variant 6 --no-cpu-baseline --no-e2e > $O/bench_cartpole_v6.json 2>&1; cut -c1-200 $O/bench_cartpole_v6.json
timeout 300 python bench.py --no-cpu-baseline > $O/bench_quadrotor.json 2>&1; cut -c1-200 $O/bench_quadrotor.json
