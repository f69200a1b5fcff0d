set -u
O=gpurun_out/r01f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or mixed" > $O/pytest_streamed.log 2>&1; tail -3 $O/pytest_streamed.log
timeout 600 python profiles/tools/e2e_sweep.py quadrotor 0 > $O/e2e_quadrotor_v0.jsonl 2>&1; cat $O/e2e_quadrotor_v0.jsonl | cut -c1-250
timeout 600 python profiles/tools/e2e_sweep.py quadrotor 5 > $O/e2e_quadrotor_v5.jsonl 2>&1; cat $O/e2e_quadrotor_v5.jsonl | cut -c1-250
timeout 600 python profiles/tools/e2e_sweep.py cartpole 0 > $O/e2e_cartpole_v0.jsonl 2>&1; cat $O/e2e_cartpole_v0.jsonl | cut -c1-250
timeout 600 python profiles/tools/e2e_sweep.py rocket 0 > $O/e2e_rocket_v0.jsonl 2>&1; cat $O/e2e_rocket_v0.jsonl | cut -c1-250
