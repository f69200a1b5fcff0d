python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_2.log; tail -5 gpurun_out/r02_pytest_2.log
./profiles/microbench/fp64_bench > gpurun_out/r02_fp64_bench.txt 2>&1; cat gpurun_out/r02_fp64_bench.txt
for f in -1 0 8 12 16 24 32; do python bench.py --fixer-sms $f --no-cpu-baseline --parity-n 0 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(dict(fixer_sms=d['run']['fixer_sms'], value=round(d['value']/1e6,2), ms=round(d['ms_per_step'],3), e2e=round(d['e2e']['value']/1e6,2), e2e_ms=round(d['e2e']['ms_per_step'],2), marked=d['fp64_resolved'], kernel=d['roofline']['kernel'])))" | tee -a gpurun_out/r02_fixer_sweep.jsonl; done
python bench.py > gpurun_out/r02_bench_exact2.json 2> gpurun_out/r02_bench_exact2.err; cat gpurun_out/r02_bench_exact2.json
