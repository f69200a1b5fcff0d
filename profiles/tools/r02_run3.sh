# round-2 GPU run 3: every step under its own timeout
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r02_pytest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_3.log; tail -25 gpurun_out/r02_pytest_3.log
rm -f gpurun_out/r02_fixer_sweep.jsonl
for f in -1 0 12 16 24 32; do timeout 150 python bench.py --fixer-sms $f --no-cpu-baseline --parity-n 0 --steps 10 2>gpurun_out/r02_sweep_$f.err | python -c "
import json,sys
t=sys.stdin.read().strip()
if not t: print('fixer_sms $f: no output (timeout?)'); sys.exit()
d=json.loads(t); print(json.dumps(dict(fixer_sms=d['run']['fixer_sms'], value=round(d['value']/1e6,2), ms=round(d['ms_per_step'],3), e2e=round(d['e2e']['value']/1e6,2), e2e_ms=round(d['e2e']['ms_per_step'],2), e2e_other=round(d['e2e_other']['value']/1e6,2) if d.get('e2e_other') else None, marked=d['fp64_resolved'], kernel=d['roofline']['kernel'])))" | tee -a gpurun_out/r02_fixer_sweep.jsonl; done
timeout 400 python bench.py > gpurun_out/r02_bench_quadrotor.json 2> gpurun_out/r02_bench_quadrotor.err; cat gpurun_out/r02_bench_quadrotor.json
for args in "--config rocket" "--config rocket --mixed 0" "--config rocket --mixed 0 --variant 3"; do timeout 300 python bench.py $args --no-cpu-baseline 2>>gpurun_out/r02_bench_rocket.err | tee -a gpurun_out/r02_bench_rocket.jsonl | python -c "
import json,sys
t=sys.stdin.read().strip()
if not t: print('rocket $args: no output'); sys.exit()
d=json.loads(t); print(json.dumps(dict(args='$args', value=round(d['value']/1e6,2), ms=round(d['ms_per_step'],3), frac=round(d['roofline']['frac'],3), e2e=round(d['e2e']['value']/1e6,2), marked=d['fp64_resolved'], kernel=d['roofline']['kernel'], parity=d['parity'])))"; done
