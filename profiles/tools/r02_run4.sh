timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_4.log; cat gpurun_out/r02_pytest_4.log
for f in -1 0; do timeout 300 python bench.py --fixer-sms $f --no-cpu-baseline --steps 10 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fixer_sms $f value', round(d['value']/1e6,2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,2), d['e2e']['pipeline'], 'e2e_other', round(d['e2e_other']['value']/1e6,2), d['e2e_other']['pipeline'], d['parity'])"; done
mkdir -p gpurun_out/r02_gpp1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpp -s 2 -c 1 -f -o gpurun_out/r02_gpp1/full_gpp python bench.py --precision 64 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 --batch 262144 > gpurun_out/r02_gpp1/full.log 2>&1
python profiles/tools/ncu_summary.py gpurun_out/r02_gpp1/full_gpp.ncu-rep > gpurun_out/r02_gpp1/ncu_full_gpp.json
ncu -i gpurun_out/r02_gpp1/full_gpp.ncu-rep --page source --csv > gpurun_out/r02_gpp1/source.csv
python profiles/tools/sass_hist.py gpurun_out/r02_gpp1/source.csv | head -24
grep -E "issue_active|registers_per_thread\"|time_duration|warps_active|stalled_(short|long|wait|math|not_sel|dispatch|no_inst|branch|mio|barrier|lg)" gpurun_out/r02_gpp1/ncu_full_gpp.json
python profiles/tools/line_hist.py gpurun_out/r02_gpp1/full_gpp.ncu-rep 30 > gpurun_out/r02_gpp1/line_hist.txt; head -32 gpurun_out/r02_gpp1/line_hist.txt
