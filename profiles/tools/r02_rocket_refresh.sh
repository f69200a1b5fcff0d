O=gpurun_out/r02_final; mkdir -p $O
timeout 400 python bench.py --config rocket --cpu-seconds 4 > $O/bench_rocket_n1.json 2> $O/bench_rocket_n1.err
python -c "
import json; d=json.load(open('$O/bench_rocket_n1.json')); print(round(d['value']/1e6,2), round(d['ms_per_step'],3), d['roofline']['kernel_ms'], d['roofline']['fp64_pass_ms'], round(d['roofline']['frac'],3), round(d['e2e']['value']/1e6,2), d['parity']['pass'], d['cpu_baseline']['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp4 -s 3 -c 1 -f -o $O/full_rocket_tpp4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 --config rocket > $O/full_rocket_tpp4.log 2>&1
python profiles/tools/ncu_summary.py $O/full_rocket_tpp4.ncu-rep > $O/ncu_full_rocket_tpp4.json
ncu -i $O/full_rocket_tpp4.ncu-rep --page source --csv > $O/src.csv; python profiles/tools/sass_hist.py $O/src.csv > $O/sass_hist_rocket_tpp4.txt; rm -f $O/src.csv $O/*.ncu-rep
grep -E "time_duration|issue_active|registers_per_thread\"|no_instruction|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|dram__bytes_(read|write).sum\"" $O/ncu_full_rocket_tpp4.json
head -8 $O/sass_hist_rocket_tpp4.txt
