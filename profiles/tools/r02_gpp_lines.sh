O=gpurun_out/r02_gpp2; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpp -s 2 -c 1 -f -o $O/full_gpp python bench.py --precision 64 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --parity-n 0 --batch 65536 > $O/full.log 2>&1
python profiles/tools/line_hist.py $O/full_gpp.ncu-rep 60 > $O/line_hist.txt; head -64 $O/line_hist.txt
ncu -i $O/full_gpp.ncu-rep --page source --csv > $O/source.csv 2>/dev/null
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_gpp2/source.csv', encoding='latin-1')))
hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
print([h for h in hdr if 'stall' in h.lower() or 'Sampl' in h][:40])
PY
rm -f $O/*.ncu-rep
