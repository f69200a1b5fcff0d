#!/bin/bash
# One `ncu --set full` capture of the solve kernel of a bench config, with its summary, SASS opcode histogram and the
# per-source-line export.  Usage (GPU box, repo root): bash profiles/tools/ncu_one.sh <config> <tag> [extra bench args]
set -u
CFG=$1; TAG=$2; shift 2
O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tpp -s 3 -c 1 -f -o $O/full_$CFG \
   python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/full_$CFG.log 2>&1
python profiles/tools/ncu_summary.py $O/full_$CFG.ncu-rep > $O/ncu_full_$CFG.json 2>> $O/full_$CFG.log
ncu -i $O/full_$CFG.ncu-rep --page source --csv > $O/full_${CFG}_source.csv 2>> $O/full_$CFG.log
python profiles/tools/sass_hist.py $O/full_${CFG}_source.csv > $O/sass_hist_$CFG.txt 2>> $O/full_$CFG.log
head -30 $O/sass_hist_$CFG.txt
grep -E "issue_active|registers_per_thread\"|time_duration|stalled_(short|long|wait|math|not_sel|dispatch|no_inst|branch|mio)" $O/ncu_full_$CFG.json
