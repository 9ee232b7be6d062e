#!/bin/bash
# round-2 run L: wavefront integrator with the state-machine extend stage on C5
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine or tessellated or c5_two or wavefront" > gpurun_out/gputest_l.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_l.log
timeout 900 python tools/sweep_tune.py c5 4 "" "5=2" 2>&1 | tee gpurun_out/sweep_c5_l.txt
SWEEP_INTEGRATOR=wavefront timeout 900 python tools/sweep_tune.py c5 4 "13=1" "" "7=4" "7=12" "7=16" "11=6" "11=14" "7=12,11=14" 2>&1 | tee -a gpurun_out/sweep_c5_l.txt
SWEEP_INTEGRATOR=wavefront SWEEP_FPB=2 timeout 900 python tools/sweep_tune.py c5 4 "" 2>&1 | tee -a gpurun_out/sweep_c5_l.txt
SWEEP_INTEGRATOR=wavefront timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5_wavefront_sm.csv python tools/sweep_tune.py c5 1 "" > gpurun_out/ncu_list_l.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c5_wavefront_sm.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    agg.setdefault((int(r[0]), r[4][:40]),{})[r[12]]=r[14]
for (i,k),m in sorted(agg.items()):
    if i>=40: break
    print(i,k,m.get('gpu__time_duration.sum'),m.get('smsp__inst_executed.sum'),m.get('smsp__thread_inst_executed_per_inst_executed.ratio'),m.get('smsp__issue_active.avg.pct_of_peak_sustained_active'))
PY
