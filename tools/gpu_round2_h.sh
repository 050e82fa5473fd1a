#!/bin/bash
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh plain "TC_FILTER_NO_TMA=1" safe "TC_FILTER_NO_TMA=1 TC_B5_DRAIN=4" persample "TC_FILTER_NO_TMA=1 TC_B5_DRAIN=8" nob5 "TC_FILTER_NO_B5=1"
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 --baselines 16"
M="smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio"
for v in 4 8 0; do
TC_FILTER_NO_TMA=1 TC_B5_DRAIN=$v timeout 300 $B > gpurun_out/plain_h.log 2>&1 && \
TC_FILTER_NO_TMA=1 TC_B5_DRAIN=$v timeout 900 ncu --metrics $M --clock-control none -k regex:k_box5b -s 30 -c 12 --csv --log-file gpurun_out/inst_h$v.csv $B > gpurun_out/ncu_h.log 2>&1
done
TC_FILTER_NO_B5=1 timeout 900 ncu --metrics $M --clock-control none -k regex:k_box4 -s 30 -c 12 --csv --log-file gpurun_out/inst_hk4.csv $B > gpurun_out/ncu_h.log 2>&1
python - <<'PY'
import csv
for f in ['h4','h8','h0','hk4']:
    rows=[r for r in csv.reader(l for l in open('gpurun_out/inst_%s.csv'%f) if not l.startswith('=='))]
    h=rows[0]; cur={}
    for r in rows[1:]:
        d=dict(zip(h,r)); cur.setdefault(d['ID'],{'k':d['Kernel Name'][:28]})[d['Metric Name'].split('.')[0][-22:]]=d['Metric Value']
    for i,v in cur.items(): print(f,i,v)
PY
