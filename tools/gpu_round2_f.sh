#!/bin/bash
set -x
mkdir -p gpurun_out
Q="python tools/filter_probe.py 16 512 4096 10,8 43,34 21,17"
timeout 300 $Q > gpurun_out/probe_f.json 2>&1; grep -v summary gpurun_out/probe_f.json
timeout 300 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background or filter" 2>&1 | tail -2
timeout 300 $Q > gpurun_out/plain_q.log 2>&1 && \
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_box5 --csv --log-file gpurun_out/inst_f.csv $Q > gpurun_out/ncu_f.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/inst_f.csv') if not l.startswith('=='))]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r))
    if d.get('Metric Name') in ('smsp__inst_executed.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active'):
        print(d['ID'], d['Kernel Name'][:40], d['Metric Name'][:40], d['Metric Value'])
PY
