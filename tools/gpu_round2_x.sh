#!/bin/bash
# A/B: thresholds applied by the collecting sweep's tail (default) against the separate k_sel_update launch
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh tail "TC_X=1" sep "TC_BRK_UPDATE_IN_TAIL=0" tail1024 "TC_BRK_THREADS=1024" tail_b "TC_X=2"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_tail.json | cut -c1-140
grep -o '"gpu_launches": [0-9]*' gpurun_out/ab_tail.json gpurun_out/ab_sep.json
export AB_ARGS="--config 3 --parity-planes 2"
tools/gpu_ab.sh c3tail "TC_X=1" c3sep "TC_BRK_UPDATE_IN_TAIL=0"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_c3tail.json | cut -c1-140
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "median or background or sum_threshold or uvcontsub or golden or missed" > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_x.log
