#!/bin/bash
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh c4 "TC_BRK_CLUSTER=4" c1 "TC_BRK_CLUSTER=1" c4b "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_c4.json | cut -c1-140
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "median or background or sum_threshold or uvcontsub or golden or missed" > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_w.log
