#!/bin/bash
# SumThreshold scan with its samples fetched one block of 8 steps ahead: timing against st_scan 60.3 ms per
# 64-baseline step (tools/gpu_round2_z.sh) + parity
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh pref "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_pref.json | cut -c1-140
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "sum_threshold or golden or flagger" > gpurun_out/pytest_s2.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_s2.log
