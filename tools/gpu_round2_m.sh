#!/bin/bash
# line median with pointer-walk loads and an approximate pivot quotient: timing against line_median 48.3 ms per
# 64-baseline step + parity
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh med "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_med.json | cut -c1-140
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "median or sum_threshold" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_m.log
