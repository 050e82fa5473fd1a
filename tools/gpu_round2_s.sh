#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "median or background or sum_threshold or uvcontsub or golden or select" > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_s.log
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh b64 "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_b64.json | cut -c1-140
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh b32 "TC_X=1"
