#!/bin/bash
# B5 input loads coalesced per stream (lane permutation before the publish): timing against the numbers of
# tools/gpu_round2_x.sh (box_filter 233.7 / axis0 160.3 / single axis 8.8 ms per 64-baseline step) + parity
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh coal "TC_X=1" coal2 "TC_X=2"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_coal.json | cut -c1-140
timeout 900 python -m pytest tests/test_parity.py tests/test_gpu_pipeline.py -m gpu -x -q -k "filter or background or golden or tma or flagger" > gpurun_out/pytest_z.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_z.log
