#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "median or background or sum_threshold or uvcontsub or golden or select" > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_p.log
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh base "TC_X=1" t256 "TC_BRK_THREADS=256" t1024 "TC_BRK_THREADS=1024"
