#!/bin/bash
# B5 with prefetched residual samples / unrolled ping-pong loop / running fetch pointers
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_m.log
TC_B5_B_MAXR=300 TC_B5_A_MINR=4 TC_FILTER_NO_TMA=1 timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background" > gpurun_out/pytest_m2.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_m2.log
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh base "TC_X=1" b300 "TC_B5_B_MAXR=300" b300_a18 "TC_B5_B_MAXR=300 TC_B5_A_MINR=18" b300_a28 "TC_B5_B_MAXR=300 TC_B5_A_MINR=28" b300_notma "TC_B5_B_MAXR=300 TC_FILTER_NO_TMA=1"
