#!/bin/bash
# A/B of the two-line-groups-per-warp form of the B5 filter (TC_B5_U) on the default bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background" > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_j.log
TC_B5_U=2 TC_B5_B_MAXR=300 TC_B5_A_MINR=18 TC_FILTER_NO_TMA=1 timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background" > gpurun_out/pytest_j2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_j2.log
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh u1 "TC_B5_U=1" \
  u2 "TC_X=1" \
  u1_b300 "TC_B5_U=1 TC_B5_B_MAXR=300" \
  u2_b300 "TC_B5_B_MAXR=300" \
  u2all_b300 "TC_B5_U=2 TC_B5_B_MAXR=300 TC_FILTER_NO_TMA=1" \
  u2_b300_a18 "TC_B5_B_MAXR=300 TC_B5_A_MINR=18" \
  u2_b300_a10 "TC_B5_B_MAXR=300 TC_B5_A_MINR=10"
export AB_ARGS="--baselines 128"
TC_WORKSPACE_MB=120000 tools/gpu_ab.sh bl128 "TC_B5_U=1"
