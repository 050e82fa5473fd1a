#!/bin/bash
# B5 with TMA output tiles (k_box5bo): parity, then A/B against the plain-store form
set -x
mkdir -p gpurun_out
TC_B5O_MAXR=300 TC_FILTER_TRACE=1 timeout 600 python -m pytest tests/test_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -s -k "gaussian or background or config0 or fullsize or plane" > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"
grep -c "b5o filter" gpurun_out/pytest_o.log; grep -E "passed|failed|error" gpurun_out/pytest_o.log | tail -3
export AB_ARGS="--baselines 32 --parity-planes 2"
tools/gpu_ab.sh base "TC_X=1" b5o_all "TC_B5O_MAXR=300" b5o_ge13 "TC_B5O_MAXR=300 TC_B5O_MINR=13" b5o_le12 "TC_B5O_MAXR=12"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_b5o_all.json | cut -c1-160
