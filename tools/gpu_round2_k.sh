#!/bin/bash
# select tail fallback + small-range update + ST scan widening (TC_ST_V3) + non-negative widening experiment
set -x
mkdir -p gpurun_out
K="median or background or sum_threshold or uvcontsub or golden or select"
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "$K" > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_k.log
TC_ST_V3=1 TC_BRK_K=0.0 timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "$K" > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_k2.log
export AB_ARGS="--baselines 32"
tools/gpu_ab.sh base "TC_X=1" v3 "TC_ST_V3=1"
cp tricolour_b200/libtricolour_b200.so /tmp/lib_main.so
cp tools/_var/libnonneg.so tricolour_b200/libtricolour_b200.so
export AB_ARGS="--baselines 32 --parity-planes 4"
tools/gpu_ab.sh nonneg "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_nonneg.json
cp /tmp/lib_main.so tricolour_b200/libtricolour_b200.so
