#!/bin/bash
# last pass of round 2: A/B of the vector transpose (TC_TRANSPOSE_GENERIC=1 = the 32 x 32 kernel), then the whole
# GPU test suite, smoke() and the default bench the way the driver runs it
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh tr_new "TC_X=1" tr_old "TC_TRANSPOSE_GENERIC=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_tr_new.json | cut -c1-140
bash tools/gpu_round2_final_b.sh
