#!/bin/bash
# select knobs at the default block size (64 baselines)
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64"
tools/gpu_ab.sh base "TC_X=1" t512 "TC_BRK_THREADS=512" notail "TC_BRK_TAIL_MAX=0" notail512 "TC_BRK_TAIL_MAX=0 TC_BRK_THREADS=512" base2 "TC_X=2"
