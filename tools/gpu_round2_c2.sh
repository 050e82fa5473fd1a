#!/bin/bash
# collecting sweep with a bounds-free vector loop + one checked tail iteration: timing against chunk_select 117.9 ms
# per 64-baseline step + parity
set -x
mkdir -p gpurun_out
export AB_ARGS="--baselines 64 --parity-planes 2"
tools/gpu_ab.sh col "TC_X=1"
grep -o '"parity_check": {[^}]*}' gpurun_out/ab_col.json | cut -c1-140
