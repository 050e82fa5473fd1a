#!/bin/bash
# source-level capture of the B5 second-axis filter at large radii (where it loses to k_box4)
set -x
mkdir -p gpurun_out
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 --baselines 16"
export TC_B5_B_MAXR=300 TC_FILTER_NO_TMA=1
timeout 300 $S > gpurun_out/plain_l.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_box5b" -s 12 -c 2 -o gpurun_out/r02_b5_src $S > gpurun_out/ncu_l.log 2>&1
echo "rc=$?"
ls -la gpurun_out
