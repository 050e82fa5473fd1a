#!/bin/bash
# third GPU pass: per-radius times of the filter forms + ncu captures of the thread-per-line kernels
set -x
mkdir -p gpurun_out
P="python tools/filter_probe.py 16 512 4096"
$P > gpurun_out/probe_default.json 2>&1
TC_TPL_A=1 TC_TPL_B=1 TC_TPL_B_MAXR=43 TC_TPL_B_MINW=2 $P > gpurun_out/probe_tpl.json 2>&1
Q="python tools/filter_probe.py 16 512 4096 10,8 43,34"
$Q > gpurun_out/plain_q.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_box -s 8 -c 4 -o gpurun_out/r02_box_default $Q > gpurun_out/ncu_q.log 2>&1
TC_TPL_A=1 TC_TPL_B=1 TC_TPL_B_MAXR=43 TC_TPL_B_MINW=2 $Q > gpurun_out/plain_q2.log 2>&1 && \
TC_TPL_A=1 TC_TPL_B=1 TC_TPL_B_MAXR=43 TC_TPL_B_MINW=2 ncu --set full --clock-control none --import-source on -k regex:k_box -s 8 -c 4 -o gpurun_out/r02_box_tpl $Q > gpurun_out/ncu_q2.log 2>&1
ls -la gpurun_out
