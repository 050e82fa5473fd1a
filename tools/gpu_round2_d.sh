#!/bin/bash
# B5 filter: parity, per-radius times against the k_filter2 forms, ncu capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background or filter or sum_threshold or golden" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -3 gpurun_out/pytest_d.log
P="python tools/filter_probe.py 16 512 4096"
$P > gpurun_out/probe_b5.json 2>&1
TC_FILTER_NO_B5=1 $P > gpurun_out/probe_nob5.json 2>&1
Q="python tools/filter_probe.py 16 512 4096 10,8 43,34"
$Q > gpurun_out/plain_q.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_box5 -s 2 -c 2 -o gpurun_out/r02_box5_r8 $Q > gpurun_out/ncu_q.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_box5 -s 8 -c 2 -o gpurun_out/r02_box5_r43 $Q > gpurun_out/ncu_q2.log 2>&1
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-light"
timeout 300 python bench.py $B > gpurun_out/bench_b5.json 2> gpurun_out/bench_b5.err
tail -c 1500 gpurun_out/bench_b5.json
