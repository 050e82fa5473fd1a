#!/bin/bash
# B5T (TMA) second-axis filter: parity, per-radius times, ncu capture, companions
set -x
mkdir -p gpurun_out
TC_FILTER_TRACE=1 timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background or filter or sum_threshold or golden" > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e.log
grep -c "b5t filter" gpurun_out/pytest_e.log; grep -v "filter:" gpurun_out/pytest_e.log | tail -15
P="python tools/filter_probe.py 16 512 4096"
timeout 300 $P > gpurun_out/probe_b5t.json 2>&1
TC_FILTER_NO_TMA=1 timeout 300 $P > gpurun_out/probe_b5.json 2>&1
cat gpurun_out/probe_b5t.json | grep -v summary
Q="python tools/filter_probe.py 16 512 4096 10,8 43,34"
timeout 300 $Q > gpurun_out/plain_q.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box5t -s 1 -c 1 -o gpurun_out/r02_box5t_r8 $Q > gpurun_out/ncu_q.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box5t -s 4 -c 1 -o gpurun_out/r02_box5t_r34 $Q > gpurun_out/ncu_q2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_e2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e2.log
tail -3 gpurun_out/pytest_e2.log
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-light"
timeout 300 python bench.py $B > gpurun_out/bench_b5t.json 2> gpurun_out/bench_b5t.err
tail -c 1500 gpurun_out/bench_b5t.json
timeout 300 python tools/companions.py > gpurun_out/companions2.json 2> gpurun_out/companions2.err
grep -v summary gpurun_out/companions2.json | cut -c1-250
