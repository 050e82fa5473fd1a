#!/bin/bash
# end of round: the two subprocess tests on the GPU, the default bench the way the driver runs it, recapture of the select kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "missed_bracket or tma_form or median_abs" > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_v.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c1_driver.json 2> gpurun_out/bench_c1_driver.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_c1_driver.json
bash tools/gpu_round2_profile.sh other
