#!/bin/bash
# final pass A: the whole GPU test suite, every config, the companions
set -x
mkdir -p gpurun_out
python -c "import os; print(os.cpu_count())" > gpurun_out/cores.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "rc=$?"
tail -c 400 gpurun_out/bench_c1.json
for c in 0 3 4 2; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "config $c rc=$?"
  tail -c 300 gpurun_out/bench_c$c.err
done
timeout 300 python tools/companions.py > gpurun_out/companions.json 2> gpurun_out/companions.err; echo "companions rc=$?"
