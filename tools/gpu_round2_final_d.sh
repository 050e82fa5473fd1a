#!/bin/bash
# long-line median (32768-channel mode) with pointer-walk sweeps: parity of the paths that reach it + configs[2]
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "median or config2 or config3 or sum_threshold" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_d.log
timeout 300 python bench.py --config 2 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_d_c2.json 2> gpurun_out/bench_d_c2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_d_c2.json') if l.startswith('{')][-1])
print('c2 value', round(d['value'],3), 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_check',{}).get('ndiff'))
print({k:round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
