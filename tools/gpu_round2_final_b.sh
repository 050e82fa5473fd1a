#!/bin/bash
# end of round 2: whole GPU test suite, smoke(), the default bench the way the driver runs it
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c1_driver.json 2> gpurun_out/bench_c1_driver.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_c1_driver.json') if l.startswith('{')][-1])
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'ms', round(d['ms_per_step'],1), 'launches', d['gpu_launches'], 'parity', d['parity_check']['ndiff'], 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['clocks'])
print({k:round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()})
print(d['extra'])
PY
