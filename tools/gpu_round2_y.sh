#!/bin/bash
# rows -> pack -> flag -> unpack through the pipelined executor (pre / post hooks): test + configs 2 and 4 end to end
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "pipelined" > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_y.log
for c in 4 2; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y_c$c.json 2> gpurun_out/bench_y_c$c.err; echo "bench c$c rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_y_c$c.json') if l.startswith('{')][-1])
    print('c$c value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms', round(d['ms_per_step'],1), 'parity', d.get('parity_check',{}).get('ndiff'), d['e2e'].get('result_shape'))
except Exception as e: print('c$c failed', e)
PY
  tail -3 gpurun_out/bench_y_c$c.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-light --parity-planes 0 > gpurun_out/bench_y_c1.json 2> gpurun_out/bench_y_c1.err; echo "bench c1 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_y_c1.json') if l.startswith('{')][-1])
print('c1 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms', round(d['ms_per_step'],1))
PY
