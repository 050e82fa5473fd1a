#!/bin/bash
# A/B of runtime knobs on the default bench: tools/gpu_ab.sh name1 "ENV=.. ENV=.." name2 "..." ...
mkdir -p gpurun_out
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0"
while [ $# -gt 1 ]; do
  name=$1; envs=$2; shift 2
  env $envs timeout 300 python bench.py $Q $AB_ARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/ab_$name.json') if l.startswith('{')][-1])
    print('$name', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()})
except Exception as e: print('$name failed rc=$rc', e)
PY
done
