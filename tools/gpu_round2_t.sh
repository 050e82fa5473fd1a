#!/bin/bash
# (1) does the NVML sampler thread slow the timed region?  (2) new thread-per-line first-axis loop against the old one
set -x
mkdir -p gpurun_out
Q="--no-e2e --no-cpu-baseline --no-light --parity-planes 0"
run() { # name, lib, args...
  name=$1; lib=$2; shift 2
  cp tools/_var/lib_$lib.so tricolour_b200/libtricolour_b200.so
  timeout 600 python bench.py $Q "$@" > gpurun_out/t_$name.json 2> gpurun_out/t_$name.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/t_$name.json') if l.startswith('{')][-1])
k=d['roofline']['kernel_ms_per_step']
print('$name', round(d['ms_per_step'],1), 'sum', round(sum(k.values()),1), 'axis0', round(k['box_filter_axis0'],1), 'sel', round(k['chunk_select'],1), d['clocks'])
PY
}
run base_33 base --steps 3 --warmup 3
run base_33_slow base --steps 3 --warmup 3 --clock-interval 1.5
run base_21 base --steps 2 --warmup 1
run new_33_slow new --steps 3 --warmup 3 --clock-interval 1.5
run new_21 new --steps 2 --warmup 1
run base_33b base --steps 3 --warmup 3
cp tools/_var/lib_new.so tricolour_b200/libtricolour_b200.so
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background or golden" > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_t.log
