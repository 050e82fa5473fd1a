#!/bin/bash
# same-box comparison of library builds at the default bench settings (64 baselines, light workload included)
set -x
mkdir -p gpurun_out
run() {
  cp tools/_var/lib_$1.so tricolour_b200/libtricolour_b200.so
  timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --parity-planes 0 > gpurun_out/q_$2.json 2> gpurun_out/q_$2.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/q_$2.json') if l.startswith('{')][-1])
k=d['roofline']['kernel_ms_per_step']
print('$2', round(d['ms_per_step'],1), 'sum', round(sum(k.values()),1), 'light', round(d['extra']['light_workload']['value'],4), {a:round(b,1) for a,b in k.items()})
PY
}
run old old1
run head head1
run new new1
run old old2
run new new2
