#!/bin/bash
# second GPU pass: new filter / median kernels: parity, then A/B timing of the variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q -k "gaussian or background or median or sum_threshold or golden" > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -5 gpurun_out/pytest_b.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_b2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b2.log
tail -5 gpurun_out/pytest_b2.log
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py $Q > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/ab_$name.json') if l.startswith('{')][-1])
    print('$name', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()})
except Exception as e: print('$name failed', e)
PY
}
run base TC_FILTER_NO_TPL=1 TC_MEDIAN_BITS=1
run med TC_FILTER_NO_TPL=1
run tpl TC_X=1
run tpl_b24 TC_TPL_B_MAXR=24
run tpl_b43 TC_TPL_B_MAXR=43 TC_TPL_B_MINW=2
run tpl_a28 TC_TPL_A_MAXR=28
run tpl_a17 TC_TPL_A_MAXR=17
run tpl_bonly TC_TPL_A_MAXR=1
run tpl_aonly TC_TPL_B_MAXR=1
