#!/bin/bash
# two GPUs: configs[4] (rows -> pack -> stats -> strategy -> stats -> unpack, stats all-reduce per block) through the
# pipelined executor's packing hooks
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_c4.json 2> gpurun_out/bench_2gpu_c4.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_2gpu_c4.json') if l.startswith('{')][-1])
    print('c4 x2 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('stats_allreduce'), 'parity', d.get('parity_check',{}).get('ndiff'))
except Exception as e: print('failed', e)
PY
tail -5 gpurun_out/bench_2gpu_c4.err
