#!/bin/bash
# source-level (SASS + stall samples) captures of the top kernels of a 16-baseline step; only the CSV pages come back
set -x
mkdir -p gpurun_out/src
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 --baselines 16"
timeout 300 $S > gpurun_out/plain_n.log 2>&1 || exit 1
cap() { # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/cap_$1 $S > gpurun_out/src/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/cap_$1.ncu-rep --page raw --csv > gpurun_out/src/$1_raw.csv 2>/dev/null
  for i in $(seq 0 $(($4 - 1))); do
    ncu -i /tmp/cap_$1.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/src/$1_src$i.csv 2>/dev/null
  done
}
cap box5b k_box5b 6 2
cap t4a k_box_t4a 2 2
cap box5a k_box5a 0 1
cap box8 k_box8 0 1
cap collect k_brk_collect 0 1
cap collect_uv k_brk_collect 26 1
cap scan k_st_scan 1 2
cap median k_line_median2 1 1
gzip -f gpurun_out/src/*_src*.csv
du -sh gpurun_out/src
