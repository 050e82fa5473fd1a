#!/bin/bash
# source-level captures of the remaining kernels (line median, uvcontsub, interpolation, sample, update)
set -x
mkdir -p gpurun_out/src2
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 --baselines 16"
timeout 300 $S > gpurun_out/plain_u.log 2>&1 || exit 1
cap() { # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/cap_$1 $S > gpurun_out/src2/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/cap_$1.ncu-rep --page raw --csv > gpurun_out/src2/$1_raw.csv 2>/dev/null
  for i in $(seq 0 $(($4 - 1))); do
    ncu -i /tmp/cap_$1.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/src2/$1_src$i.csv 2>/dev/null
  done
}
cap median k_line_median2 4 2
cap uvmean k_uv_mean 0 1
cap uvabs k_uv_absres 0 1
cap interp k_interp_nans_rows 1 1
cap sample k_brk_sample 6 1
cap update k_sel_update 6 1
cap prep k_prep_c64_tile 0 1
cap combine "k_combine_time_v16|k_dilate_rows_v16|k_finalize_flags_v4|k_colcnt_v4" 0 4
gzip -f gpurun_out/src2/*_src*.csv
du -sh gpurun_out/src2
