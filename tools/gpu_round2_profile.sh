#!/bin/bash
# round-2 profile pass (gpurun_out must stay below 64 MiB per call, so it comes in three parts):
#   lists : launch list of the default bench command + DRAM bytes / instructions of every launch of one 16-baseline step
#   box   : --set full capture of the box filter kernels
#   other : --set full capture of the select / scan / median / transpose kernels
set -x
mkdir -p gpurun_out
S="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 --baselines 16"
case "$1" in
lists)
  B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-light --parity-planes 0"
  timeout 600 $B > gpurun_out/plain_bench.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
  timeout 300 $S > gpurun_out/plain_small.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_all.csv $S > gpurun_out/ncu_traffic.log 2>&1
  echo "traffic rc=$?"
  gzip -f gpurun_out/r02_launches.csv gpurun_out/r02_traffic_all.csv
  ;;
box)
  timeout 300 $S > gpurun_out/plain_small.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_box4|k_box5a|k_box5b|k_box8|k_box_t4a" -s 44 -c 8 -o gpurun_out/r02_box_filter $S > gpurun_out/ncu_full1.log 2>&1
  echo "full1 rc=$?"
  ;;
other)
  timeout 300 $S > gpurun_out/plain_small.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none -k regex:"k_brk_collect|k_st_scan|k_line_median|k_sel_update|k_brk_sample|k_interp|k_transpose" -s 70 -c 10 -o gpurun_out/r02_other $S > gpurun_out/ncu_full2.log 2>&1
  echo "full2 rc=$?"
  ;;
esac
du -sh gpurun_out
