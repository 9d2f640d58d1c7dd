#!/bin/bash
# the final default kernel (F compiled in, 256-bit accesses): launch list of the bench command + one ncu --set full capture
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r54_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r54_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgd_flagged -s 1 -c 1 -o gpurun_out/prof_r54_sgd_flagged_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r54_ncu2.log 2>&1
tail -1 gpurun_out/r54_ncu2.log
