#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_serial -s 1 -c 1 -o gpurun_out/prof_r30_serial_pipelined python scripts/config2.py 3 > gpurun_out/r30_ncu_serial.log 2>&1
tail -2 gpurun_out/r30_ncu_serial.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r30_launches_fm.csv python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r30_fm.log 2>&1
tail -1 gpurun_out/r30_fm.log | cut -c1-200
