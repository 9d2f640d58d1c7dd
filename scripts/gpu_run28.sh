#!/bin/bash
# launch lists: default SGD bench (with the device-built schedule) and the FM bench, per-kernel time + DRAM bytes
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r28_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r28_b.log 2>&1
tail -2 gpurun_out/r28_b.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/r28_launches_fm.csv python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r28_fm.log 2>&1
tail -2 gpurun_out/r28_fm.log
