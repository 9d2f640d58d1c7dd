#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload fm_k64_5Mx500Kx32c_125M_per_gpu --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r40_bench_fm_125M_n1.json 2> gpurun_out/r40_bench_fm_125M_n1.log; cut -c1-400 gpurun_out/r40_bench_fm_125M_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r40_launches_fm_125M.csv python bench.py --workload fm_k64_5Mx500Kx32c_125M_per_gpu --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r40_fm.log 2>&1
