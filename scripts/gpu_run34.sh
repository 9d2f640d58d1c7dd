#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fm_dense_reduce -s 8 -c 1 -o gpurun_out/prof_r34_fm_dense python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r34_ncu.log 2>&1
tail -2 gpurun_out/r34_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fm_piece_reduce -s 12 -c 2 -o gpurun_out/prof_r34_fm_piece python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r34_ncu2.log 2>&1
tail -2 gpurun_out/r34_ncu2.log
