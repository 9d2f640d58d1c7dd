#!/bin/bash
mkdir -p gpurun_out
CARS_SCHED_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r32_bench_default.json 2> gpurun_out/r32_bench_default.log; grep -v "^$" gpurun_out/r32_bench_default.log | tail -30; grep -o '"e2e": {[^}]*}' gpurun_out/r32_bench_default.json
