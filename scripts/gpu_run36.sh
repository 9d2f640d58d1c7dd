#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fm" > gpurun_out/r36_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r36_pytest.log
tail -3 gpurun_out/r36_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 > gpurun_out/r36_bench_fm.json 2> gpurun_out/r36_bench_fm.log; cat gpurun_out/r36_bench_fm.json
CARS_SCHED_TRACE=1 timeout 600 python bench.py > gpurun_out/r36_bench_default.json 2> gpurun_out/r36_bench_default.log; grep -v "^$" gpurun_out/r36_bench_default.log | tail -24; cat gpurun_out/r36_bench_default.json
