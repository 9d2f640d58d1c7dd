#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fm" > gpurun_out/r38_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r38_pytest.log
tail -12 gpurun_out/r38_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r38_bench_fm.json 2> gpurun_out/r38_bench_fm.log; tail -3 gpurun_out/r38_bench_fm.log; cat gpurun_out/r38_bench_fm.json
