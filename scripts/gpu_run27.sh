#!/bin/bash
# device-built schedule + staged copies + CAMF_CUCI: tests, then the default bench (pinned e2e) and the pageable e2e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r27_gpu.txt
nproc >> gpurun_out/r27_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r27_pytest.log
tail -5 gpurun_out/r27_pytest.log
timeout 600 python bench.py > gpurun_out/r27_bench_default.json 2> gpurun_out/r27_bench_default.log; tail -4 gpurun_out/r27_bench_default.log; cat gpurun_out/r27_bench_default.json
CARS_BENCH_PAGEABLE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r27_bench_pageable.json 2> gpurun_out/r27_bench_pageable.log; tail -3 gpurun_out/r27_bench_pageable.log; cat gpurun_out/r27_bench_pageable.json
