#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r35_pytest.log
tail -5 gpurun_out/r35_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r35_bench_fm.json 2> gpurun_out/r35_bench_fm.log; cut -c1-330 gpurun_out/r35_bench_fm.json
CARS_FM_FUSE=0 timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r35_bench_fm_nofuse.json 2> gpurun_out/r35_bench_fm_nofuse.log; cut -c1-330 gpurun_out/r35_bench_fm_nofuse.json
