#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fm" > gpurun_out/r33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r33_pytest.log
tail -6 gpurun_out/r33_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r33_bench_fm.json 2> gpurun_out/r33_bench_fm.log; cut -c1-330 gpurun_out/r33_bench_fm.json
CARS_FM_FUSE=0 timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r33_bench_fm_nofuse.json 2> gpurun_out/r33_bench_fm_nofuse.log; cut -c1-330 gpurun_out/r33_bench_fm_nofuse.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r33_launches_fm.csv python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r33_fm.log 2>&1
