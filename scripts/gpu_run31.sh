#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r31_pytest.log
tail -6 gpurun_out/r31_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r31_bench_fm.json 2> gpurun_out/r31_bench_fm.log; cut -c1-330 gpurun_out/r31_bench_fm.json
for pl in 1 0; do CARS_SERIAL_PIPELINE=$pl python scripts/config2.py 10; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r31_bench_default.json 2> gpurun_out/r31_bench_default.log; tail -3 gpurun_out/r31_bench_default.log; grep -o '"e2e": {[^}]*}' gpurun_out/r31_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r31_launches_fm.csv python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r31_fm.log 2>&1
