#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fm or call_order" > gpurun_out/r45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r45_pytest.log
tail -5 gpurun_out/r45_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r45_bench_fm.json 2> gpurun_out/r45_bench_fm.log; cut -c1-300 gpurun_out/r45_bench_fm.json
timeout 900 python bench.py --workload fm_k64_5Mx500Kx32c_125M_per_gpu --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r45_bench_fm_125M_n1.json 2> gpurun_out/r45_bench_fm_125M_n1.log; cut -c1-300 gpurun_out/r45_bench_fm_125M_n1.json
