#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r51_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r51_pytest.log
tail -3 gpurun_out/r51_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r51_bench_default.json 2> gpurun_out/r51_bench_default.log; grep "bench\]" gpurun_out/r51_bench_default.log | tail -3; cat gpurun_out/r51_bench_default.json
for v in 7 3; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload camf_cu_f128_2Mx200Kx64c_200M 2>&1 >gpurun_out/r51_bench_cu_f128_v$v.json | grep "epochs in" | sed "s/^/F=128 variant=$v /"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgd_flagged -s 1 -c 1 -o gpurun_out/prof_r51_sgd_flagged_wide python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r51_ncu.log 2>&1
tail -1 gpurun_out/r51_ncu.log
