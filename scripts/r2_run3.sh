#!/bin/bash
# round 2, GPU call 3: full GPU suite (reference-bytecode vectors included), 100 M lines with the pool allocator,
# ncu captures of the FAST kernel (uniform and Zipf) + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -s > gpurun_out/r2c_pytest_fast.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_fast.log; tail -4 gpurun_out/r2c_pytest_fast.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fast_gpu.py > gpurun_out/r2c_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_all.log; tail -4 gpurun_out/r2c_pytest_all.log
B=camf_ci_f64_1Mx100Kx32c_100M
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_exact100M.json 2> gpurun_out/r2c_exact100M.log
timeout 900 python bench.py --workload ${B}_zipf1.0 --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_fast100Mz.json 2> gpurun_out/r2c_fast100Mz.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_fast -s 1 -c 1 -o gpurun_out/prof_r2c_fast_uniform python bench.py --workload $B --mode fast --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_fast -s 1 -c 1 -o gpurun_out/prof_r2c_fast_zipf python bench.py --workload ${B}_zipf1.0 --mode fast --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches_fast.csv python bench.py --workload $B --mode fast --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_ncu3.log 2>&1
for f in gpurun_out/r2c_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d.get("parity"), d["e2e"]["value"], d["e2e"]["seconds"], d["e2e_pageable"]["value"], d["config"].get("fast_min_item_scale"), d.get("cpu_baseline"))
except Exception as e:
    print("ERR", e)
PY
done
