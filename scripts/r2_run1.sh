#!/bin/bash
# round 2, GPU call 1: new FAST / parity tests, full GPU suite, EXACT + FAST bench lines, shape sweep, strict-acquire A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -s > gpurun_out/r2a_pytest_fast.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_fast.log; tail -5 gpurun_out/r2a_pytest_fast.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fast_gpu.py > gpurun_out/r2a_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_all.log; tail -5 gpurun_out/r2a_pytest_all.log
W=camf_ci_f64_100Kx10Kx32c_10M
for sh in 0 1 2; do
  timeout 300 python bench.py --workload $W --mode fast --tuning shape=$sh --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2a_fast10M_shape$sh.json 2> gpurun_out/r2a_fast10M_shape$sh.log
  timeout 300 python bench.py --workload ${W}_zipf1.0 --mode fast --tuning shape=$sh --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2a_fast10Mz_shape$sh.json 2> gpurun_out/r2a_fast10Mz_shape$sh.log
done
timeout 600 python bench.py --workload ${W}_zipf1.0 --mode exact --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_exact10Mz.json 2> gpurun_out/r2a_exact10Mz.log
timeout 300 python bench.py --workload $W --mode exact --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_exact10M.json 2> gpurun_out/r2a_exact10M.log
CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_strict.so timeout 300 python bench.py --workload $W --mode exact --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2a_exact10M_strict.json 2> gpurun_out/r2a_exact10M_strict.log
for f in gpurun_out/r2a_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d.get("parity") and d["parity"].get("ok"), d["e2e"]["value"], d["e2e_pageable"]["value"])
except Exception as e:
    print("ERR", e)
PY
done
