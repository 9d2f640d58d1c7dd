#!/bin/bash
# round 2, GPU call 2: FAST with hot-row accumulation; 100 M-rating lines (EXACT with parity, FAST uniform, FAST Zipf)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -s > gpurun_out/r2b_pytest_fast.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_fast.log; tail -5 gpurun_out/r2b_pytest_fast.log
W=camf_ci_f64_100Kx10Kx32c_10M
for fl in 1 4 16 64; do
  timeout 300 python bench.py --workload ${W}_zipf1.0 --mode fast --tuning fast_hot_flush=$fl --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2b_fast10Mz_flush$fl.json 2> gpurun_out/r2b_fast10Mz_flush$fl.log
done
timeout 300 python bench.py --workload ${W}_zipf1.0 --mode fast --tuning fast_hot_rows=0 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2b_fast10Mz_nohot.json 2> gpurun_out/r2b_fast10Mz_nohot.log
timeout 300 python bench.py --workload ${W} --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_fast10M.json 2> gpurun_out/r2b_fast10M.log
B=camf_ci_f64_1Mx100Kx32c_100M
timeout 900 python bench.py --workload ${B}_zipf1.0 --mode fast --steps 10 --warmup 3 > gpurun_out/r2b_fast100Mz.json 2> gpurun_out/r2b_fast100Mz.log
timeout 900 python bench.py --workload ${B} --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_fast100M.json 2> gpurun_out/r2b_fast100M.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_exact100M.json 2> gpurun_out/r2b_exact100M.log
for f in gpurun_out/r2b_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d.get("parity"), d["e2e"]["value"], d["e2e_pageable"]["value"], d["config"].get("fast_min_item_scale"))
except Exception as e:
    print("ERR", e)
PY
done
