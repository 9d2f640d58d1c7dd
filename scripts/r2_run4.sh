#!/bin/bash
# round 2, GPU call 4: FAST after record prefetch x2 / icBias accumulators / grid shrink; JNI glue under the fake JNIEnv
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fast_gpu.py tests/test_jni.py -q -s > gpurun_out/r2d_pytest_fast.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_fast.log; grep -E "RMSE|passed|failed|rc=" gpurun_out/r2d_pytest_fast.log
W=camf_ci_f64_100Kx10Kx32c_10M
B=camf_ci_f64_1Mx100Kx32c_100M
timeout 300 python bench.py --workload ${W} --mode fast --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_fast10M.json 2> gpurun_out/r2d_fast10M.log
timeout 300 python bench.py --workload ${W}_zipf1.0 --mode fast --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_fast10Mz.json 2> gpurun_out/r2d_fast10Mz.log
timeout 900 python bench.py --workload ${B} --mode fast --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_fast100M.json 2> gpurun_out/r2d_fast100M.log
timeout 900 python bench.py --workload ${B}_zipf1.0 --mode fast --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_fast100Mz.json 2> gpurun_out/r2d_fast100Mz.log
for f in gpurun_out/r2d_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e"]["seconds"], d["e2e_pageable"]["value"], d["config"].get("fast_min_item_scale"))
except Exception as e:
    print("ERR", e)
PY
done
