#!/bin/bash
# round 2, GPU call 8: K1t (tagged rows) parity + speed; FAST after the damping change; config 2 FAST
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "tagged" --tb=short > gpurun_out/r2h_pytest_tagged.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest_tagged.log; tail -12 gpurun_out/r2h_pytest_tagged.log
timeout 600 python -m pytest tests/test_fast_gpu.py -q -x --tb=short > gpurun_out/r2h_pytest_fast.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest_fast.log; tail -4 gpurun_out/r2h_pytest_fast.log
W=camf_ci_f64_100Kx10Kx32c_10M
for c in 2 3; do
timeout 300 python bench.py --workload $W --tuning "tagged=1;tagged_ctas=$c" --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_tagged10M_c$c.json 2> gpurun_out/r2h_tagged10M_c$c.log; tail -2 gpurun_out/r2h_tagged10M_c$c.log | cut -c1-300
timeout 600 python bench.py --tuning "tagged=1;tagged_ctas=$c" --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2h_tagged100M_c$c.json 2> gpurun_out/r2h_tagged100M_c$c.log; tail -2 gpurun_out/r2h_tagged100M_c$c.log | cut -c1-300
done
timeout 600 python bench.py --workload camf_c_f10_frappe_shaped --mode fast --steps 20 --warmup 3 > gpurun_out/r2h_config2_fast.json 2> gpurun_out/r2h_config2_fast.log; tail -3 gpurun_out/r2h_config2_fast.log | cut -c1-300
for f in gpurun_out/r2h_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity") and d["parity"].get("ok"))
except Exception as e:
    print("ERR", e)
PY
done
