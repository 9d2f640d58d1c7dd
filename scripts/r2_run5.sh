#!/bin/bash
# round 2, GPU call 5 (2 GPUs): multi-GPU tests (torchrun route + single-process handle + c_client --gpus 2), N=2 bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_smi.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_fast_gpu.py tests/test_jni.py -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log; tail -15 gpurun_out/r2e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.log; grep "bench\]" gpurun_out/r2e_bench_n2.log | tail -8 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --mode fast > gpurun_out/r2e_bench_n2_fast.json 2> gpurun_out/r2e_bench_n2_fast.log
for f in gpurun_out/r2e_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("rmse_vs_serial"))
except Exception as e:
    print("ERR", e)
PY
done
