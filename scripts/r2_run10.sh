#!/bin/bash
# round 2, GPU call 10 (8 GPUs): the scaling bench exactly as the driver launches it, N = 8 (and the reference arm)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2j_ngpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.log; echo "rc=$?"; grep "bench\] rank 0" gpurun_out/r2j_bench_n8.log | tail -5 | cut -c1-250
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "single_process or c_client" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2j_bench_n8.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("rmse_vs_serial"))
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/r2j_bench_n8.log | cut -c1-300
