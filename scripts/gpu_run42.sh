#!/bin/bash
# BASELINE config 4 at full size: FM k=64, 500 M rows over 4 GPUs
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 1 --workload fm_k64_5Mx500Kx32c_125M_per_gpu --no-cpu-baseline > gpurun_out/r42_bench_fm_cfg4_n4.json 2> gpurun_out/r42_bench_fm_cfg4_n4.log; cat gpurun_out/r42_bench_fm_cfg4_n4.json
