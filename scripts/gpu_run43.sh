#!/bin/bash
# BASELINE config 5 at full size: CAMF_CU F=128, 10 M users x 1 M items x 64 conditions, 1 B ratings over 8 GPUs
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 3 --warmup 3 --workload camf_cu_f128_1250Kx1Mx64c_125M_per_gpu > gpurun_out/r43_bench_cfg5_n8.json 2> gpurun_out/r43_bench_cfg5_n8.log; grep "bench\]" gpurun_out/r43_bench_cfg5_n8.log | grep "rank 0" | cut -c1-250; cat gpurun_out/r43_bench_cfg5_n8.json
