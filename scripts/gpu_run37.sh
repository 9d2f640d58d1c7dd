#!/bin/bash
# 2 GPUs: the multi-GPU tests (skipped on one GPU), the sharded SGD bench and the row-sharded FM bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "multi or shard" > gpurun_out/r37_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r37_pytest.log
tail -4 gpurun_out/r37_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r37_bench_n2.json 2> gpurun_out/r37_bench_n2.log; grep "bench\]" gpurun_out/r37_bench_n2.log | tail -6; cat gpurun_out/r37_bench_n2.json | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --workload fm_k64_250Kx25Kx32c_25M --no-cpu-baseline > gpurun_out/r37_bench_fm_n2.json 2> gpurun_out/r37_bench_fm_n2.log; cat gpurun_out/r37_bench_fm_n2.json | cut -c1-500
