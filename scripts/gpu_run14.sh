mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_recommender_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; grep "epochs in" gpurun_out/bench_n$N.log | head -3; cat gpurun_out/bench_n$N.json
