mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.log; cat gpurun_out/bench_n2.json
for v in 3 4; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fl2_v$v.json 2> gpurun_out/bench_fl2_v$v.log; tail -3 gpurun_out/bench_fl2_v$v.log | head -2
done
