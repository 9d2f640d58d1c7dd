mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 3 --warmup 1 --workload fm_k64_250Kx25Kx32c_25M > gpurun_out/bench_fm_n$N.json 2> gpurun_out/bench_fm_n$N.log; tail -3 gpurun_out/bench_fm_n$N.log; cat gpurun_out/bench_fm_n$N.json
