mkdir -p gpurun_out
CARS_SCHEDULE=flagged timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flag or level or orders or edge" 2>&1 | tail -8
for v in 0 1 2 3; do
CARS_SCHEDULE=flagged CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_flr_v$v.json 2> gpurun_out/bench_flr_v$v.log; tail -3 gpurun_out/bench_flr_v$v.log | head -2
done
