mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in 2 0; do
CARS_SCHEDULE=flagged CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_flp_v$v.json 2> gpurun_out/bench_flp_v$v.log; tail -3 gpurun_out/bench_flp_v$v.log
done
