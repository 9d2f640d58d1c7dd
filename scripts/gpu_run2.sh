mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in 0 1 2; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_df_v$v.json 2> gpurun_out/bench_df_v$v.log; tail -3 gpurun_out/bench_df_v$v.log
done
