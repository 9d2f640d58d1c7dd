mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 2 0; do
CARS_SCHEDULE=flagged CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fl2_v$v.json 2> gpurun_out/bench_fl2_v$v.log; tail -3 gpurun_out/bench_fl2_v$v.log | head -2
done
for s in flagged wavefront dataflow; do
CARS_SCHEDULE=$s CARS_WF_VARIANT=2 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep epochs
done
