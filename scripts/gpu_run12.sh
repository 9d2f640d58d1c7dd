mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for v in 3 2; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fl3_v$v.json 2> gpurun_out/bench_fl3_v$v.log; tail -3 gpurun_out/bench_fl3_v$v.log | head -2
done
CARS_WF_VARIANT=3 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep epochs
