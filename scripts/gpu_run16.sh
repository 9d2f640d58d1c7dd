mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 3 5; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=$v 100M /"
done
CARS_WF_VARIANT=3 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=3 10M /"
