mkdir -p gpurun_out
for v in 5 6; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=$v 100M /"
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=$v 10M /"
done
CARS_WF_VARIANT=3 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=3 10M /"
