mkdir -p gpurun_out
for la in 0 1; do
CARS_LOOK_AHEAD=$la timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/look_ahead=$la 100M /"
CARS_LOOK_AHEAD=$la timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep "epochs in" | sed "s/^/look_ahead=$la 10M /"
done
export CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_trace.so
CARS_LOOK_AHEAD=1 timeout 600 python scripts/trace_flagged.py camf_ci_f64_1Mx100Kx32c_100M gpurun_out/trace_100M_lookahead.json 2>&1 | grep -E "mean_us|tries_mean|epoch_ms|\": \{"
