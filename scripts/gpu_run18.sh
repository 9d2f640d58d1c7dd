mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
CARS_DEFER_RELEASE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flag or level or orders or edge or golden" 2>&1 | tail -2
for d in 0 1; do
CARS_DEFER_RELEASE=$d timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/defer=$d 100M /"
CARS_DEFER_RELEASE=$d timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --workload camf_ci_f64_100Kx10Kx32c_10M 2>&1 >/dev/null | grep "epochs in" | sed "s/^/defer=$d 10M /"
done
export CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_trace.so
CARS_DEFER_RELEASE=1 timeout 600 python scripts/trace_flagged.py camf_ci_f64_1Mx100Kx32c_100M gpurun_out/trace_100M_defer.json 2>&1 | grep -A4 -E "poll_turn|gather|compute_scatter|release\"|tries_mean|epoch_ms" | grep -E "mean_us|tries_mean|epoch_ms|\": \{"
