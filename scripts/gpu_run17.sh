mkdir -p gpurun_out
export CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_trace.so
timeout 600 python scripts/trace_flagged.py camf_ci_f64_100Kx10Kx32c_10M gpurun_out/trace_10M.json 2>&1 | tail -40
timeout 600 python scripts/trace_flagged.py camf_ci_f64_1Mx100Kx32c_100M gpurun_out/trace_100M.json 2>&1 | tail -40
