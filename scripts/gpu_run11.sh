mkdir -p gpurun_out
timeout 600 python bench.py --workload fm_k16_tiny --steps 3 --warmup 1 2>&1 | tail -2
timeout 900 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 > gpurun_out/bench_fm_25M.json 2> gpurun_out/bench_fm_25M.log; tail -3 gpurun_out/bench_fm_25M.log; cat gpurun_out/bench_fm_25M.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_launches.log
bash scripts/gpu_prof.sh r1_default sgd_flagged
