mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; lscpu | grep "Model name"; java -version 2>&1 | head -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 2 > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.log; tail -5 gpurun_out/bench_v0.log; cat gpurun_out/bench_v0.json
CARS_WF_VARIANT=1 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.log; tail -3 gpurun_out/bench_v1.log
CARS_WF_VARIANT=2 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.log; tail -3 gpurun_out/bench_v2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_10M.csv python bench.py --steps 2 --warmup 1 --workload camf_ci_f64_100Kx10Kx32c_10M --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:sgd_wavefront -s 1 -c 1 -o gpurun_out/prof_wf_10M python bench.py --steps 1 --warmup 1 --workload camf_ci_f64_100Kx10Kx32c_10M --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_full.log
tail -3 gpurun_out/ncu_full.log
