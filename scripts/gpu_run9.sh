mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.log; tail -4 gpurun_out/bench_r1_n1.log; cat gpurun_out/bench_r1_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.log; cat gpurun_out/bench_r1_ref.json
