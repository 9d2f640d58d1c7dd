#!/bin/bash
# round-1 closing run on one GPU: the whole GPU suite, smoke, the bench line, the reference arm, the launch list,
# and compute-sanitizer over a small subset of the tests that exercise the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r44_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r44_pytest.log
tail -4 gpurun_out/r44_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CARS_SCHED_TRACE=1 timeout 900 python bench.py > gpurun_out/r44_bench_default.json 2> gpurun_out/r44_bench_default.log; grep -v "^$" gpurun_out/r44_bench_default.log | tail -14; cat gpurun_out/r44_bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r44_bench_reference.json 2> gpurun_out/r44_bench_reference.log; cat gpurun_out/r44_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r44_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r44_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fm_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "fm_matches_sparse_oracle or device_built_levels or golden or ragged" > gpurun_out/r44_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r44_memcheck.log; tail -6 gpurun_out/r44_memcheck.log
