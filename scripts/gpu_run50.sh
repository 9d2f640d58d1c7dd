#!/bin/bash
mkdir -p gpurun_out
CARS_WF_VARIANT=7 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact_mode_bit_identical or flag_schedules_large or golden or orders_and_skew" > gpurun_out/r50_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r50_pytest.log
tail -4 gpurun_out/r50_pytest.log
for v in 7 3 7 3; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=$v /"
done
