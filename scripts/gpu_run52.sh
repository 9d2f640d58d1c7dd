#!/bin/bash
mkdir -p gpurun_out
CARS_WF_VARIANT=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact_mode_bit_identical or flag_schedules_large or golden or orders_and_skew" > gpurun_out/r52_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r52_pytest.log
tail -3 gpurun_out/r52_pytest.log
for v in 8 7 8; do
CARS_WF_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/variant=$v /"
done
