#!/bin/bash
# final state: whole GPU suite + one ncu --set full capture of the dominant kernel (for roofline.traffic)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r47_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r47_pytest.log
tail -4 gpurun_out/r47_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgd_flagged -s 1 -c 1 -o gpurun_out/prof_r47_sgd_flagged python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r47_ncu.log 2>&1
tail -2 gpurun_out/r47_ncu.log
