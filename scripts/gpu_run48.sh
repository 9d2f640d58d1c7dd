#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recommender_gpu.py -m gpu -x -q > gpurun_out/r48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r48_pytest.log
tail -15 gpurun_out/r48_pytest.log
