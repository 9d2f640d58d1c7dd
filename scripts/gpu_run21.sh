mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_fm_gpu.py -m gpu -x -q 2>&1 | tail -8
