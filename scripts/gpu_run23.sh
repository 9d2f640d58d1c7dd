mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rank_gpu.py -m gpu -x -q 2>&1 | tail -15
