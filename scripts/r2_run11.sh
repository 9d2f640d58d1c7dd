#!/bin/bash
# round 2, GPU call 11: full GPU suite after CAMF_ICS / tagged / multi-GPU / ingest changes
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest_all.log; tail -8 gpurun_out/r2k_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
timeout 600 python scripts/bench_eval_kernels.py > gpurun_out/r2k_eval_kernels.jsonl 2> gpurun_out/r2k_eval_kernels.log; cat gpurun_out/r2k_eval_kernels.jsonl
