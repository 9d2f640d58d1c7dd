#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r53_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r53_pytest.log
tail -3 gpurun_out/r53_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r53_bench_default.json 2> gpurun_out/r53_bench_default.log; grep "bench\]" gpurun_out/r53_bench_default.log | tail -2; cat gpurun_out/r53_bench_default.json
