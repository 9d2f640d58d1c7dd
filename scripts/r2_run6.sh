#!/bin/bash
# round 2, GPU call 6: rank / predict kernels (tests + launch list + ncu), ingest tests, strict-acquire A/B at 100 M
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rank_gpu.py tests/test_recommender_gpu.py tests/test_gpu_parity.py tests/test_fast_gpu.py tests/test_ingest.py -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log; tail -5 gpurun_out/r2f_pytest.log
timeout 600 python scripts/bench_eval_kernels.py > gpurun_out/r2f_eval_kernels.jsonl 2> gpurun_out/r2f_eval_kernels.log; cat gpurun_out/r2f_eval_kernels.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_eval.csv python scripts/bench_eval_kernels.py 10000000 4000 > gpurun_out/r2f_ncu_eval.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_group -s 1 -c 1 -o gpurun_out/prof_r2f_predict python scripts/bench_eval_kernels.py 10000000 64 > gpurun_out/r2f_ncu_predict.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_score -s 1 -c 1 -o gpurun_out/prof_r2f_rank_score python scripts/bench_eval_kernels.py 1000 4000 > gpurun_out/r2f_ncu_rank_score.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_select -s 1 -c 1 -o gpurun_out/prof_r2f_rank_select python scripts/bench_eval_kernels.py 1000 4000 > gpurun_out/r2f_ncu_rank_select.log 2>&1
CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_strict.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2f_exact100M_strict.json 2> gpurun_out/r2f_exact100M_strict.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2f_exact100M_default.json 2> gpurun_out/r2f_exact100M_default.log
timeout 600 python bench.py --workload camf_c_f10_frappe_shaped --mode fast --steps 20 --warmup 3 > gpurun_out/r2f_config2_fast.json 2> gpurun_out/r2f_config2_fast.log
timeout 600 python bench.py --workload camf_c_f10_frappe_shaped --mode exact --steps 5 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_config2_exact.json 2> gpurun_out/r2f_config2_exact.log
for f in gpurun_out/r2f_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity"))
except Exception as e:
    print("ERR", e)
PY
done
