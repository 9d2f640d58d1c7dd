#!/bin/bash
# round 2, GPU call 7: line-atomicity litmus; FAST edge-case failure; eval-kernel launch list + ncu summaries (reports are
# summarised ON the box: gpurun_out/ merges at most 64 MiB back)
mkdir -p gpurun_out
timeout 600 ./scripts/litmus/line_atomicity > gpurun_out/r2g_litmus.txt 2>&1; echo "rc=$?" >> gpurun_out/r2g_litmus.txt; cat gpurun_out/r2g_litmus.txt
timeout 600 python -m pytest tests/test_fast_gpu.py -q -x --tb=short -k edge > gpurun_out/r2g_pytest_edge.log 2>&1; tail -30 gpurun_out/r2g_pytest_edge.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_eval.csv python scripts/bench_eval_kernels.py 10000000 4000 > gpurun_out/r2g_ncu_eval.log 2>&1
prof() { # tag kernel-regex args...
  tag=$1; re=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -o gpurun_out/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep 25 > gpurun_out/ncu_summary_$tag.txt 2>&1
  rm -f gpurun_out/prof_$tag.ncu-rep
}
prof r2g_predict predict_group python scripts/bench_eval_kernels.py 10000000 64
prof r2g_rank_score rank_score python scripts/bench_eval_kernels.py 1000 4000
prof r2g_rank_select rank_select python scripts/bench_eval_kernels.py 1000 4000
CARSKIT_B200_LIB=$PWD/carskit_b200/libcarskit_b200_strict.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2g_exact100M_strict.json 2> gpurun_out/r2g_exact100M_strict.log
timeout 600 python bench.py --workload camf_c_f10_frappe_shaped --mode fast --steps 20 --warmup 3 > gpurun_out/r2g_config2_fast.json 2> gpurun_out/r2g_config2_fast.log; tail -3 gpurun_out/r2g_config2_fast.log
for f in gpurun_out/r2g_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity"))
except Exception as e:
    print("ERR", e)
PY
done
du -sh gpurun_out
