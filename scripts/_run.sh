#!/bin/bash
mkdir -p gpurun_out
B=camf_ci_f64_1Mx100Kx32c_100M
timeout 900 python -m pytest tests/test_fast_gpu.py tests/test_rank_gpu.py tests/test_recommender_gpu.py -q -x --tb=short > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log; tail -4 gpurun_out/r2m_pytest.log
timeout 600 python scripts/bench_eval_kernels.py > gpurun_out/r2m_eval_kernels.jsonl 2> gpurun_out/r2m_eval_kernels.log; cat gpurun_out/r2m_eval_kernels.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_eval.csv python scripts/bench_eval_kernels.py 10000000 4000 > gpurun_out/r2m_ncu_eval.log 2>&1
python scripts/launch_summary.py gpurun_out/r2m_launches_eval.csv | head -8
prof() { tag=$1; re=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -o gpurun_out/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep 25 > gpurun_out/ncu_summary_$tag.txt 2>&1
  rm -f gpurun_out/prof_$tag.ncu-rep; }
prof r2m_rank_score rank_score python scripts/bench_eval_kernels.py 1000 4000
prof r2m_fast_bulk sgd_fast python bench.py --workload $B --mode fast --steps 1 --warmup 1 --no-cpu-baseline --no-parity
timeout 900 python bench.py --workload ${B}_zipf1.0 --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_fast100Mz.json 2> gpurun_out/r2m_fast100Mz.log
timeout 900 python bench.py --workload ${B} --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_fast100M.json 2> gpurun_out/r2m_fast100M.log
timeout 600 python bench.py --workload camf_c_f10_frappe_shaped --mode fast --steps 20 --warmup 3 > gpurun_out/r2m_config2_fast.json 2> gpurun_out/r2m_config2_fast.log
F=fm_k64_5Mx500Kx32c_125M_per_gpu
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2m_launches_fm.csv python bench.py --workload $F --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2m_ncu_fm.log 2>&1
python scripts/launch_summary.py gpurun_out/r2m_launches_fm.csv | head -16
for k in fm_piece_reduce fm_row_update fm_prepare; do prof r2m_$k $k python bench.py --workload $F --steps 1 --warmup 0 --no-cpu-baseline; done
for f in gpurun_out/r2m_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity") and (d["parity"]["ok"], d["parity"]["loss_rel"]))
except Exception as e:
    print("ERR", e)
PY
done
du -sh gpurun_out
