#!/bin/bash
# round 2, GPU call 9: K1t with the cheap retry path: parity, speed, ncu summary
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "tagged" --tb=short > gpurun_out/r2i_pytest_tagged.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest_tagged.log; tail -4 gpurun_out/r2i_pytest_tagged.log
W=camf_ci_f64_100Kx10Kx32c_10M
for c in 2 3; do
timeout 300 python bench.py --workload $W --tuning "tagged=1;tagged_ctas=$c" --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2i_tagged10M_c$c.json 2> gpurun_out/r2i_tagged10M_c$c.log
timeout 600 python bench.py --tuning "tagged=1;tagged_ctas=$c" --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2i_tagged100M_c$c.json 2> gpurun_out/r2i_tagged100M_c$c.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_tagged -s 1 -c 1 -o gpurun_out/prof_r2i_tagged python bench.py --tuning "tagged=1;tagged_ctas=2" --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_r2i_tagged.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r2i_tagged.ncu-rep 30 > gpurun_out/ncu_summary_r2i_tagged.txt 2>&1; rm -f gpurun_out/prof_r2i_tagged.ncu-rep
for f in gpurun_out/r2i_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"])
except Exception as e:
    print("ERR", e)
PY
done
head -40 gpurun_out/ncu_summary_r2i_tagged.txt
