#!/bin/bash
# FM row order + wide coord sums, record gather, pipelined CAMF_C serial kernel: tests + FM bench + config 2 + default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r29_pytest.log
tail -15 gpurun_out/r29_pytest.log
timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 > gpurun_out/r29_bench_fm.json 2> gpurun_out/r29_bench_fm.log; cut -c1-400 gpurun_out/r29_bench_fm.json
CARS_FM_BLOCK_ROWS=0 timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r29_bench_fm_noblock.json 2> gpurun_out/r29_bench_fm_noblock.log; cut -c1-300 gpurun_out/r29_bench_fm_noblock.json
for pl in 1 0; do
CARS_SERIAL_PIPELINE=$pl python - <<'PY'
# BASELINE config 2: CAMF_C F=10 on a Frappe-shaped synthetic (serial kernel; every rating touches condBias)
import os, time, numpy as np
from carskit_b200 import recommender, synth
ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1)
rec = recommender.CAMF_C(ts, test, conf={"num.factors": "10", "num.max.iter": "10"})
rec.initModel(seed=1)
t0 = time.time(); rec.keep_engine = True; rec.buildModel(); dt = time.time() - t0
print("config2 CAMF_C F=10 pipelined=%s nnz" % os.environ["CARS_SERIAL_PIPELINE"], ts.nnz, "10 epochs", round(dt, 3), "s;", round(ts.nnz * len(rec.iter_losses) / dt / 1e6, 3), "M updates/s; kernel ms/epoch", rec.engine.stats().last_epoch_ms, "RMSE", rec.evalRatings()["RMSE"])
rec.close_engine()
PY
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r29_bench_default.json 2> gpurun_out/r29_bench_default.log; tail -2 gpurun_out/r29_bench_default.log; grep -o '"e2e": {[^}]*}' gpurun_out/r29_bench_default.json
