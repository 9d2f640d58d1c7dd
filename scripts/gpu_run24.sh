mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload camf_cu_f128_2Mx200Kx64c_200M > gpurun_out/bench_camf_cu_f128_200M.json 2> gpurun_out/bench_camf_cu_f128_200M.log; grep -E "cars_create|epochs in|e2e" gpurun_out/bench_camf_cu_f128_200M.log; cut -c1-300 gpurun_out/bench_camf_cu_f128_200M.json
python - <<'PY'
# BASELINE config 2: CAMF_C F=10 on a Frappe-shaped synthetic (serial kernel; every rating touches condBias)
import time, numpy as np
from carskit_b200 import recommender, synth
ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1)
rec = recommender.CAMF_C(ts, test, conf={"num.factors": "10", "num.max.iter": "10"})
rec.initModel(seed=1)
t0 = time.time(); rec.keep_engine = True; rec.buildModel(); dt = time.time() - t0
print("config2 CAMF_C F=10 nnz", ts.nnz, "10 epochs", round(dt, 3), "s;", round(ts.nnz * len(rec.iter_losses) / dt / 1e6, 3), "M updates/s; kernel ms/epoch", rec.engine.stats().last_epoch_ms, "RMSE", rec.evalRatings()["RMSE"])
rec.close_engine()
PY
