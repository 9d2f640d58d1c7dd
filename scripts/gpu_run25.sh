mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "camf_c or golden or serial or execute" 2>&1 | tail -3
python - <<'PY'
import time, numpy as np
from carskit_b200 import recommender, synth
ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1)
for it in range(2):
    rec = recommender.CAMF_C(ts, test, conf={"num.factors": "10", "num.max.iter": "10"})
    rec.initModel(seed=1)
    t0 = time.time(); rec.keep_engine = True; rec.buildModel(); dt = time.time() - t0
    print("config2 CAMF_C F=10 nnz", ts.nnz, "10 epochs", round(dt, 3), "s;", round(ts.nnz * len(rec.iter_losses) / dt / 1e6, 3), "M updates/s e2e; kernel ms/epoch", round(rec.engine.stats().last_epoch_ms, 2), "RMSE", rec.evalRatings()["RMSE"])
    rec.close_engine()
PY
