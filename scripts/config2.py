"""BASELINE.json configs[1]: CAMF_C, 10 factors, Frappe-shaped synthetic (957 x 4 082, 8 context dimensions with
7/7/2/3/2/9/80/233 conditions, 96 203 ratings, 90/10 split) through recommender.CAMF_C on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carskit_b200 import recommender, synth
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1)
rec = recommender.CAMF_C(ts, test, conf={"num.factors": "10", "num.max.iter": str(epochs)})
rec.initModel(seed=1)
t0 = time.time(); rec.keep_engine = True; rec.buildModel(); dt = time.time() - t0
print("config2 CAMF_C F=10 nnz", ts.nnz, epochs, "epochs", round(dt, 3), "s;",
      round(ts.nnz * len(rec.iter_losses) / dt / 1e6, 3), "M updates/s; kernel ms/epoch", rec.engine.stats().last_epoch_ms,
      "RMSE", rec.evalRatings()["RMSE"])
rec.close_engine()
