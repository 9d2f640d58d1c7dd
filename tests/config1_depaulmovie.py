"""BASELINE.json configs[0]: BiasedMF, 10 factors, on DePaulMovie through the reference's own pipeline stages as
restated here: DataTransformer (compact -> binary) -> DataDAO.readData -> DataSplitter (cv -k 5, --rand-seed 1)
-> toTraditionalSparseMatrix -> BiasedMF.buildModel (100 iterations, learn.rate 2e-2 -bold-driver) -> evalRatings.

    python -m tests.config1_depaulmovie <path to Movie_DePaulMovie/ratings.txt> [--gpu]

Lives under tests/ because it checks against the CPU oracle (test infrastructure).

Plumbing check (the config is "CPU, no GPU" in BASELINE.json): the CPU oracle trains every fold; with --gpu the
engine trains the same folds from the same initial arrays and must return bit-identical MAE / RMSE.  The
reference's own initial factors are wall-clock seeded, so its RMSE is comparable only statistically."""
import math
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carskit_b200 import capi, data  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

if __name__ != "__main__":
    raise SystemExit("run as: python -m tests.config1_depaulmovie <ratings.txt> [--gpu]")
path = sys.argv[1]
use_gpu = "--gpu" in sys.argv
jdk = 8
text = data.transform_to_binary(path, jdk=jdk)
with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
    f.write(text)
ts, dao = data.read_binary_csv(f.name)
os.unlink(f.name)
print(f"users {dao.numUsers()} items {dao.numItems()} pairs {dao.numUserItems()} contexts {dao.numContexts()} "
      f"conditions {dao.numConditions()} dims {dao.numContextDims()} ratings {ts.nnz} globalMean {ts.global_mean:.6f}")
K, SEED, F, ITERS = 5, 1, 10, 100
regs = dict(reg_u=capi.f32(1e-4), reg_i=capi.f32(1e-4), reg_b=capi.f32(1e-4), reg_c=capi.f32(1e-3))
sp = data.DataSplitter(ts, K, SEED)
orc.build()
avg = {"MAE": 0.0, "RMSE": 0.0}
for k in range(1, K + 1):
    train3, test = sp.getKthFold(k)
    train = data.to_traditional(train3)
    g = orc.JavaRandom(100 + k)
    init = {n: g.gaussian(s) for n, s in capi.member_shapes(capi.BIASEDMF, train.num_users, train.num_items, 0, F).items()}
    ref = {n: v.copy() for n, v in init.items()}
    desc = capi.make_desc(train, capi.BIASEDMF, F, **regs)
    n_it, losses = orc.build_model(desc, ref, orc.new_state(capi.f32(2e-2), bold_driver=True), ITERS)
    sa, ss, cnt = orc.eval_ratings(desc, ref, test["u"], test["j"], None, test["r"], 1.0, 5.0)
    mae, rmse = sa / cnt, math.sqrt(ss / cnt)
    line = f"fold {k}: train {train.nnz} (u,i) cells from {train3.nnz} ratings, test {cnt}; {n_it} iterations; oracle MAE {mae:.6f} RMSE {rmse:.6f}"
    if use_gpu:
        from carskit_b200 import recommender
        rec = recommender.BiasedMF(train, {**test, "ctx": None}, fold=k,
                                   conf={"num.factors": str(F), "num.max.iter": str(ITERS), "learn.rate": "2e-2 -max -1 -bold-driver",
                                         "reg.lambda": "0.0001 -c 0.001"})
        m = rec.execute(init={n: v.copy() for n, v in init.items()})
        same = m["MAE"] == mae and m["RMSE"] == rmse and all(np.array_equal(ref[n], rec.model[n]) for n in ref)
        line += f"; B200 MAE {m['MAE']:.6f} RMSE {m['RMSE']:.6f} bit-identical={same}"
        assert same
    print(line)
    avg["MAE"] += mae / K
    avg["RMSE"] += rmse / K
print(f"average over {K} folds: MAE {avg['MAE']:.6f} RMSE {avg['RMSE']:.6f}")
