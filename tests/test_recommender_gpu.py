"""The host-side mirror of the reference classes (carskit_b200/recommender.py) on the GPU, against the
oracle's buildModel(): same hooks in the same order -- initModel -> buildModel (epoch + isConverged +
updateLRate per iteration) -> evalRatings -- so these read like tests of the Java classes."""
import math

import numpy as np
import pytest

from carskit_b200 import capi, recommender, synth
from tests.golden.make_golden import REGS, init_arrays

pytestmark = pytest.mark.gpu

CONF = {"num.max.iter": "6", "learn.rate": "2e-2 -max -1 -bold-driver", "reg.lambda": "0.0001 -c 0.001"}


@pytest.mark.parametrize("name,F,dims", [("pmf", 10, None), ("biasedmf", 10, None), ("camf_c", 10, [7, 7, 2, 3]),
                                          ("camf_ci", 64, [8, 8, 8, 8]), ("camf_cu", 128, [16, 16]),
                                          ("camf_cuci", 32, [5, 4, 3])])
def test_execute_matches_oracle_build_model(oracle, cars_lib, name, F, dims):
    model = capi.MODEL_NAMES[name]
    nnz = 3000 if name == "camf_c" else 30000
    ts, test = synth.make_training_set(400, 150, dims, nnz, seed=17, holdout=0.1)
    init = init_arrays(oracle, model, ts, F, seed=23)
    rec = recommender.getRecommender(name)(ts, test, conf={**CONF, "num.factors": str(F)})
    assert (rec.regU, rec.regC, rec.lRate) == (capi.f32(1e-4), capi.f32(1e-3), capi.f32(2e-2))  # float widening
    measures = rec.execute(init={k: v.copy() for k, v in init.items()})

    ref = {k: v.copy() for k, v in init.items()}
    desc = capi.make_desc(ts, model, F, **REGS)
    state = oracle.new_state(capi.f32(2e-2), bold_driver=True)
    n, losses = oracle.build_model(desc, ref, state, 6)
    assert n == len(rec.iter_losses) == 6
    np.testing.assert_allclose(rec.iter_losses, losses, rtol=1e-11)
    for k in ref:
        assert np.array_equal(ref[k], rec.model[k]), k
    assert rec.lRate == state.lRate  # the bold driver took the same decisions
    sa, ss, cnt = oracle.eval_ratings(desc, ref, test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
    assert measures["RMSE"] == math.sqrt(ss / cnt) and measures["MAE"] == sa / cnt


def test_early_stop_rmse_uses_the_device_resident_model(oracle, cars_lib):
    # evaluation.setup ... --early-stop RMSE: isConverged() calls evalRatings() every iteration
    # (IterativeRecommender.java:156-163); the mirror evaluates on the engine without a download
    ts, test = synth.make_training_set(300, 100, [4, 4], 20000, seed=5, holdout=0.1)
    rec = recommender.CAMF_CI(ts, test, conf={**CONF, "num.factors": "16", "num.max.iter": "30",
                                                "evaluation.setup": "test-set --early-stop RMSE"})
    m = rec.execute(seed=3)
    assert 1 <= len(rec.iter_losses) <= 30 and math.isfinite(m["RMSE"])
    assert rec.measure == m["RMSE"]


def test_nan_loss_raises_like_the_reference_exits(cars_lib):
    ts, _ = synth.make_training_set(50, 20, [2, 2], 1500, seed=3)
    rec = recommender.CAMF_CI(ts, None, conf={**CONF, "num.factors": "8", "learn.rate": "1e6", "num.max.iter": "10"})
    rec.initModel(seed=1)
    with pytest.raises(FloatingPointError):  # IterativeRecommender.java:181-184 calls System.exit(-1)
        rec.buildModel()
    assert rec.engine is None  # the handle is released on the error path


def test_cross_validation_folds_in_parallel_match_folds_one_by_one(oracle, cars_lib):
    # CARSKit.runCrossValidation: K folds from DataSplitter, one recommender (and one engine handle) per thread;
    # the same folds trained one after the other, and by the oracle, must give the same measures bit for bit
    from carskit_b200 import data
    ts, _ = synth.make_training_set(300, 120, [4, 3], 20000, seed=8)
    conf = {**CONF, "num.factors": "16", "num.max.iter": "4"}
    K, seed = 3, 5
    sp = data.DataSplitter(ts, K, seed)
    inits = [init_arrays(oracle, capi.CAMF_CI, sp.getKthFold(k + 1)[0], 16, seed=40 + k) for k in range(K)]
    avg_p, algos_p = recommender.runCrossValidation(ts, "camf_ci", conf, kFold=K, rand_seed=seed, parallel=True,
                                                    inits=[{k: v.copy() for k, v in a.items()} for a in inits])
    avg_s, algos_s = recommender.runCrossValidation(ts, "camf_ci", conf, kFold=K, rand_seed=seed, parallel=False,
                                                    inits=[{k: v.copy() for k, v in a.items()} for a in inits])
    assert avg_p == avg_s and [a.fold for a in algos_p] == [1, 2, 3]
    want = {}
    for k in range(K):
        train, test = sp.getKthFold(k + 1)
        assert algos_p[k].measures == algos_s[k].measures
        ref = {n: v.copy() for n, v in inits[k].items()}
        desc = capi.make_desc(train, capi.CAMF_CI, 16, **REGS)
        oracle.build_model(desc, ref, oracle.new_state(capi.f32(2e-2), bold_driver=True), 4)
        sa, ss, cnt = oracle.eval_ratings(desc, ref, test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
        assert algos_p[k].measures["RMSE"] == math.sqrt(ss / cnt) and algos_p[k].measures["MAE"] == sa / cnt
        for m, v in algos_p[k].measures.items():
            want[m] = want.get(m, 0.0) + v / K
    assert avg_p == want


def test_get_recommender_names():
    assert recommender.getRecommender("CAMF_CI") is recommender.CAMF_CI
    with pytest.raises(ValueError):
        recommender.getRecommender("slim")
