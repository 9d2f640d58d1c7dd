"""Host-side mirror of the reference's epoch control and configuration parsing (carskit_b200/recommender.py), on
the CPU: hyper-parameters are Java floats widened to double (IterativeRecommender.java:76-108), and isConverged /
updateLRate (:145-229) take the same decisions as the oracle's restatement when fed the same loss sequence.  No
engine is opened here (the constructor does not touch the GPU)."""
import math
import os

import numpy as np
import pytest

from carskit_b200 import capi, recommender, synth


def make(conf=None, name="camf_ci"):
    ts, _ = synth.make_training_set(20, 10, [2, 2], 200, seed=1)
    return recommender.getRecommender(name)(ts, None, conf=conf or {})


def test_line_configer():
    lc = recommender.LineConfiger("2e-2 -max -1 -bold-driver -decay 0.95")
    assert lc.getMainParam() == "2e-2" and lc.contains("-bold-driver") and not lc.contains("-momentum")
    assert lc.getFloat("-max", 7) == -1.0 and lc.getFloat("-decay", -1) == float(np.float32(0.95))
    assert lc.getFloat("-momentum", 50) == 50.0
    ev = recommender.LineConfiger("cv -k 5 -p on --rand-seed 1 --test-view all --early-stop RMSE")
    assert ev.getMainParam() == "cv" and ev.getString("--early-stop") == "RMSE" and ev.getString("-k") == "5"


def test_hyper_parameters_are_floats_widened_to_double():
    rec = make({"learn.rate": "0.02 -max 0.05 -decay 0.9", "reg.lambda": "0.0001 -u 0.001 -c 0.01", "num.factors": "7"})
    assert rec.initLRate == rec.lRate == capi.f32(0.02) == 0.019999999552965164
    assert (rec.reg, rec.regU, rec.regI, rec.regB, rec.regC) == (capi.f32(1e-4), capi.f32(1e-3), capi.f32(1e-4),
                                                                 capi.f32(1e-4), capi.f32(1e-2))
    assert rec.maxLRate == capi.f32(0.05) and rec.decay == capi.f32(0.9) and not rec.isBoldDriver and rec.numFactors == 7
    d = rec._desc()
    assert (d.reg_u, d.reg_c, d.num_factors, d.model) == (capi.f32(1e-3), capi.f32(1e-2), 7, capi.CAMF_CI)
    dflt = make()  # setting.conf:52-59
    assert dflt.isBoldDriver and dflt.numFactors == 10 and dflt.numIters == 100 and dflt.regC == capi.f32(1e-3)


@pytest.mark.parametrize("conf,early", [({"learn.rate": "2e-2 -max -1 -bold-driver"}, 0),
                                         ({"learn.rate": "2e-2 -max 0.021 -bold-driver"}, 0),
                                         ({"learn.rate": "0.05 -decay 0.9"}, 0),
                                         ({"learn.rate": "2e-2 -bold-driver", "evaluation.setup": "cv --early-stop loss"}, 1)])
def test_epoch_control_follows_the_oracle(oracle, conf, early):
    # a loss sequence with rises, falls, a plateau below the 1e-5 delta and a tiny value
    losses = [100.0, 80.0, 90.0, 70.0, 69.0, 69.0 - 4e-6, 50.0, 3e-6, 1.0]
    rec = make(conf)
    st = oracle.new_state(rec.lRate, bold_driver=rec.isBoldDriver, decay=rec.decay, max_lrate=rec.maxLRate, early_stop=early)
    for it, loss in enumerate(losses, start=1):
        rec.loss = loss
        st.loss = loss
        got = rec.isConverged(it)
        want = oracle.lib().oracle_is_converged(st, it)
        assert int(got) == want, (it, loss)
        assert rec.lRate == st.lRate and rec.last_loss == st.last_loss
        if got:
            break
    else:
        pytest.fail("the sequence should have converged")


def test_nan_loss_raises():
    rec = make()
    rec.loss = float("nan")
    with pytest.raises(FloatingPointError):  # IterativeRecommender.java:181-184 exits the JVM
        rec.isConverged(1)
    rec.loss = math.inf
    with pytest.raises(FloatingPointError):
        rec.isConverged(2)


def test_java_hashset_order_known_answer():
    # Integer keys: the spread is h ^ (h >>> 16), so small ids iterate in ascending order until the table wraps
    assert recommender.java_hashset_order([5, 3, 9, 3, 1]).tolist() == [1, 3, 5, 9]
    ids = [17, 1, 33, 16, 0]   # capacity 16: 17, 1 and 33 share bucket 1 in insertion order; 16 and 0 share bucket 0
    assert recommender.java_hashset_order(ids).tolist() == [16, 0, 17, 1, 33]
    big = list(range(70000, 70013))  # 13 > 12 entries: capacity 32; 70000 = 0x11170 -> (h ^ h >> 16) & 31
    out = recommender.java_hashset_order(big).tolist()
    assert sorted(out) == big and out == sorted(big, key=lambda v: ((v ^ (v >> 16)) & 31, big.index(v)))


def test_recommender_registry_and_model_members():
    """Every recommender name CARSKit.getRecommender would dispatch for this path maps to a class whose MODEL, member list and
    member shapes agree with the C ABI's enum (include/carskit_b200.h) -- incl. the one-chain models (ICS / LCS / MCS / SVD++)."""
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "carskit_b200.h")).read()
    enum = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(CARS_[A-Z_]+?)\s*=\s*(\d+)\s*[,/ ]", hdr.split("enum cars_model")[1].split("};")[0])}
    want = {"pmf": "CARS_PMF", "biasedmf": "CARS_BIASEDMF", "svd++": "CARS_SVDPP", "camf_c": "CARS_CAMF_C", "camf_ci": "CARS_CAMF_CI",
            "camf_cu": "CARS_CAMF_CU", "camf_cuci": "CARS_CAMF_CUCI", "camf_ics": "CARS_CAMF_ICS", "camf_lcs": "CARS_CAMF_LCS",
            "camf_mcs": "CARS_CAMF_MCS", "fm": "CARS_FM"}
    for name, sym in want.items():
        cls = recommender.getRecommender(name)
        assert cls.MODEL == enum[sym], (name, cls.MODEL, enum[sym])
    shapes = capi.member_shapes(capi.CAMF_LCS, 7, 5, 9, 4, num_context_factors=3)
    assert shapes == {"P": (7, 4), "Q": (5, 4), "cf_lcs": (9, 3)}
    assert capi.member_shapes(capi.CAMF_MCS, 7, 5, 9, 4) == {"P": (7, 4), "Q": (5, 4), "c_mcs": (9,)}
    assert capi.member_shapes(capi.SVDPP, 7, 5, 0, 4) == {"P": (7, 4), "Q": (5, 4), "user_bias": (7,), "item_bias": (5,), "Y": (5, 4)}
    # the ctypes mirrors have one field per header member, in the header's order
    fields = [n for n, _ in capi.CarsModelArrays._fields_]
    hdr_fields = re.findall(r"double\*\s+(\w+);", hdr.split("typedef struct cars_model_arrays {")[1].split("} cars_model_arrays;")[0])
    assert fields == hdr_fields
