"""Mints tests/golden/sgd_golden.json from the CPU oracle (run: python -m tests.golden.make_golden).

The reference cannot run here (Java, no JVM) and ships no golden outputs, so these vectors pin the
ORACLE (regression) and give the GPU tests a committed, box-independent expectation: per-epoch losses,
a SHA-256 over the trained arrays' bytes and the test RMSE, for every model on small seeded inputs.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from carskit_b200 import capi, synth  # noqa: E402

REGS = dict(reg_u=capi.f32(1e-4), reg_i=capi.f32(1e-4), reg_b=capi.f32(1e-4), reg_c=capi.f32(1e-3))

SPECS = [
    dict(name="pmf_f10", model="pmf", users=60, items=25, dims=None, nnz=1500, F=10, epochs=5, seed=1, order="user_sorted"),
    dict(name="biasedmf_f10", model="biasedmf", users=97, items=79, dims=None, nnz=3000, F=10, epochs=5, seed=2, order="user_sorted"),
    dict(name="camf_c_f10", model="camf_c", users=90, items=120, dims=[7, 7, 2, 3, 2, 9], nnz=4000, F=10, epochs=4, seed=3, order="shuffled"),
    dict(name="camf_ci_f64", model="camf_ci", users=300, items=80, dims=[8, 8, 8, 8], nnz=6000, F=64, epochs=4, seed=4, order="user_sorted"),
    dict(name="camf_ci_f7_shuffled", model="camf_ci", users=50, items=40, dims=[3, 2], nnz=1800, F=7, epochs=6, seed=5, order="shuffled"),
    dict(name="camf_cu_f128", model="camf_cu", users=200, items=60, dims=[16, 16, 16, 16], nnz=5000, F=128, epochs=3, seed=6, order="user_sorted"),
    dict(name="camf_cuci_f32", model="camf_cuci", users=150, items=90, dims=[5, 4, 3], nnz=5000, F=32, epochs=4, seed=7, order="shuffled"),
]


def init_arrays(oracle, model, ts, F, seed):
    g = oracle.JavaRandom(seed)
    shapes = capi.member_shapes(model, ts.num_users, ts.num_items, ts.num_conditions, F)
    # icBias / ucBias ~ U(0,1) (CAMF_CI.java:58-59, CAMF_CU.java:55-56); CAMF_CUCI's tables are Gaussian (:58-64)
    out = {}
    for k, s in shapes.items():
        if k == "cc_sim":  # CAMF_ICS.java:45-48: every similarity starts at 1.0
            out[k] = np.ones(s)
        elif k == "c_mcs":  # CAMF_MCS.java:47-48: cVector_MCS.init(upbound) -> U(0, 1 / sqrt(numContextDims))
            out[k] = g.uniform(s) / np.sqrt(len(ts.empty_conditions))
        elif model in (capi.CAMF_ICS, capi.CAMF_LCS, capi.CAMF_MCS):  # isRankingPred: P.init(), Q.init(), cfMatrix_LCS.init() -> U(0, 1)
            out[k] = g.uniform(s)
        elif k in ("ic_bias", "uc_bias") and model != capi.CAMF_CUCI:
            out[k] = g.uniform(s)
        else:
            out[k] = g.gaussian(s)
    return out


def make_inputs(oracle, spec):
    model = capi.MODEL_NAMES[spec["model"]]
    ts, test = synth.make_training_set(spec["users"], spec["items"], spec["dims"], spec["nnz"], seed=spec["seed"],
                                       order=spec["order"], holdout=0.1)
    desc = capi.make_desc(ts, model, spec["F"], **REGS)
    arrs = init_arrays(oracle, model, ts, spec["F"], seed=spec["seed"] + 100)
    return model, ts, test, desc, arrs


def digest(arrs) -> str:
    h = hashlib.sha256()
    for k in sorted(arrs):
        h.update(k.encode())
        h.update(np.ascontiguousarray(arrs[k]).tobytes())
    return h.hexdigest()


def bold_driver(lr, last_loss, loss, it):
    # IterativeRecommender.updateLRate (:216-229) with -bold-driver, no decay, no max
    if it > 1:
        lr = lr * 1.05 if abs(last_loss) > abs(loss) else lr * 0.5
    return lr


def run_case(oracle, spec):
    model, ts, test, desc, arrs = make_inputs(oracle, spec)
    lr, last = capi.f32(0.02), 0.0
    losses = []
    for it in range(1, spec["epochs"] + 1):
        loss = oracle.epoch(desc, arrs, lr)
        losses.append(loss)
        lr = bold_driver(lr, last, loss, it)
        last = loss
    sa, ss, cnt = oracle.eval_ratings(desc, arrs, test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
    rmse = math.sqrt(ss / cnt)
    return {"losses_hex": [float(x).hex() for x in losses], "digest": digest(arrs), "rmse_hex": float(rmse).hex(),
            "rmse": rmse, "nnz": ts.nnz}


def main():
    from oracle import oracle_py as oracle
    oracle.build()
    cases = []
    for spec in SPECS:
        out = run_case(oracle, spec)
        cases.append({"name": spec["name"], "spec": spec, **out})
        print(spec["name"], out["losses_hex"][-1], out["rmse"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sgd_golden.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "oracle": oracle.lib().oracle_version().decode(),
                   "cases": cases}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
