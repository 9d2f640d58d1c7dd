"""Mints tests/golden/fm_golden.json from the LITERAL dense FM oracle (python -m tests.golden.make_fm_golden).
Values are stored as hex floats: the CPU test reproduces them bit for bit, the GPU test within the stated
tolerances (tree-ordered sums, closed-form denominators)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SPEC = dict(users=30, items=20, dims=[2, 3], nnz=600, k=4, seed=41, iters=4, holdout=0.15)


def run(oracle):
    from tests.test_fm_oracle import clone, fm_inputs
    ts, test, prob, arrs = fm_inputs(oracle, SPEC["users"], SPEC["items"], SPEC["dims"], SPEC["nnz"], SPEC["k"],
                                     seed=SPEC["seed"], holdout=SPEC["holdout"])
    m = clone(arrs)
    oracle.fm_dense_build(prob, m, SPEC["iters"])
    pred = oracle.fm_predict(prob, m, test["u"], test["j"], test["ctx"], bound=True, lo=1.0, hi=5.0)
    return {"w0": float(m["w0"][0]).hex(), "w": [float(x).hex() for x in m["w"]],
            "V": [float(x).hex() for x in m["V"].reshape(-1)], "pred": [float(x).hex() for x in pred]}


if __name__ == "__main__":
    from oracle import oracle_py as oracle
    oracle.build()
    out = {"generator": "tests/golden/make_fm_golden.py", "spec": SPEC, **run(oracle)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fm_golden.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, len(out["w"]), len(out["V"]), len(out["pred"]))
