"""Mints tests/golden/jvm_golden.json by EXECUTING THE REFERENCE'S OWN BYTECODE (run: python -m tests.golden.make_jvm_golden).

The image has no JVM, but /root/reference/jar/CARSKit-v0.4.0.jar and lib/librec-v1.4-alpha.jar hold the class files of
the hot path.  tests/tools/minijvm.py interprets buildModel() / predict() / isConverged() / updateLRate() and librec's
DenseMatrix / DenseVector / SparseMatrix iterator from those class files (tests/tools/carskit_jvm.py lists the few
container natives the harness supplies).  The vectors written here are therefore OUTPUTS OF THE REFERENCE ITSELF run in
this container, on seeded inputs: they pin the CPU oracle (tests/test_reference_bytecode.py, everywhere) and the CUDA
engine (tests/test_gpu_parity.py, on the GPU box, where /root/reference does not exist).

Per case: SHA-256 of the inputs (so a drifted input generator is reported as such), per-iteration loss and learning
rate (hex doubles), SHA-256 over the trained arrays' bytes, bounded predictions on held-out ratings (hex).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from carskit_b200 import capi, synth  # noqa: E402
from tests.golden.make_golden import digest, init_arrays  # noqa: E402

SPECS = [
    dict(name="pmf_f10", model="pmf", users=40, items=25, dims=None, nnz=700, F=10, iters=4, seed=11, order="user_sorted"),
    dict(name="biasedmf_f10", model="biasedmf", users=50, items=30, dims=None, nnz=900, F=10, iters=5, seed=12, order="user_sorted"),
    dict(name="camf_c_f10_8dims", model="camf_c", users=40, items=60, dims=[7, 7, 2, 3, 2, 9, 5, 4], nnz=900, F=10, iters=4, seed=13,
         order="shuffled"),
    dict(name="camf_ci_f64", model="camf_ci", users=60, items=25, dims=[8, 8, 8, 8], nnz=600, F=64, iters=3, seed=14, order="user_sorted"),
    dict(name="camf_ci_f7_shuffled_bold_halving", model="camf_ci", users=30, items=20, dims=[3, 2], nnz=700, F=7, iters=8, seed=15,
         order="shuffled", lrate=0.2),  # a rate large enough that the loss goes up twice and the bold driver halves it
    dict(name="camf_cu_f16_decay", model="camf_cu", users=40, items=20, dims=[4, 4, 4], nnz=800, F=16, iters=4, seed=16,
         order="user_sorted", bold_driver=False, decay=0.9, max_lrate=0.019),
    dict(name="camf_cuci_f12", model="camf_cuci", users=35, items=30, dims=[5, 4, 3], nnz=800, F=12, iters=4, seed=17, order="shuffled"),
    dict(name="camf_ics_f9", model="camf_ics", users=35, items=25, dims=[4, 3, 3], nnz=800, F=9, iters=4, seed=19, order="shuffled",
         lrate=0.005),
    dict(name="svdpp_f10", model="svdpp", users=40, items=25, dims=None, nnz=600, F=10, iters=4, seed=23, order="user_sorted"),
    dict(name="camf_lcs_f8", model="camf_lcs", users=35, items=25, dims=[4, 3, 3], nnz=800, F=8, iters=4, seed=21, order="shuffled",
         lrate=0.0002),  # U(0, 1) condition vectors of 10 factors: similarities of ~2.5 per dimension, predictions of ~30
    dict(name="camf_mcs_f8", model="camf_mcs", users=35, items=25, dims=[3, 4, 2], nnz=800, F=8, iters=4, seed=22, order="user_sorted",
         lrate=0.005),
]
FM_SPEC = dict(name="fm_k4", users=14, items=10, dims=[2, 3], nnz=160, k=4, iters=3, seed=18, reg_lw=0.01, reg_lf=0.02)


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def sgd_inputs(oracle, spec):
    model = capi.MODEL_NAMES[spec["model"]]
    ts, test = synth.make_training_set(spec["users"], spec["items"], spec["dims"], spec["nnz"], seed=spec["seed"],
                                       order=spec["order"], holdout=0.15)
    arrs = init_arrays(oracle, model, ts, spec["F"], seed=spec["seed"] + 100)
    return model, ts, test, arrs


def hyper(spec):
    return dict(lrate=spec.get("lrate", 0.02), bold_driver=spec.get("bold_driver", True), decay=spec.get("decay", -1.0),
                max_lrate=spec.get("max_lrate", -1.0))


def fm_inputs(oracle, spec):
    ts, test = synth.make_training_set(spec["users"], spec["items"], spec["dims"], spec["nnz"], seed=spec["seed"], holdout=0.2)
    p = ts.num_users + ts.num_items + ts.num_conditions
    g = oracle.JavaRandom(spec["seed"] + 100)
    arrs = {"w0": np.zeros(1), "w": g.uniform((p,)), "V": g.gaussian((p, spec["k"]))}
    return ts, test, arrs


def main():
    from oracle import oracle_py as oracle
    from tests.tools.carskit_jvm import ReferenceFM, ReferenceRun, available
    if not available():
        raise SystemExit("the reference jars are not here (/root/reference): the vectors can only be minted where they are")
    cases = []
    for spec in SPECS:
        model, ts, test, arrs = sgd_inputs(oracle, spec)
        t0 = time.time()
        rr = ReferenceRun(model, ts, arrs, spec["F"], **hyper(spec)).build_model(spec["iters"])
        out = rr.arrays()
        pred = rr.predict(test["u"], test["j"], test["ctx"], True)
        cases.append({"name": spec["name"], "spec": spec, "nnz": ts.nnz, "input_sha": sha(ts.u, ts.j, ts.ctx, ts.r, *[arrs[k] for k in sorted(arrs)]),
                      "losses_hex": [float(x).hex() for x in rr.losses], "lrates_hex": [float(x).hex() for x in rr.lrates],
                      "digest": digest(out), "pred_hex": [float(x).hex() for x in pred[:40]],
                      "bytecodes_executed": rr.jvm.ops_executed})
        print(spec["name"], f"{time.time() - t0:.1f}s", rr.jvm.ops_executed, "bytecodes; loss", rr.losses[-1], "lRate", rr.lrates[-1])
    spec = FM_SPEC
    ts, test, arrs = fm_inputs(oracle, spec)
    fr = ReferenceFM(ts, arrs, spec["k"], len(spec["dims"]), spec["reg_lw"], spec["reg_lf"]).build_model(spec["iters"])
    fm_case = {"name": spec["name"], "spec": spec, "nnz": ts.nnz, "input_sha": sha(ts.u, ts.j, ts.ctx, ts.r, arrs["w"], arrs["V"]),
               "digest": digest(fr.arrays()), "loss_hex": float(fr.rec.f["loss"]).hex(),
               "pred_hex": [float(x).hex() for x in fr.predict(test["u"], test["j"], test["ctx"], True)],
               "bytecodes_executed": fr.jvm.ops_executed}
    print(spec["name"], fr.jvm.ops_executed, "bytecodes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jvm_golden.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_jvm_golden.py",
                   "source": "reference class files interpreted by tests/tools/minijvm.py: jar/CARSKit-v0.4.0.jar, lib/librec-v1.4-alpha.jar",
                   "cases": cases, "fm": fm_case}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
