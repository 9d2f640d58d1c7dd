"""The parity pin: the CPU oracle against the REFERENCE'S OWN BYTECODE.

Three layers, from strongest to weakest:
  1. live      -- where /root/reference exists (this container): buildModel() / predict() of every model, isConverged(),
                  updateLRate() and librec's containers are interpreted from the class files in jar/CARSKit-v0.4.0.jar and
                  lib/librec-v1.4-alpha.jar by tests/tools/minijvm.py, and the oracle must reproduce P, Q, every bias, the
                  per-iteration loss and the learning rate BIT FOR BIT;
  2. committed -- everywhere (the GPU box has no /root/reference): tests/golden/jvm_golden.json was minted by the same
                  execution (tests/golden/make_jvm_golden.py); the oracle must reproduce its digests;
  3. skeleton  -- the floating-point instruction sequence of the hot methods, read from the class files: what the oracle's
                  statement order assumes (separately rounded dmul / dadd / dsub, f2d on every regulariser, f ascending,
                  no fused operations), asserted opcode by opcode.
"""
import json
import os

import numpy as np
import pytest

from carskit_b200 import capi, synth
from tests.golden import make_jvm_golden as G
from tests.golden.make_golden import REGS, digest, init_arrays
from tests.tools import carskit_jvm

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "jvm_golden.json")))
needs_reference = pytest.mark.skipif(not carskit_jvm.available(), reason="the reference jars (/root/reference) are not on this box")


def oracle_run(oracle, model, ts, arrs, F, iters, lrate=0.02, bold_driver=True, decay=-1.0, max_lrate=-1.0):
    """buildModel() on the oracle: epoch + isConverged/updateLRate exactly as the tests of the product drive it."""
    desc = capi.make_desc(ts, model, F, **REGS)
    import ctypes as C
    st = oracle.new_state(capi.f32(lrate), bold_driver=bold_driver, decay=decay, max_lrate=max_lrate)
    losses, lrates = [], []
    for it in range(1, iters + 1):
        lrates.append(st.lRate)
        st.loss = oracle.epoch(desc, arrs, st.lRate)
        losses.append(st.loss)
        if oracle.lib().oracle_is_converged(C.byref(st), it):
            break
    return losses, lrates


# ---- 1. live ---------------------------------------------------------------------------------------------------------
LIVE = [
    ("pmf", None, "user_sorted", 5), ("biasedmf", None, "user_sorted", 6), ("svdpp", None, "user_sorted", 5),
    ("camf_c", [3, 2, 2], "shuffled", 5),
    ("camf_ci", [3, 2, 2], "user_sorted", 9), ("camf_cu", [4, 3], "shuffled", 4), ("camf_cuci", [2, 2, 3], "user_sorted", 5),
    ("camf_ics", [3, 3, 2], "shuffled", 6), ("camf_lcs", [3, 3, 2], "shuffled", 6), ("camf_mcs", [3, 2, 3], "user_sorted", 5),
]


@needs_reference
@pytest.mark.parametrize("name,dims,order,F", LIVE)
def test_oracle_reproduces_the_reference_bytecode_live(oracle, name, dims, order, F):
    model = capi.MODEL_NAMES[name]
    ts, test = synth.make_training_set(25, 14, dims, 350, seed=F, order=order, holdout=0.1)
    arrs = init_arrays(oracle, model, ts, F, seed=F + 7)
    lrate = 0.005 if name in ("camf_ics", "camf_lcs", "camf_mcs") else 0.02  # the similarity models start from U(0, 1) factors: predictions of ~F/4
    ref = carskit_jvm.ReferenceRun(model, ts, arrs, F, lrate=lrate).build_model(3)
    got = {k: v.copy() for k, v in arrs.items()}
    losses, lrates = oracle_run(oracle, model, ts, got, F, 3, lrate=lrate)
    out = ref.arrays()
    for k in out:
        assert np.array_equal(out[k], got[k]), k
    assert [x.hex() for x in losses] == [x.hex() for x in ref.losses]
    assert [float(x).hex() for x in lrates] == [float(x).hex() for x in ref.lrates]
    desc = capi.make_desc(ts, model, F, **REGS)
    p_ref = ref.predict(test["u"], test["j"], test["ctx"], True)
    p_got = oracle.predict(desc, got, test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    assert np.array_equal(p_ref, p_got)


@needs_reference
def test_ragged_contexts_and_empty_rows_live(oracle):
    # contexts of different lengths (one empty); user-item pairs whose CRS rows the iterator has to skip over
    rng = np.random.default_rng(3)
    ts, _ = synth.make_training_set(12, 9, [3, 4, 2], 200, seed=8)
    ctxs = [[], [0], [3, 0], [1, 4, 7], [7], [8, 2], [5, 8, 1]]
    ptr = np.cumsum([0] + [len(c) for c in ctxs]).astype(np.int32)
    cond = np.array([x for c in ctxs for x in c], dtype=np.int32)
    ts2 = capi.TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=ts.u, j=ts.j, r=ts.r,
                           ctx=np.sort(rng.integers(0, len(ctxs), ts.nnz)).astype(np.int32) * 0 + rng.integers(0, len(ctxs), ts.nnz).astype(np.int32),
                           num_conditions=9, num_contexts=len(ctxs), ctx_ptr=ptr, ctx_cond=cond, global_mean=ts.global_mean)
    # CRS order inside a pair is context-ascending: re-sort the drawn contexts inside every (u, j) run
    key = ts2.u.astype(np.int64) * ts2.num_items + ts2.j
    order = np.lexsort((ts2.ctx, key))
    ts2.ctx = np.ascontiguousarray(ts2.ctx[order])
    keep = np.ones(ts2.nnz, dtype=bool)
    keep[1:] = (key[order][1:] != key[order][:-1]) | (ts2.ctx[1:] != ts2.ctx[:-1])  # unique (pair, ctx)
    ts2 = capi.TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=ts2.u[keep], j=ts2.j[keep], r=ts2.r[keep],
                           ctx=ts2.ctx[keep], num_conditions=9, num_contexts=len(ctxs), ctx_ptr=ptr, ctx_cond=cond,
                           global_mean=ts.global_mean)
    for model in (capi.CAMF_C, capi.CAMF_CI):
        arrs = init_arrays(oracle, model, ts2, 6, seed=2)
        ref = carskit_jvm.ReferenceRun(model, ts2, arrs, 6).build_model(2)
        got = {k: v.copy() for k, v in arrs.items()}
        losses, _ = oracle_run(oracle, model, ts2, got, 6, 2)
        out = ref.arrays()
        for k in out:
            assert np.array_equal(out[k], got[k]), k
        assert losses == ref.losses


@needs_reference
def test_fm_dense_oracle_reproduces_the_reference_bytecode_live(oracle):
    spec = dict(G.FM_SPEC, users=9, items=7, nnz=70, k=3, iters=2, seed=5)
    ts, test, arrs = G.fm_inputs(oracle, spec)
    ref = carskit_jvm.ReferenceFM(ts, arrs, spec["k"], len(spec["dims"]), spec["reg_lw"], spec["reg_lf"]).build_model(spec["iters"])
    prob = oracle.fm_problem(ts, spec["k"], len(spec["dims"]), np.float32(spec["reg_lw"]), np.float32(spec["reg_lf"]))
    got = {k: v.copy() for k, v in arrs.items()}
    oracle.fm_dense_build(prob, got, spec["iters"])
    out = ref.arrays()
    for k in out:
        assert np.array_equal(out[k], got[k]), k
    p_ref = ref.predict(test["u"], test["j"], test["ctx"], True)
    p_got = oracle.fm_predict(prob, got, test["u"], test["j"], test["ctx"], bound=True, lo=1.0, hi=5.0)
    assert np.array_equal(p_ref, p_got)


# ---- 2. committed vectors ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_oracle_reproduces_the_committed_reference_vectors(oracle, case):
    spec = case["spec"]
    model, ts, test, arrs = G.sgd_inputs(oracle, spec)
    assert G.sha(ts.u, ts.j, ts.ctx, ts.r, *[arrs[k] for k in sorted(arrs)]) == case["input_sha"], \
        "the seeded INPUTS differ from the ones the reference bytecode was run on (generator drift, not a parity failure)"
    losses, lrates = oracle_run(oracle, model, ts, arrs, spec["F"], spec["iters"], **G.hyper(spec))
    assert [float(x).hex() for x in losses] == case["losses_hex"]
    assert [float(x).hex() for x in lrates] == case["lrates_hex"]
    assert digest(arrs) == case["digest"]
    desc = capi.make_desc(ts, model, spec["F"], **REGS)
    pred = oracle.predict(desc, arrs, test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    assert [float(x).hex() for x in pred[:40]] == case["pred_hex"]


def test_fm_oracle_reproduces_the_committed_reference_vector(oracle):
    case = GOLDEN["fm"]
    spec = case["spec"]
    ts, test, arrs = G.fm_inputs(oracle, spec)
    assert G.sha(ts.u, ts.j, ts.ctx, ts.r, arrs["w"], arrs["V"]) == case["input_sha"]
    prob = oracle.fm_problem(ts, spec["k"], len(spec["dims"]), np.float32(spec["reg_lw"]), np.float32(spec["reg_lf"]))
    oracle.fm_dense_build(prob, arrs, spec["iters"])
    assert digest(arrs) == case["digest"]
    pred = oracle.fm_predict(prob, arrs, test["u"], test["j"], test["ctx"], bound=True, lo=1.0, hi=5.0)
    assert [float(x).hex() for x in pred] == case["pred_hex"]


# ---- 3. opcode skeleton ---------------------------------------------------------------------------------------------------
@needs_reference
def test_arithmetic_skeleton_of_the_hot_methods():
    from tests.tools.classfile import Jar, arithmetic_skeleton
    jar = Jar(*carskit_jvm.JARS)
    dm = jar.load("librec/data/DenseMatrix")
    # res += m.data[mrow][j] * n.data[nrow][j], j ascending: ONE multiply then ONE add per factor, no fused op exists in
    # the class file (and javac never emits one)
    assert arithmetic_skeleton(dm.method("rowMult")) == ["AssertionError.<init>", "DenseMatrix.get", "DenseMatrix.get", "dmul", "dadd"]
    row_mult = dm.method("rowMult").code
    assert [i.op for i in row_mult if i.op == "iinc"] == ["iinc"] and [i.args for i in row_mult if i.op == "iinc"] == [(6, 1)]
    assert arithmetic_skeleton(dm.method("add", "(IID)V")) == ["dadd"]
    assert arithmetic_skeleton(jar.load("librec/data/DenseVector").method("add", "(ID)V")) == ["dadd"]
    # DenseMatrix.init(mean, sigma): row-major Randoms.gaussian (the time-seeded librec generator)
    assert [s for s in arithmetic_skeleton(dm.method("init", "(DD)V")) if "." in s] == ["Randoms.gaussian"]
    ci = jar.load(carskit_jvm.CLASS_OF[capi.CAMF_CI]).method("buildModel")
    assert ci.lines[0] == 77  # the class file was compiled from the CAMF_CI.java under src/ (line numbers agree)
    want = ["SparseMatrix.iterator", "Iterator.hasNext", "Iterator.next", "MatrixEntry.row", "DataDAO.getUserIdFromUI",
            "DataDAO.getItemIdFromUI", "MatrixEntry.column", "MatrixEntry.get", "CAMF_CI.predict",
            "dsub",                                    # e = r - pred                         :89
            "dmul", "dadd",                            # loss += e * e                        :91
            "DenseVector.get", "f2d", "dmul", "dsub",  # sgd = e - regB * bu                  :94-95
            "dmul", "DenseVector.add",                 # userBias.add(u, lRate * sgd)         :96
            "f2d", "dmul", "dmul", "dadd",             # loss += regB * bu * bu               :98
            "CAMF_CI.getConditions", "List.iterator", "Iterator.hasNext", "Iterator.next", "Integer.intValue",
            "DenseMatrix.get", "Math.pow", "dadd",     # Bic_sum += pow(Bic, 2)               :102-103
            "f2d", "dmul", "dsub",                     # sgd = e - regC * Bic                 :104
            "dmul", "dadd", "DenseMatrix.set",         # icBias.set(j, c, Bic + lRate * sgd)  :105
            "f2d", "dmul", "dadd",                     # loss += regC * Bic_sum               :108
            "DenseMatrix.get", "DenseMatrix.get",      # puf, qjf                             :111-112
            "dmul", "f2d", "dmul", "dsub",             # delta_u = e * qjf - regU * puf       :114
            "dmul", "f2d", "dmul", "dsub",             # delta_j = e * puf - regI * qjf       :115
            "dmul", "DenseMatrix.add", "dmul", "DenseMatrix.add",  # P.add(lRate * delta_u); Q.add(lRate * delta_j)  :117-118
            "f2d", "dmul", "dmul", "f2d", "dmul", "dmul", "dadd", "dadd",  # loss += regU*puf*puf + regI*qjf*qjf  :120
            "dmul",                                    # loss *= 0.5                          :124
            "CAMF_CI.isConverged"]
    assert arithmetic_skeleton(ci) == want
    up = jar.load("carskit/generic/IterativeRecommender").method("updateLRate")
    consts = [i.ref[1] for i in up.code if i.op == "ldc2_w"]
    assert consts == [1.05, 0.5]
