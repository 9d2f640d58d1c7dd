"""evalRankings' scoring loop and top-N cut (cars_rank_topn) against the oracle's literal restatement of
Recommender.java:797-824: ranked item ids, their scores and the survivor counts must be IDENTICAL (the north
star asks for bit-exact top-N indices), including tie order, the binThold filter and the rated-item exclusion."""
import numpy as np
import pytest

from carskit_b200 import capi, recommender, synth
from tests.golden.make_golden import REGS, init_arrays

pytestmark = pytest.mark.gpu
CTX_MODELS = (capi.CAMF_C, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI)


def make_queries(ts, test, rng, n):
    has_ctx = ts.ctx is not None
    idx = rng.choice(len(test["u"]), size=min(n, len(test["u"])), replace=False)
    keys = sorted({(int(test["u"][i]), int(test["ctx"][i]) if has_ctx else 0) for i in idx})
    qu = np.array([k[0] for k in keys], dtype=np.int32)
    qc = np.array([k[1] for k in keys], dtype=np.int32)
    trc = ts.ctx if has_ctx else np.zeros(ts.nnz, dtype=np.int32)
    rated = {}
    for u, j, c in zip(ts.u.tolist(), ts.j.tolist(), trc.tolist()):
        rated.setdefault((u, c), []).append(j)
    rptr, ritems = [0], []
    for k in keys:
        ritems.extend(rated.get(k, []))
        rptr.append(len(ritems))
    return qu, (qc if has_ctx else None), np.array(rptr, dtype=np.int64), np.array(ritems, dtype=np.int32)


@pytest.mark.parametrize("model,F", [(capi.PMF, 10), (capi.BIASEDMF, 7), (capi.CAMF_C, 16), (capi.CAMF_CI, 64),
                                     (capi.CAMF_CU, 200), (capi.CAMF_CUCI, 24)])
def test_topn_identical_to_oracle(oracle, cars_lib, model, F):
    dims = [4, 3] if model in CTX_MODELS else None
    ts, test = synth.make_training_set(300, 700, dims, 20000, seed=31, holdout=0.1)
    arrs = init_arrays(oracle, model, ts, F, seed=9)
    # exact ties: three items share one factor row and all their biases
    for a in ("Q", "item_bias", "ic_bias"):
        if a in arrs:
            arrs[a][[10, 400, 555]] = arrs[a][10]
    desc = capi.make_desc(ts, model, F, **REGS)
    rng = np.random.default_rng(1)
    qu, qc, rptr, ritems = make_queries(ts, test, rng, 400)
    cand = recommender.java_hashset_order(ts.j)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        for thold, nrec in ((-1.0, 10), (3.2, 5), (-1e9, 50), (1e9, 10)):
            got = eng.rank_topn(qu, qc, cand, rptr, ritems, thold, nrec)
            ref = oracle.rank_topn(desc, arrs, qu, qc, cand, rptr, ritems, thold, nrec)
            for g, r, name in zip(got, ref, ("items", "scores", "count", "kept")):
                assert np.array_equal(g, r), (name, thold, nrec)
        # without an exclusion list, and with a candidate list in another order (ties follow the candidate order)
        got = eng.rank_topn(qu, qc, cand[::-1].copy(), None, None, -1.0, 10)
        ref = oracle.rank_topn(desc, arrs, qu, qc, cand[::-1].copy(), None, None, -1.0, 10)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


def test_java_hashset_order_known_answers():
    # HashSet<Integer>: ids below the table capacity come out ascending; larger ids wrap around
    assert recommender.java_hashset_order([5, 3, 9, 3, 1]).tolist() == [1, 3, 5, 9]
    # 13 elements -> capacity 32 (13 > 0.75 * 16); 33 and 65 share bucket 1 with 1, in insertion order
    ids = [65, 33, 1] + list(range(2, 12))
    assert recommender.java_hashset_order(ids).tolist() == [65, 33, 1] + list(range(2, 12))
    # 70000 = 0x11170: hash 70000 ^ (70000 >>> 16) = 70001 -> bucket 1; 4464 = 0x1170 -> bucket 0
    assert recommender.java_hashset_order([70000, 4464]).tolist() == [4464, 70000]


def test_eval_rankings_mirror(oracle, cars_lib):
    ts, test = synth.make_training_set(120, 200, [3, 3], 6000, seed=8, holdout=0.15)
    rec = recommender.CAMF_CI(ts, test, conf={"num.factors": "16", "num.max.iter": "3"})
    rec.initModel(seed=2)
    rec.buildModel()
    lists = rec.evalRankings(numRecs=10, binThold=3.0)
    assert lists and all(len(x["ranked"]) <= 10 and x["correct"] for x in lists)
    for x in lists[:20]:
        s = x["scores"]
        assert all(s[i] >= s[i + 1] for i in range(len(s) - 1)) and all(v > 3.0 for v in s)
        trained = {int(j) for u, j, c in zip(ts.u, ts.j, ts.ctx) if u == x["u"] and c == x["c"]}
        assert not trained & set(x["ranked"])
