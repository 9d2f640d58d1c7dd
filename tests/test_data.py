"""The input contract: DataDAO.readData's id / layout rules (carskit_b200/data.py) against the known answer
hand-traced from the reference's own sample file.  The file lives in the reference tree,
which only exists in the build container; elsewhere the test skips."""
import os

import numpy as np
import pytest

from carskit_b200 import data

SAMPLE = "/root/reference/sampleData/train_binary.csv"


@pytest.mark.skipif(not os.path.exists(SAMPLE), reason="reference sample data not present on this box")
def test_sample_layout_known_answer():
    ts, dao = data.read_binary_csv(SAMPLE)
    assert (dao.numUsers(), dao.numItems(), dao.numUserItems(), dao.numContexts(), dao.numConditions()) == (17, 2, 18, 8, 10)
    assert dao.dimIds == {"companion": 0, "location": 1, "time": 2} and dao.numContextDims() == 3
    assert dao.EmptyContextConditions == [2, 6, 7]
    assert dao.ctxIds == {"0,5,9": 0, "0,5,8": 1, "1,4,9": 2, "1,5,9": 3, "0,4,8": 4, "1,4,8": 5, "1,5,8": 6, "0,4,9": 7}
    assert ts.nnz == 20 and ts.global_mean == 3.95
    want = [(0, 0, 0, 0, 4), (1, 1, 1, 0, 5), (2, 2, 2, 1, 5), (2, 3, 2, 1, 5), (3, 3, 3, 1, 4), (4, 4, 4, 0, 5),
            (5, 5, 5, 1, 4), (6, 1, 6, 0, 4), (7, 6, 7, 1, 4), (8, 5, 8, 1, 4), (8, 6, 8, 1, 4), (9, 6, 9, 1, 1),
            (10, 3, 10, 1, 4), (11, 2, 11, 1, 4), (12, 4, 12, 0, 3), (13, 7, 13, 0, 2), (14, 4, 14, 0, 3),
            (15, 7, 10, 0, 4), (16, 5, 15, 1, 5), (17, 2, 16, 1, 5)]
    got = list(zip(ts.pair_ids.tolist(), ts.ctx.tolist(), ts.u.tolist(), ts.j.tolist(), ts.r.tolist()))
    assert got == [(a, b, c, d, float(e)) for a, b, c, d, e in want]
    assert ts.ctx_ptr.tolist() == [0, 3, 6, 9, 12, 15, 18, 21, 24]
    assert ts.ctx_cond[:3].tolist() == [0, 5, 9] and dao.ratingScale == [1.0, 2.0, 3.0, 4.0, 5.0]


def test_rules_on_a_synthetic_file(tmp_path):
    p = tmp_path / "r.csv"
    p.write_text("user,item,rating,time:na,time:day,time:night,place:na,place:home\n"
                 "b,y,3,0,1,0,0,1\n"
                 "a,x,5,1,0,0,1,0\n"
                 "b,y,4,0,0,1,0,1\n"
                 "b,x,0,1,0,0,1,0\n"      # a zero rating is not stored
                 "b,y,2,0,1,0,0,1\n"      # the same (pair, context) again: the last one wins
                 "a,y,1,0,0,1,0,1\n")
    ts, dao = data.read_binary_csv(str(p))
    assert dao.userIds == {"b": 0, "a": 1} and dao.itemIds == {"y": 0, "x": 1}
    assert dao.uiIds == {"0,0": 0, "1,1": 1, "0,1": 2, "1,0": 3}
    assert dao.ctxIds == {"1,4": 0, "0,3": 1, "2,4": 2} and dao.EmptyContextConditions == [0, 3]
    # CRS order: pair 0 (ctx 0, ctx 2), pair 1 (ctx 1), [pair 2 only had the zero], pair 3 (ctx 2)
    assert list(zip(ts.u.tolist(), ts.j.tolist(), ts.ctx.tolist(), ts.r.tolist())) == \
        [(0, 0, 0, 2.0), (0, 0, 2, 4.0), (1, 1, 1, 5.0), (1, 0, 2, 1.0)]
    assert ts.global_mean == 12.0 / 4 and ts.num_conditions == 5 and ts.ctx_cond.tolist() == [1, 4, 0, 3, 2, 4]


# ---------------------------------------------------------------------------------------------------
# DataSplitter (k-fold assignment) and the vectorised java.util.Random behind it
# ---------------------------------------------------------------------------------------------------
def test_java_random_doubles_known_answers(oracle):
    # published java.util.Random outputs
    assert data.java_random_doubles(0, 1)[0] == 0.730967787376657
    assert data.java_random_doubles(42, 2).tolist() == [0.7275636800328681, 0.6832234717598454]
    for seed in (1, 20261017, -5, 2 ** 40 + 3):
        g = oracle.JavaRandom(seed)
        want = [g.next_double() for _ in range(1000)]
        assert data.java_random_doubles(seed, 1000).tolist() == want
    assert data.java_random_doubles(7, 0).shape == (0,)


def literal_split_folds(n, kfold, seed, oracle):
    """DataSplitter.splitFolds (:102-133) as written, with a plain Lomuto quicksort carrying the labels."""
    g = oracle.JavaRandom(seed)
    num_fold = min(kfold, n)
    indv = (n + 0.0) / num_fold
    rdm = [g.next_double() for _ in range(n)]
    fold = [int(i / indv) + 1 for i in range(n)]

    def qs(lo, hi):
        while lo < hi:
            p, i = rdm[hi], lo
            for k in range(lo, hi):
                if rdm[k] < p:
                    rdm[i], rdm[k] = rdm[k], rdm[i]
                    fold[i], fold[k] = fold[k], fold[i]
                    i += 1
            rdm[i], rdm[hi] = rdm[hi], rdm[i]
            fold[i], fold[hi] = fold[hi], fold[i]
            qs(lo, i - 1)
            lo = i + 1
    qs(0, n - 1)
    return fold


@pytest.mark.parametrize("n,kfold,seed", [(20, 5, 1), (997, 5, 1), (1000, 10, 20261017), (3, 5, 9)])
def test_data_splitter_matches_the_literal_algorithm(oracle, n, kfold, seed):
    from carskit_b200 import synth
    ts, _ = synth.make_training_set(50, 40, [3, 2], 4 * n, seed=3, order="shuffled")
    keep = np.arange(ts.nnz) < n
    ts = data.TrainingSet(num_users=50, num_items=40, u=ts.u[keep], j=ts.j[keep], r=ts.r[keep], ctx=ts.ctx[keep],
                          num_conditions=ts.num_conditions, num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr,
                          ctx_cond=ts.ctx_cond, global_mean=0.0)
    assert ts.nnz == n
    sp = data.DataSplitter(ts, kfold, seed)
    assert sp.assign.tolist() == literal_split_folds(n, kfold, seed, oracle)
    assert sp.getKthFold(0) is None and sp.getKthFold(sp.numFold + 1) is None
    sizes = []
    for k in range(1, sp.numFold + 1):
        train, test = sp.getKthFold(k)
        sizes.append(len(test["r"]))
        assert train.nnz + len(test["r"]) == n
        m = sp.assign == k
        assert train.u.tolist() == ts.u[~m].tolist() and test["ctx"].tolist() == ts.ctx[m].tolist()  # order kept
        assert train.global_mean == (sum(train.r.tolist()) / train.nnz if train.nnz else 0.0)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
