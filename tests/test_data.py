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
