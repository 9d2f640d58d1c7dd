"""FM oracle (oracle/fm_oracle.cpp): the literal dense restatement of FM.java against its sparse form.

Parity chain for FM:
  FM.java  ==(line-by-line reading, unpinned)==  dense oracle  ==(bit-identical)==  sparse oracle, sequential
  denominators  ==(<= 1e-10 relative: summation order of N equal terms)==  sparse oracle, closed-form
  denominators  ==(tree-ordered sums, tests/test_fm_gpu.py)==  CUDA engine.
"""
import numpy as np
import pytest

from carskit_b200 import capi, synth


def fm_inputs(oracle, users, items, dims, nnz, k, seed, holdout=0.0):
    ts, test = synth.make_training_set(users, items, dims, nnz, seed=seed, order="user_sorted", holdout=holdout)
    D = len(dims)
    prob = oracle.fm_problem(ts, k, D, np.float32(0.01), np.float32(0.02))
    p = ts.num_users + ts.num_items + ts.num_conditions
    g = oracle.JavaRandom(seed + 50)
    arrs = {"w0": np.zeros(1), "w": g.uniform((p,)), "V": g.gaussian((p, k))}
    return ts, test, prob, arrs


def clone(arrs):
    return {k: v.copy() for k, v in arrs.items()}


def test_dense_and_sparse_predict_agree(oracle):
    # contexts 0..5 but only 5 conditions: context id 5 has index U+I+5 == p, i.e. NO feature (FM.java:81)
    ts, _, prob, arrs = fm_inputs(oracle, 12, 9, [2, 3], 150, 3, seed=1)
    assert ts.num_contexts == 6 and ts.num_conditions == 5 and int(ts.ctx.max()) == 5
    arrs["w0"][0] = 0.3
    for n in range(0, ts.nnz, 7):
        a = oracle.fm_dense_predict(prob, arrs, int(ts.u[n]), int(ts.j[n]), int(ts.ctx[n]))
        b = oracle.fm_predict(prob, arrs, ts.u[n:n + 1], ts.j[n:n + 1], ts.ctx[n:n + 1])[0]
        assert a == b


@pytest.mark.parametrize("dims,k", [([2, 3], 3), ([4], 2), ([2, 2, 2], 4)])
def test_sparse_form_is_bit_identical_to_the_literal_algorithm(oracle, dims, k):
    ts, _, prob, arrs = fm_inputs(oracle, 12, 9, dims, 160, k, seed=2)
    dense, sparse = clone(arrs), clone(arrs)
    iters = 3
    _, e_d, Q_d = oracle.fm_dense_build(prob, dense, iters)
    e_s, Q_s = oracle.fm_prepare(prob, sparse)
    for _ in range(iters):
        oracle.fm_iteration(prob, sparse, e_s, Q_s, closed_den=False)
    for name in ("w0", "w", "V"):
        assert np.array_equal(dense[name], sparse[name]), name
    assert np.array_equal(e_d, e_s) and np.array_equal(Q_d, Q_s)


def test_closed_form_denominators_differ_by_rounding_only(oracle):
    ts, _, prob, arrs = fm_inputs(oracle, 60, 40, [3, 4], 2500, 5, seed=3)
    a, b = clone(arrs), clone(arrs)
    ea, Qa = oracle.fm_prepare(prob, a)
    eb, Qb = oracle.fm_prepare(prob, b)
    la = lb = 0.0
    for _ in range(4):
        la = oracle.fm_iteration(prob, a, ea, Qa, closed_den=False)
        lb = oracle.fm_iteration(prob, b, eb, Qb, closed_den=True)
    for name in ("w0", "w", "V"):
        np.testing.assert_allclose(b[name], a[name], rtol=1e-10, atol=1e-13)
    assert abs(la - lb) <= 1e-10 * abs(la)
    assert lb < 1e9 and np.isfinite(lb)


def test_reference_quirks_are_reproduced_not_repaired(oracle):
    """FM.java:133 caches errors = r - predict while its ALS steps (:159, :176-185) are the textbook formulas
    for e = predict - r, and the V step moves the cache by delta*x (:209) where the prediction moves by
    delta*h.  The cache is therefore NOT r - predict after training; the sweep minimises the cached objective
    only.  This repo reproduces that arithmetic: the test pins the observable
    consequences so nobody "fixes" the oracle or the kernels independently of the reference."""
    ts, _, prob, arrs = fm_inputs(oracle, 80, 50, [3, 3], 3000, 4, seed=4)
    pred0 = oracle.fm_predict(prob, arrs, ts.u, ts.j, ts.ctx)
    e, Q = oracle.fm_prepare(prob, arrs)
    np.testing.assert_allclose(e, ts.r - pred0, rtol=0, atol=1e-12)
    sse0 = float(np.sum(e * e))
    for _ in range(5):
        oracle.fm_iteration(prob, arrs, e, Q, closed_den=True)
    assert float(np.sum(e * e)) < 0.5 * sse0  # the cached objective goes down ...
    pred = oracle.fm_predict(prob, arrs, ts.u, ts.j, ts.ctx)
    true_sse0, true_sse = float(np.sum((ts.r - pred0) ** 2)), float(np.sum((ts.r - pred) ** 2))
    assert true_sse > true_sse0  # ... while the model's real training error goes UP (sign convention)
    assert np.max(np.abs(e - (ts.r - pred))) > 1.0


def test_dense_oracle_reproduces_fm_golden(oracle):
    import json
    import os
    from tests.golden.make_fm_golden import run
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fm_golden.json")))
    got = run(oracle)
    for k in ("w0", "w", "V", "pred"):
        assert got[k] == g[k], k
