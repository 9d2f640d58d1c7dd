"""FM on the GPU (cars_fm_* through the C ABI) against the sparse CPU oracle with the same closed-form
denominators.  The engine's sums are tree-ordered, the oracle's sequential: equality is up to summation
order -- tolerances below; the model-level bar of the north star (predictions / RMSE within 1e-5) is
asserted against the LITERAL dense oracle on an input it can afford."""
import math

import numpy as np
import pytest

from carskit_b200 import capi, recommender
from tests.test_fm_oracle import clone, fm_inputs

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-12


def run_gpu(ts, arrs, k, D, iters, tuning=None):
    desc = capi.make_desc(ts, capi.FM, k, reg_lw=float(np.float32(0.01)), reg_lf=float(np.float32(0.02)), num_context_dims=D,
                          tuning=tuning)
    got = clone(arrs)
    losses = []
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(got)
        eng.prepare()
        for _ in range(iters):
            losses.append(eng.iteration())
        eng.download(got)
        st = eng.stats()
    return got, losses, st


@pytest.mark.parametrize("users,items,dims,nnz,k", [(12, 9, [2, 3], 150, 3), (300, 120, [4, 8], 20000, 16),
                                                    (2000, 50, [32], 60000, 8), (50, 2000, [3, 3, 3], 30000, 64)])
@pytest.mark.parametrize("block_rows", [None, "700"])
def test_fm_matches_sparse_oracle(oracle, cars_lib, users, items, dims, nnz, k, block_rows):
    # block_rows: the engine's internal row order (item block, context, caller's order) with tiny blocks, so the
    # permuted layout the big inputs use is exercised at oracle sizes; results do not depend on it beyond rounding
    # (fm_dense_min_rows=0: and the streaming reduce of the context field); knobs travel in cars_desc.tuning
    tuning = None if block_rows is None else f"fm_block_rows={block_rows};fm_dense_min_rows=0"
    ts, _, prob, arrs = fm_inputs(oracle, users, items, dims, nnz, k, seed=7)
    iters = 3
    got, losses, st = run_gpu(ts, arrs, k, len(dims), iters, tuning)
    ref = clone(arrs)
    e, Q = oracle.fm_prepare(prob, ref)
    ref_losses = [oracle.fm_iteration(prob, ref, e, Q, closed_den=True) for _ in range(iters)]
    for name in ("w0", "w", "V"):
        np.testing.assert_allclose(got[name], ref[name], rtol=RTOL, atol=ATOL, err_msg=name)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-10)
    assert st.kernel_launches > 0 and st.nnz == ts.nnz


@pytest.mark.parametrize("tuning", ["fm_block_rows=700;fm_dense_min_rows=0;fm_run_users=256",
                                    "fm_block_rows=300;fm_dense_min_rows=0;fm_run_users=64",
                                    "fm_block_rows=700;fm_dense_min_rows=0;fm_run_fuse_update=0",
                                    "fm_block_rows=700;fm_dense_min_rows=0;fm_runs=0"])
@pytest.mark.parametrize("order", ["user_sorted", "shuffled"])
def test_fm_user_field_run_reduce(oracle, cars_lib, tuning, order):
    """fm_run_reduce_kernel (users: contiguous runs of the (item block, user) storage order; step (b) in its epilogue)
    against the oracle -- for a caller's order that is NOT user-sorted too (the engine sorts the rows of a block by user
    itself), with groups wider than the field, and with the gathering piece reduce it replaces (fm_runs=0)."""
    ts, _, prob, arrs = fm_inputs(oracle, 700, 90, [4, 8], 40000, 8, seed=21)
    if order == "shuffled":
        perm = np.random.default_rng(5).permutation(ts.nnz)
        ts.u, ts.j, ts.ctx, ts.r = (np.ascontiguousarray(a[perm]) for a in (ts.u, ts.j, ts.ctx, ts.r))
        prob = oracle.fm_problem(ts, 8, 2, np.float32(0.01), np.float32(0.02))
    iters = 3
    got, losses, _ = run_gpu(ts, arrs, 8, 2, iters, tuning)
    ref = clone(arrs)
    e, Q = oracle.fm_prepare(prob, ref)
    ref_losses = [oracle.fm_iteration(prob, ref, e, Q, closed_den=True) for _ in range(iters)]
    for name in ("w0", "w", "V"):
        np.testing.assert_allclose(got[name], ref[name], rtol=RTOL, atol=ATOL, err_msg=name)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-10)


def test_fm_sharded_entry_point_with_one_shard_equals_the_plain_iteration(oracle, cars_lib):
    """cars_fm_iteration_sharded with a no-op all-reduce (one shard): the split coordinate step (partial sums ->
    all-reduce -> finish) after fm_run_reduce_kernel gives the fused epilogue's result bit for bit."""
    import torch
    ts, _, prob, arrs = fm_inputs(oracle, 500, 80, [4, 8], 30000, 8, seed=23)
    tuning = "fm_block_rows=700;fm_dense_min_rows=0"
    a, la, _ = run_gpu(ts, arrs, 8, 2, 2, tuning)
    desc = capi.make_desc(ts, capi.FM, 8, reg_lw=float(np.float32(0.01)), reg_lf=float(np.float32(0.02)), num_context_dims=2,
                          tuning=tuning)
    b = clone(arrs)
    lb = []
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(b)
        eng.prepare()
        buf = torch.zeros(eng.exchange_doubles(), dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        for _ in range(2):
            lb.append(eng.iteration_sharded(buf.data_ptr(), lambda ptr, count: None))
        eng.download(b)
    assert la == lb
    for name in ("w0", "w", "V"):
        assert np.array_equal(a[name], b[name]), name


def test_fm_tiled_prepass_is_bit_identical_to_the_plain_one(oracle, cars_lib):
    """fm_prepare_tiled_kernel (coordinate-major scratch copy of V, 16 factors of 32 rows staged per step) performs
    fm_prepare_kernel's operations in the same order: e and Qc, hence everything trained from them, are bit-identical.
    k = 21: a ragged last factor step; contexts with index >= p (no feature) are in fm_inputs' data."""
    for users, items, dims, nnz, k in [(300, 120, [4, 8], 20000, 21), (12, 9, [2, 3], 150, 3), (500, 64, [6], 9000, 64)]:
        ts, _, prob, arrs = fm_inputs(oracle, users, items, dims, nnz, k, seed=31)
        a, la, _ = run_gpu(ts, arrs, k, len(dims), 2, "fm_prepare_tiled=0")
        b, lb, _ = run_gpu(ts, arrs, k, len(dims), 2, "fm_prepare_tiled=1")
        assert la == lb
        for name in ("w0", "w", "V"):
            assert np.array_equal(a[name], b[name]), (name, k)


def test_fm_within_1e5_of_the_literal_algorithm(oracle, cars_lib):
    ts, test, prob, arrs = fm_inputs(oracle, 40, 30, [2, 3], 900, 4, seed=9, holdout=0.15)
    dense = clone(arrs)
    iters = 5
    oracle.fm_dense_build(prob, dense, iters)
    got, _, _ = run_gpu(ts, arrs, 4, 2, iters)
    desc = capi.make_desc(ts, capi.FM, 4, reg_lw=float(np.float32(0.01)), reg_lf=float(np.float32(0.02)), num_context_dims=2)
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(got)
        p_gpu = eng.predict(test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    p_ref = oracle.fm_predict(prob, dense, test["u"], test["j"], test["ctx"], bound=True, lo=1.0, hi=5.0)
    assert np.max(np.abs(p_gpu - p_ref)) < 1e-5
    rmse = lambda p: math.sqrt(float(np.mean((test["r"] - p) ** 2)))
    assert abs(rmse(p_gpu) - rmse(p_ref)) < 1e-5


def test_fm_predict_bit_identical_to_oracle(oracle, cars_lib):
    ts, test, prob, arrs = fm_inputs(oracle, 100, 60, [3, 5], 4000, 10, seed=11, holdout=0.2)
    arrs["w0"][0] = 0.25
    desc = capi.make_desc(ts, capi.FM, 10, reg_lw=0.01, reg_lf=0.02, num_context_dims=2)
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        for bound in (False, True):
            a = eng.predict(test["u"], test["j"], test["ctx"], bound=bound, min_rate=1.0, max_rate=5.0)
            b = oracle.fm_predict(prob, arrs, test["u"], test["j"], test["ctx"], bound=bound, lo=1.0, hi=5.0)
            assert np.array_equal(a, b)


def test_fm_recommender_mirror_and_errors(oracle, cars_lib):
    ts, test, prob, arrs = fm_inputs(oracle, 150, 80, [4, 4], 8000, 8, seed=13, holdout=0.1)
    rec = recommender.getRecommender("fm")(ts, test, conf={"num.factors": "8", "num.max.iter": "4", "FM": "-lw 0.01 -lf 0.02"})
    m = rec.execute(init=arrs)
    assert len(rec.iter_losses) == 4 and rec.iter_losses[-1] < rec.iter_losses[0]
    ref = clone(arrs)
    e, Q = oracle.fm_prepare(prob, ref)
    for _ in range(4):
        oracle.fm_iteration(prob, ref, e, Q, closed_den=True)
    p_ref = oracle.fm_predict(prob, ref, test["u"], test["j"], test["ctx"], bound=True, lo=1.0, hi=5.0)
    assert abs(m["RMSE"] - math.sqrt(float(np.mean((test["r"] - p_ref) ** 2)))) < 1e-9
    # call-order and argument errors
    desc = capi.make_desc(ts, capi.FM, 8, reg_lw=0.01, reg_lf=0.02, num_context_dims=2)
    with capi.FmEngine(desc, keepalive=ts) as eng:
        with pytest.raises(capi.CarsError) as ex:
            eng.iteration()
        assert ex.value.code == -6
    with pytest.raises(capi.CarsError):
        capi.FmEngine(capi.make_desc(ts, capi.FM, 8, num_context_dims=0), keepalive=ts)
    keep = int(ts.j[11])
    ts.j[11] = ts.num_items  # range check on the device
    with pytest.raises(capi.CarsError) as ex:
        capi.FmEngine(capi.make_desc(ts, capi.FM, 8, num_context_dims=2), keepalive=ts)
    assert ex.value.code == -1 and "rating 11 " in str(ex.value)
    ts.j[11] = keep
    with pytest.raises(capi.CarsError):
        capi.Engine(capi.make_desc(ts, capi.FM, 8, num_context_dims=2), keepalive=ts)


def test_fm_against_committed_golden_vectors(oracle, cars_lib):
    import json
    import os
    from tests.golden.make_fm_golden import SPEC
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fm_golden.json")))
    ts, test, prob, arrs = fm_inputs(oracle, SPEC["users"], SPEC["items"], SPEC["dims"], SPEC["nnz"], SPEC["k"],
                                     seed=SPEC["seed"], holdout=SPEC["holdout"])
    got, _, _ = run_gpu(ts, arrs, SPEC["k"], len(SPEC["dims"]), SPEC["iters"])
    np.testing.assert_allclose(got["w0"][0], float.fromhex(g["w0"]), rtol=1e-8)
    np.testing.assert_allclose(got["w"], [float.fromhex(x) for x in g["w"]], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(got["V"].reshape(-1), [float.fromhex(x) for x in g["V"]], rtol=1e-8, atol=1e-11)
    desc = capi.make_desc(ts, capi.FM, SPEC["k"], reg_lw=float(np.float32(0.01)), reg_lf=float(np.float32(0.02)),
                          num_context_dims=len(SPEC["dims"]))
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(got)
        p = eng.predict(test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    assert np.max(np.abs(p - np.array([float.fromhex(x) for x in g["pred"]]))) < 1e-5  # the north star's tolerance
