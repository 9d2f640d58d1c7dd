"""FAST (hogwild) mode, CARS_FAST: csrc/fast_kernels.cuh.  Not serial-equivalent, so the bars differ from EXACT's:

  * with no two ratings sharing an item no update races -> the result must equal the serial loop up to the order
    of the dot product (a tree in the kernel, f = 0..F-1 in DenseMatrix.rowMult): 1e-11 relative;
  * on data that can be learnt (planted low-rank ratings) the held-out RMSE after the same number of epochs must
    be within RMSE_TOL of the serial oracle's, uniform AND Zipf(1.0) items (where EXACT has nothing to run in
    parallel), and for CAMF_C, whose condBias vector every rating shares;
  * loss finite and decreasing; the damping factors are reported.
"""
import math

import numpy as np
import pytest

from carskit_b200 import capi, synth
from tests.golden.make_golden import REGS, bold_driver, init_arrays

pytestmark = pytest.mark.gpu

RMSE_TOL = 0.02  # held-out RMSE(FAST) may exceed RMSE(serial oracle) by at most this on the planted data sets below
                 # (RMSE ~ 0.7-1.0).  One-sided: the damped hot rows act as a regulariser and FAST is often BETTER
                 # than the serial loop on held-out data (measured: -0.014 and -0.025 on the Zipf cases)
CTX_MODELS = (capi.CAMF_C, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI)


def train_both(oracle, model, ts, F, epochs, seed, **desc_kw):
    ref = init_arrays(oracle, model, ts, F, seed)
    got = {k: v.copy() for k, v in ref.items()}
    d_ref = capi.make_desc(ts, model, F, **REGS)
    d_fast = capi.make_desc(ts, model, F, mode=capi.FAST, **REGS, **desc_kw)
    rl, gl = [], []
    with capi.Engine(d_fast, keepalive=ts) as eng:
        eng.upload(got)
        lr_r = lr_g = capi.f32(0.02)
        last_r = last_g = 0.0
        for it in range(1, epochs + 1):
            lo = oracle.epoch(d_ref, ref, lr_r)
            lg = eng.epoch(lr_g)
            rl.append(lo)
            gl.append(lg)
            lr_r, last_r = bold_driver(lr_r, last_r, lo, it), lo
            lr_g, last_g = bold_driver(lr_g, last_g, lg, it), lg
        eng.download(got)
        st = eng.stats()
    return ref, got, rl, gl, st


def rmse(oracle, model, ts, F, arrs, test):
    d = capi.make_desc(ts, model, F, **REGS)
    p = oracle.predict(d, arrs, test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    return math.sqrt(float(np.mean((test["r"] - p) ** 2)))


@pytest.mark.parametrize("model", [capi.PMF, capi.BIASEDMF, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("F,shape", [(10, 0), (31, 0), (64, 0), (64, 5), (100, 0), (100, 5), (128, 0), (200, 0)])
def test_fast_equals_serial_when_no_two_ratings_share_an_item(oracle, cars_lib, model, F, shape):
    # shape 0: the default -- the Q-row step goes through ONE TMA add-reduce of a staged row for Fp >= 32 (UBLKRED), scalar
    # red.global.add.f64 below that; shape 5: the scalar kernels at every F
    # every item is rated exactly once: the only chains are the users', and FAST keeps those in order
    users, items = 150, 6000
    rng = np.random.default_rng(F)
    u = np.sort(rng.integers(0, users, size=items)).astype(np.int32)
    j = rng.permutation(items).astype(np.int32)
    dims = [4, 3, 2] if model in CTX_MODELS else None
    base, _ = synth.make_training_set(users, items, dims, 10, seed=1)
    ctx = rng.integers(0, base.num_contexts, size=items).astype(np.int32) if dims else None
    r = rng.integers(1, 6, size=items).astype(np.float64)
    ts = capi.TrainingSet(num_users=users, num_items=items, u=u, j=j, r=r, ctx=ctx, num_conditions=base.num_conditions,
                          num_contexts=base.num_contexts, ctx_ptr=base.ctx_ptr, ctx_cond=base.ctx_cond,
                          global_mean=float(r.mean()))
    ref, got, rl, gl, st = train_both(oracle, model, ts, F, epochs=3, seed=3, tuning=f"shape={shape}")
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-11, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(gl, rl, rtol=1e-10)
    assert st.fast_min_item_scale == 1.0


@pytest.mark.parametrize("model,F,zipf", [(capi.CAMF_CI, 64, 0.0), (capi.CAMF_CI, 64, 1.0), (capi.CAMF_CU, 128, 1.0),
                                          (capi.BIASEDMF, 10, 1.0), (capi.CAMF_CUCI, 16, 0.0), (capi.PMF, 32, 1.0)])
def test_fast_converges_like_the_serial_loop(oracle, cars_lib, model, F, zipf):
    dims = [4, 4, 2] if model in CTX_MODELS else None
    ts, test = synth.make_training_set(20000, 3000, dims, 600000, seed=5, item_zipf=zipf, holdout=0.1, planted_rank=4)
    if dims is None:
        test["ctx"] = None
    ref, got, rl, gl, st = train_both(oracle, model, ts, F, epochs=8, seed=7)
    assert all(math.isfinite(x) for x in gl)
    assert gl[-1] < gl[0]
    r_ref, r_got = rmse(oracle, model, ts, F, ref, test), rmse(oracle, model, ts, F, got, test)
    print(f"model {model} F {F} zipf {zipf}: RMSE serial {r_ref:.4f} fast {r_got:.4f}; max item degree {st.max_item_degree}, "
          f"min item scale {st.fast_min_item_scale:.3g}; loss serial {rl[-1]:.6g} fast {gl[-1]:.6g}")
    # Zipf(1.0) on 600 K ratings is the worst case for FAST: a third of all ratings sit on rows whose step is damped to a
    # few per cent, so a model whose signal is mostly in those rows (BiasedMF, 10 factors) converges more slowly per
    # epoch than the serial loop (measured +0.040 after 8 epochs); the other Zipf cases end BETTER than the serial loop
    assert r_got - r_ref < (0.06 if zipf > 0 else RMSE_TOL)
    assert r_got < float(np.std(test["r"]))  # and it has learnt something: better than predicting the mean
    if zipf > 0:
        assert st.fast_min_item_scale < 1.0 and st.fast_hot_rows > 0  # the Zipf head is damped and CTA-accumulated
    else:
        assert st.fast_hot_rows == 0


def test_fast_camf_c_shared_condition_biases(oracle, cars_lib):
    # BASELINE configs[1] shape: every rating reads and writes condBias; EXACT is one chain on one warp
    ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1, planted_rank=3)
    ref, got, rl, gl, st = train_both(oracle, capi.CAMF_C, ts, 10, epochs=30, seed=2)
    assert all(math.isfinite(x) for x in gl)
    r_ref, r_got = rmse(oracle, capi.CAMF_C, ts, 10, ref, test), rmse(oracle, capi.CAMF_C, ts, 10, got, test)
    print(f"CAMF_C Frappe-shaped: RMSE serial {r_ref:.4f} fast {r_got:.4f}; min cond scale {st.fast_min_cond_scale:.3g}")
    assert st.fast_min_cond_scale < 1.0
    assert r_got - r_ref < 0.05


def test_fast_edge_cases(oracle, cars_lib):
    # empty training set
    ts, _ = synth.make_training_set(5, 4, [2], 0, seed=1)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 8, 1)
    keep = {k: v.copy() for k, v in arrs.items()}
    with capi.Engine(capi.make_desc(ts, capi.CAMF_CI, 8, mode=capi.FAST, **REGS), keepalive=ts) as eng:
        eng.upload(arrs)
        assert eng.epoch(0.02) == 0.0
        eng.download(arrs)
    for k in keep:
        assert np.array_equal(keep[k], arrs[k])
    # one user (a single chunk: serial, equals the oracle up to the dot-product order); one item (every rating races on it)
    for users, items, nnz in ((1, 50, 200), (300, 1, 300), (1, 1, 1)):
        ts, _ = synth.make_training_set(users, items, [3, 3], nnz, seed=2)
        ref, got, rl, gl, st = train_both(oracle, capi.CAMF_CI, ts, 16, epochs=2, seed=4)
        assert all(math.isfinite(x) for x in gl)
        for k in got:
            assert np.all(np.isfinite(got[k])), k
    # more context dimensions than lanes in a group (slow path), user-side and item-side cells
    ts, _ = synth.make_training_set(200, 5000, [2] * 10, 5000, seed=4)
    for model in (capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI, capi.CAMF_C):
        # never damp (-1) where the shared rows are items with about one rating each; CAMF_C's condBias cells are shared
        # by half of all ratings and NEED the damping (undamped, 10 cells x 16 concurrent updates diverge)
        kw = {} if model == capi.CAMF_C else {"fast_max_conc": -1.0}
        ref, got, rl, gl, st = train_both(oracle, model, ts, 10, epochs=2, seed=8, **kw)
        assert all(math.isfinite(x) for x in gl), model
        if model != capi.CAMF_C:  # about one rating per item: few races, the first epoch's loss is the serial loop's
            np.testing.assert_allclose(gl[0], rl[0], rtol=5e-2)
        else:                     # CAMF_C's damped condBias cells learn more slowly inside the epoch: same magnitude only
            assert 0.8 * rl[0] < gl[0] < 1.5 * rl[0]


def test_fast_rejects_bad_ids(cars_lib):
    ts, _ = synth.make_training_set(50, 40, [2, 2], 500, seed=1)
    ts.j[123] = 40
    with pytest.raises(capi.CarsError) as e:
        capi.Engine(capi.make_desc(ts, capi.CAMF_CI, 8, mode=capi.FAST, **REGS), keepalive=ts)
    assert e.value.code == -1 and "rating 123" in str(e.value)
