"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs, and against the committed golden vectors.

Bar (DESIGN.md "Parity"): P, Q and every bias vector/matrix BIT-IDENTICAL to the oracle in EXACT mode
(np.array_equal); the epoch loss within 1e-11 relative (Java sums 67*nnz terms sequentially; the
engine sums per-lane partials); predictions bit-identical; RMSE bit-identical.
"""
import json
import math
import os

import numpy as np
import pytest

from carskit_b200 import capi, synth
from tests.golden.make_golden import REGS, bold_driver, digest, init_arrays, make_inputs

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-11
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sgd_golden.json")
CTX_MODELS = (capi.CAMF_C, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI)


def run_both(oracle, model, ts, F, epochs, seed, mode=capi.EXACT, lr0=capi.f32(0.02), schedule=capi.SCHED_FLAGGED, tuning=None):
    desc = capi.make_desc(ts, model, F, mode=mode, schedule=schedule, tuning=tuning, **REGS)
    ref = init_arrays(oracle, model, ts, F, seed)
    got = {k: v.copy() for k, v in ref.items()}
    ref_losses, got_losses = [], []
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)
        lr, last = lr0, 0.0
        for it in range(1, epochs + 1):
            lo = oracle.epoch(desc, ref, lr)
            lg = eng.epoch(lr)
            ref_losses.append(lo)
            got_losses.append(lg)
            lr = bold_driver(lr, last, lo, it)
            last = lo
        eng.download(got)
        st = eng.stats()
    return ref, got, ref_losses, got_losses, st


def assert_bit_identical(ref, got):
    for k in ref:
        assert np.array_equal(ref[k], got[k]), f"{k}: max abs diff {np.abs(ref[k] - got[k]).max()}"


@pytest.mark.parametrize("model", [capi.PMF, capi.BIASEDMF, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("F", [1, 7, 10, 16, 32, 64, 100, 128, 200])
def test_exact_mode_bit_identical(oracle, cars_lib, model, F):
    dims = [4, 3, 2] if model in CTX_MODELS else None
    ts, _ = synth.make_training_set(500, 120, dims, 20000, seed=F, order="user_sorted")
    ref, got, rl, gl, st = run_both(oracle, model, ts, F, epochs=3, seed=F + 1)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)
    assert st.num_levels > 1 and st.kernel_launches >= 6


@pytest.mark.parametrize("model", [capi.PMF, capi.BIASEDMF, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("F", [7, 64, 128])
@pytest.mark.parametrize("schedule", [capi.SCHED_WAVEFRONT, capi.SCHED_FLAGGED])
def test_level_schedules_bit_identical(oracle, cars_lib, model, F, schedule):
    dims = [4, 3, 2] if model in CTX_MODELS else None
    ts, _ = synth.make_training_set(500, 120, dims, 20000, seed=F, order="user_sorted")
    ref, got, rl, gl, st = run_both(oracle, model, ts, F, epochs=3, seed=F + 1, schedule=schedule)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


@pytest.mark.parametrize("schedule", [capi.SCHED_DATAFLOW, capi.SCHED_WAVEFRONT, capi.SCHED_FLAGGED])
@pytest.mark.parametrize("order", ["user_sorted", "shuffled"])
@pytest.mark.parametrize("zipf", [0.0, 1.0])
def test_exact_mode_orders_and_skew(oracle, cars_lib, order, zipf, schedule):
    ts, _ = synth.make_training_set(2000, 300, [8, 8, 8, 8], 60000, seed=9, order=order, item_zipf=zipf)
    ref, got, rl, gl, _ = run_both(oracle, capi.CAMF_CI, ts, 64, epochs=2, seed=3, schedule=schedule)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


@pytest.mark.parametrize("model", [capi.BIASEDMF, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("schedule", [capi.SCHED_DATAFLOW, capi.SCHED_FLAGGED])
def test_flag_schedules_large_random(oracle, cars_lib, model, schedule):
    # enough ratings to keep every resident group busy and to make cross-group dependencies frequent:
    # few items => long item chains that hop between groups on every rating
    dims = [8, 8, 8, 8] if model in CTX_MODELS else None
    for order, users, items in (("user_sorted", 20000, 64), ("shuffled", 3000, 2000)):
        ts, _ = synth.make_training_set(users, items, dims, 600000, seed=77, order=order)
        ref, got, rl, gl, _ = run_both(oracle, model, ts, 64, epochs=2, seed=5, schedule=schedule)
        assert_bit_identical(ref, got)
        np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


def test_dataflow_loss_is_deterministic(oracle, cars_lib):
    # chunks are handed out dynamically, but the loss is reduced per chunk in a fixed order
    ts, _ = synth.make_training_set(5000, 500, [8, 8], 200000, seed=13, order="shuffled")
    desc = capi.make_desc(ts, capi.CAMF_CI, 32, **REGS)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 32, 2)
    runs = []
    for _ in range(3):
        with capi.Engine(desc, keepalive=ts) as eng:
            eng.upload(arrs)
            runs.append([eng.epoch(0.02).hex() for _ in range(2)])
    assert runs[0] == runs[1] == runs[2]


@pytest.mark.parametrize("F,order", [(10, "shuffled"), (10, "user_sorted"), (1, "user_sorted"), (64, "shuffled"), (100, "user_sorted")])
def test_camf_c_exact_serial_kernel(oracle, cars_lib, F, order):
    # user_sorted order makes consecutive ratings share user, item (same pair, other context) and conditions
    ts, _ = synth.make_training_set(90, 120, [7, 7, 2, 3, 2, 9, 5, 4], 3000, seed=21, order=order)
    ref, got, rl, gl, st = run_both(oracle, capi.CAMF_C, ts, F, epochs=3, seed=5)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


def test_camf_c_ragged_contexts(oracle, cars_lib):
    # contexts of different lengths (and an empty one) put one condition id at different positions of the table
    rng = np.random.default_rng(3)
    ts, _ = synth.make_training_set(40, 30, [3, 4, 2], 1500, seed=8, order="user_sorted")
    ctxs = [[], [0], [3, 0], [1, 4, 7], [7], [8, 2], [5, 8, 1]]
    ptr = np.cumsum([0] + [len(c) for c in ctxs]).astype(np.int32)
    cond = np.array([x for c in ctxs for x in c], dtype=np.int32)
    ts2 = capi.TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=ts.u, j=ts.j, r=ts.r,
                           ctx=rng.integers(0, len(ctxs), ts.nnz).astype(np.int32), num_conditions=9,
                           num_contexts=len(ctxs), ctx_ptr=ptr, ctx_cond=cond, global_mean=ts.global_mean)
    ref, got, rl, gl, st = run_both(oracle, capi.CAMF_C, ts2, 10, epochs=3, seed=5)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


def test_many_context_dimensions(oracle, cars_lib):
    # more context dimensions (10) than lanes in a group (8): exercises the slow path
    ts, _ = synth.make_training_set(200, 50, [2] * 10, 5000, seed=4)
    for model in (capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI):
        ref, got, rl, gl, _ = run_both(oracle, model, ts, 10, epochs=2, seed=8)
        assert_bit_identical(ref, got)


@pytest.mark.parametrize("model", [capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("F,ndims", [(64, 5), (64, 7), (128, 6), (40, 9)])
def test_more_context_dimensions_than_lanes_per_rating(oracle, cars_lib, model, F, ndims):
    # the default flagged plan gives a rating 4 lanes at F in 33..64 (16 at 65..128): any data set with more context
    # dimensions than that takes the per-dimension slow path of compute_scatter on a MULTI-group kernel
    ts = synth.make_training_set(300, 80, ([2, 3] * 5)[:ndims], 12000, seed=ndims)[0]
    for schedule in (capi.SCHED_FLAGGED, capi.SCHED_DATAFLOW):
        ref, got, rl, gl, _ = run_both(oracle, model, ts, F, epochs=3, seed=8, schedule=schedule)
        assert_bit_identical(ref, got)
        np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


def test_edge_cases(oracle, cars_lib):
    # empty training set: an epoch is a no-op with loss 0
    ts, _ = synth.make_training_set(5, 4, [2], 0, seed=1)
    desc = capi.make_desc(ts, capi.CAMF_CI, 8, **REGS)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 8, 1)
    keep = {k: v.copy() for k, v in arrs.items()}
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        assert eng.epoch(0.02) == 0.0
        eng.download(arrs)
    assert_bit_identical(keep, arrs)
    # one rating; one user rating every item (a pure chain: every level has one rating)
    for users, items, nnz in ((1, 1, 1), (1, 50, 200), (50, 1, 200)):
        ts, _ = synth.make_training_set(users, items, [3, 3], nnz, seed=2)
        ref, got, rl, gl, st = run_both(oracle, capi.CAMF_CI, ts, 12, epochs=2, seed=3)
        assert_bit_identical(ref, got)
        np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


def test_nan_loss_is_returned_not_fatal(oracle, cars_lib):
    # IterativeRecommender.java:181-184 decides what to do with NaN; the engine must just report it
    ts, _ = synth.make_training_set(30, 20, [2, 2], 500, seed=3)
    desc = capi.make_desc(ts, capi.CAMF_CI, 8, **REGS)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 8, 1)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        losses = [eng.epoch(1e6) for _ in range(6)]
    assert any(math.isnan(x) or math.isinf(x) for x in losses)


def test_call_order_errors(oracle, cars_lib):
    ts, _ = synth.make_training_set(30, 20, [2, 2], 500, seed=3)
    desc = capi.make_desc(ts, capi.CAMF_CI, 8, **REGS)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 8, 1)
    with capi.Engine(desc, keepalive=ts) as eng:
        with pytest.raises(capi.CarsError) as e:
            eng.epoch(0.02)
        assert e.value.code == -6
        with pytest.raises(capi.CarsError):
            eng.upload({"P": arrs["P"], "Q": arrs["Q"]})  # user_bias / ic_bias missing
        eng.upload(arrs)
        eng.epoch_begin(0.02)
        with pytest.raises(capi.CarsError):
            eng.epoch_begin(0.02)
        eng.epoch_wait()
        with pytest.raises(capi.CarsError):
            eng.predict([999], [0], [0])
    bad = capi.make_desc(ts, capi.CAMF_CI, 8, **REGS)
    ts.u[7] = 10 ** 6   # ids are range-checked on the device; the message names the FIRST offending rating
    ts.j[3] = -1
    ts.ctx[400] = ts.num_contexts
    with pytest.raises(capi.CarsError) as e:
        capi.Engine(bad, keepalive=ts)
    assert e.value.code == -1 and "rating 3 " in str(e.value)


@pytest.mark.parametrize("model", [capi.PMF, capi.BIASEDMF, capi.CAMF_C, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
def test_predict_and_eval_bit_identical(oracle, cars_lib, model):
    dims = [3, 4] if model in CTX_MODELS else None
    ts, test = synth.make_training_set(300, 90, dims, 9000, seed=12, holdout=0.2)
    F = 20
    desc = capi.make_desc(ts, model, F, **REGS)
    arrs = init_arrays(oracle, model, ts, F, 4)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        for bound in (False, True):
            p = eng.predict(test["u"], test["j"], test["ctx"], bound=bound, min_rate=1.0, max_rate=5.0)
            q = oracle.predict(desc, arrs, test["u"], test["j"], test["ctx"], bound=bound, min_rate=1.0, max_rate=5.0)
            assert np.array_equal(p, q)
        sa, ss = eng.eval_ratings(test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
        sa2, ss2, cnt = oracle.eval_ratings(desc, arrs, test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
        assert (sa, ss) == (sa2, ss2)


def _golden_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_engine_reproduces_golden_vectors(oracle, cars_lib, case):
    spec = case["spec"]
    model, ts, test, desc, arrs = make_inputs(oracle, spec)
    want_losses = [float.fromhex(x) for x in case["losses_hex"]]
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        lr, last, losses = capi.f32(0.02), 0.0, []
        for it in range(1, spec["epochs"] + 1):
            loss = eng.epoch(lr)
            losses.append(loss)
            # drive the schedule with the GOLDEN loss so a last-bit loss difference cannot fork it
            lr = bold_driver(lr, last, want_losses[it - 1], it)
            last = want_losses[it - 1]
        eng.download(arrs)
        sa, ss = eng.eval_ratings(test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
    np.testing.assert_allclose(losses, want_losses, rtol=LOSS_RTOL, atol=0)
    assert digest(arrs) == case["digest"]  # every trained array byte-identical to the oracle's
    assert math.sqrt(ss / len(test["r"])).hex() == case["rmse_hex"]


def test_two_handles_concurrently(oracle, cars_lib):
    # CARSKit runs cross-validation folds on parallel threads (CARSKit.java:395-412): handles must be
    # independent.  Interleave epochs of two handles and compare each with its own oracle run.
    ts1, _ = synth.make_training_set(200, 60, [3, 3], 4000, seed=31)
    ts2, _ = synth.make_training_set(150, 80, [2, 5], 5000, seed=32)
    d1 = capi.make_desc(ts1, capi.CAMF_CI, 16, **REGS)
    d2 = capi.make_desc(ts2, capi.CAMF_CU, 24, **REGS)
    r1, r2 = init_arrays(oracle, capi.CAMF_CI, ts1, 16, 1), init_arrays(oracle, capi.CAMF_CU, ts2, 24, 2)
    g1, g2 = {k: v.copy() for k, v in r1.items()}, {k: v.copy() for k, v in r2.items()}
    with capi.Engine(d1, keepalive=ts1) as e1, capi.Engine(d2, keepalive=ts2) as e2:
        e1.upload(g1)
        e2.upload(g2)
        for _ in range(3):
            e1.epoch_begin(0.02)
            e2.epoch_begin(0.01)
            e2.epoch_wait()
            e1.epoch_wait()
            oracle.epoch(d1, r1, 0.02)
            oracle.epoch(d2, r2, 0.01)
        e1.download(g1)
        e2.download(g2)
    assert_bit_identical(r1, g1)
    assert_bit_identical(r2, g2)


# ---------------------------------------------------------------------------------------------------
# the schedule is built on the device (schedule_gpu.cuh): same levels as the sequential host pass
# ---------------------------------------------------------------------------------------------------
def host_levels(ts):
    lu, lj = np.zeros(ts.num_users, np.int64), np.zeros(ts.num_items, np.int64)
    sizes = {}
    for u, j in zip(ts.u.tolist(), ts.j.tolist()):
        l = 1 + max(lu[u], lj[j])
        lu[u] = lj[j] = l
        sizes[l] = sizes.get(l, 0) + 1
    return (max(sizes), max(sizes.values())) if sizes else (0, 0)


@pytest.mark.parametrize("order,users,items,nnz,zipf", [("user_sorted", 500, 120, 20000, 0.0), ("shuffled", 3000, 2000, 100000, 1.0),
                                                         ("shuffled", 1, 300, 300, 0.0), ("user_sorted", 4000, 3, 9000, 0.0)])
def test_device_built_levels_match_the_sequential_pass(oracle, cars_lib, order, users, items, nnz, zipf):
    ts, _ = synth.make_training_set(users, items, [4, 3], nnz, seed=11, order=order, item_zipf=zipf)
    want = host_levels(ts)
    arrs = init_arrays(oracle, capi.CAMF_CI, ts, 8, 1)
    out = {}
    for how in ("device", "host"):
        desc = capi.make_desc(ts, capi.CAMF_CI, 8, tuning=f"levels={how}", **REGS)
        got = {k: v.copy() for k, v in arrs.items()}
        with capi.Engine(desc, keepalive=ts) as eng:
            st = eng.stats()
            assert (st.num_levels, st.max_level_size) == want, how
            eng.upload(got)
            loss = eng.epoch(0.02)
            eng.download(got)
        out[how] = (loss.hex(), digest(got))
    assert out["device"] == out["host"]  # same record stream => same loss bits, same model


def test_large_pageable_and_pinned_transfers(oracle, cars_lib):
    # arrays well above the staging chunk (4 MiB): several host threads copy through pinned double buffers;
    # pinned inputs (torch pin_memory) are DMA'd directly.  Both must be byte-exact in both directions.
    import torch
    ts, _ = synth.make_training_set(300_000, 2_000, [4, 4], 3_000_000, seed=5, order="user_sorted")
    F = 16
    desc = capi.make_desc(ts, capi.CAMF_CU, F, **REGS)
    arrs = init_arrays(oracle, capi.CAMF_CU, ts, F, 2)
    assert arrs["P"].nbytes > 8 * (4 << 20)
    ref = {k: v.copy() for k, v in arrs.items()}
    want_loss = oracle.epoch(desc, ref, capi.f32(0.02))

    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()

    for pin in (False, True):
        keep = []
        if pin:
            fields = {}
            for name in ("u", "j", "ctx", "r"):
                t, v = pinned(getattr(ts, name))
                keep.append(t)
                fields[name] = v
            ts2 = capi.TrainingSet(num_users=ts.num_users, num_items=ts.num_items, num_conditions=ts.num_conditions,
                                   num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                                   global_mean=ts.global_mean, **fields)
            got = {}
            for k, v in arrs.items():
                t, a = pinned(v)
                keep.append(t)
                got[k] = a
        else:
            ts2, got = ts, {k: v.copy() for k, v in arrs.items()}
        d2 = capi.make_desc(ts2, capi.CAMF_CU, F, **REGS)
        with capi.Engine(d2, keepalive=ts2) as eng:
            eng.upload(got)
            back = {k: np.zeros_like(v) for k, v in got.items()}
            eng.download(back)
            assert_bit_identical(arrs, back)  # round trip without an epoch
            loss = eng.epoch(capi.f32(0.02))
            eng.download(got)
        assert_bit_identical(ref, got)
        np.testing.assert_allclose(loss, want_loss, rtol=LOSS_RTOL, atol=0)


def test_config2_frappe_shaped_camf_c_bit_identical(oracle, cars_lib):
    # BASELINE.json configs[1] at full size (bench.py --workload camf_c_f10_frappe_shaped times the same shape): CAMF_C, 10 factors, 957 users x
    # 4 082 items, 8 context dimensions with 7/7/2/3/2/9/80/233 conditions, 96 203 drawn ratings, 90/10 split.
    # EXACT mode (one warp, reference order) must reproduce the oracle bit for bit, predictions and RMSE included.
    ts, test = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1, holdout=0.1)
    ref, got, rl, gl, st = run_both(oracle, capi.CAMF_C, ts, 10, epochs=5, seed=1)
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)
    desc = capi.make_desc(ts, capi.CAMF_C, 10, **REGS)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)
        sa, ss = eng.eval_ratings(test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
    ra, rs, cnt = oracle.eval_ratings(desc, ref, test["u"], test["j"], test["ctx"], test["r"], 1.0, 5.0)
    assert (sa, ss) == (ra, rs) and cnt == len(test["r"])
    assert ts.num_conditions == 343 and st.nnz == ts.nnz


# ---- against OUTPUTS OF THE REFERENCE ITSELF -------------------------------------------------------------------------
# tests/golden/jvm_golden.json was produced by executing the reference's own class files (tests/tools/minijvm.py,
# tests/golden/make_jvm_golden.py) in the container that has /root/reference; the GPU box does not, the vectors travel.
from tests.golden import make_jvm_golden as JG  # noqa: E402

JVM_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "jvm_golden.json")))


@pytest.mark.parametrize("case", JVM_GOLDEN["cases"], ids=[c["name"] for c in JVM_GOLDEN["cases"]])
def test_engine_reproduces_the_reference_bytecode_vectors(oracle, cars_lib, case):
    from carskit_b200 import recommender
    spec = case["spec"]
    model, ts, test, arrs = JG.sgd_inputs(oracle, spec)
    assert JG.sha(ts.u, ts.j, ts.ctx, ts.r, *[arrs[k] for k in sorted(arrs)]) == case["input_sha"], "input generator drift"
    hy = JG.hyper(spec)
    lr = f"{hy['lrate']} -max {hy['max_lrate']}" + (" -bold-driver" if hy["bold_driver"] else "") + \
        (f" -decay {hy['decay']}" if hy["decay"] > 0 else "")
    conf = {"num.factors": str(spec["F"]), "num.max.iter": str(spec["iters"]), "learn.rate": lr, "reg.lambda": "0.0001 -c 0.001",
            "rating.min": "1", "rating.max": "5"}
    rec = recommender.getRecommender(spec["model"])(ts, test, conf=conf)
    rec.initModel(init=arrs)
    rec.keep_engine = True
    rec.buildModel()  # the host mirror of isConverged / updateLRate drives cars_epoch
    try:
        assert digest(rec.model) == case["digest"]  # P, Q, biases: bit-identical to what the Java class files computed
        want = [float.fromhex(x) for x in case["losses_hex"]]
        np.testing.assert_allclose(rec.iter_losses, want, rtol=LOSS_RTOL, atol=0)
        pred = rec.predict(test["u"], test["j"], test["ctx"], bound=True)
        assert [float(x).hex() for x in pred[:40]] == case["pred_hex"]
    finally:
        rec.close_engine()


def test_fm_engine_within_1e5_of_the_reference_bytecode_vector(oracle, cars_lib):
    case = JVM_GOLDEN["fm"]
    spec = case["spec"]
    ts, test, arrs = JG.fm_inputs(oracle, spec)
    assert JG.sha(ts.u, ts.j, ts.ctx, ts.r, arrs["w"], arrs["V"]) == case["input_sha"], "input generator drift"
    desc = capi.make_desc(ts, capi.FM, spec["k"], reg_lw=capi.f32(spec["reg_lw"]), reg_lf=capi.f32(spec["reg_lf"]),
                          num_context_dims=len(spec["dims"]))
    with capi.FmEngine(desc, keepalive=ts) as eng:
        eng.upload(arrs)
        eng.prepare()
        for _ in range(spec["iters"]):
            eng.iteration()
        pred = eng.predict(test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
    want = np.array([float.fromhex(x) for x in case["pred_hex"]])
    assert np.max(np.abs(pred - want)) < 1e-5  # the north star's bar for predicted ratings (sums in another order)


# ---- K1t: the flagged schedule on tagged rows (csrc/tagged_kernels.cuh, tuning "tagged=1") -------------------------------
@pytest.mark.parametrize("model", [capi.PMF, capi.BIASEDMF, capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI])
@pytest.mark.parametrize("F", [1, 7, 10, 15, 16, 32, 64, 100, 128])
@pytest.mark.parametrize("ctas", [2, 3])
def test_tagged_rows_bit_identical(oracle, cars_lib, model, F, ctas):
    dims = [4, 3, 2] if model in CTX_MODELS else None
    ts, test = synth.make_training_set(500, 120, dims, 20000, seed=F, order="user_sorted", holdout=0.1)
    ref, got, rl, gl, st = run_both(oracle, model, ts, F, epochs=3, seed=F + 1, tuning=f"tagged=1;tagged_ctas={ctas}")
    assert_bit_identical(ref, got)
    np.testing.assert_allclose(gl, rl, rtol=LOSS_RTOL, atol=0)


@pytest.mark.parametrize("order,users,items,nnz,zipf,dims", [("shuffled", 3000, 2000, 100000, 1.0, [8, 8, 8, 8]), ("user_sorted", 1, 300, 300, 0.0, [3, 3]),
                                                             ("user_sorted", 4000, 3, 9000, 0.0, [2] * 10), ("shuffled", 20000, 5000, 600000, 0.0, [8, 8, 8, 8])])
def test_tagged_rows_on_chains_skew_and_many_dimensions(oracle, cars_lib, order, users, items, nnz, zipf, dims):
    ts, test = synth.make_training_set(users, items, dims, nnz, seed=3, order=order, item_zipf=zipf, holdout=0.05)
    for model in (capi.CAMF_CI, capi.CAMF_CU):
        F = 64 if sum(dims) <= 32 else 16
        desc = capi.make_desc(ts, model, F, tuning="tagged=1", **REGS)
        ref = init_arrays(oracle, model, ts, F, 9)
        got = {k: v.copy() for k, v in ref.items()}
        with capi.Engine(desc, keepalive=ts) as eng:
            eng.upload(got)
            for it in range(4):  # several epochs through the same handle: every row's tag must be back at 0 each time
                lg = eng.epoch(capi.f32(0.02))
                lo = oracle.epoch(desc, ref, capi.f32(0.02))
                np.testing.assert_allclose(lg, lo, rtol=LOSS_RTOL)
                if it == 1:  # predictions in the middle of training come from the standard layout (converted on demand)
                    p = eng.predict(test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
                    assert np.array_equal(p, oracle.predict(desc, ref, test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0))
            eng.download(got)
        assert_bit_identical(ref, got)


# ---- N4: similarity-based CAMF (CAMF_ICS.java) ----------------------------------------------------------------------------
@pytest.mark.parametrize("F,order", [(10, "user_sorted"), (64, "shuffled"), (100, "user_sorted")])
def test_camf_ics_bit_identical(oracle, cars_lib, F, order):
    ts, test = synth.make_training_set(90, 120, [4, 3, 2, 5], 3000, seed=F, order=order, holdout=0.1)
    desc = capi.make_desc(ts, capi.CAMF_ICS, F, **REGS)
    ref = init_arrays(oracle, capi.CAMF_ICS, ts, F, 3)
    for k in ("P", "Q"):
        ref[k] *= 2.0 / np.sqrt(F)  # keep P[u].Q[j] in the rating range for any F
    got = {k: v.copy() for k, v in ref.items()}
    lr = capi.f32(0.005)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)
        for _ in range(3):
            lg, lo = eng.epoch(lr), oracle.epoch(desc, ref, lr)
            np.testing.assert_allclose(lg, lo, rtol=LOSS_RTOL)
        pred = eng.predict(test["u"], test["j"], test["ctx"], bound=False)
        eng.download(got)
    assert_bit_identical(ref, got)
    assert np.any(ref["cc_sim"] != 1.0) and np.array_equal(ref["cc_sim"], ref["cc_sim"].T)
    assert np.array_equal(pred, oracle.predict(desc, ref, test["u"], test["j"], test["ctx"]))
    with pytest.raises(capi.CarsError):
        capi.Engine(capi.make_desc(ts, capi.CAMF_ICS, F, mode=capi.FAST, **REGS), keepalive=ts)


@pytest.mark.parametrize("name,F,order,numF", [("camf_lcs", 10, "user_sorted", 10), ("camf_lcs", 64, "shuffled", 7), ("camf_lcs", 100, "user_sorted", 3),
                                               ("camf_mcs", 10, "shuffled", 0), ("camf_mcs", 64, "user_sorted", 0), ("camf_mcs", 130, "shuffled", 0)])
def test_camf_lcs_mcs_bit_identical(oracle, cars_lib, name, F, order, numF):
    """CAMF_LCS / CAMF_MCS (sim/CAMF_LCS.java, sim/CAMF_MCS.java) on the one-warp serial kernel: P, Q and the condition vectors /
    positions bit-identical to the oracle (itself bit-identical to the executed bytecode), bounded and unbounded predictions too."""
    model = capi.MODEL_NAMES[name]
    ts, test = synth.make_training_set(90, 120, [4, 3, 2, 5], 3000, seed=F, order=order, holdout=0.1)
    kw = dict(num_context_factors=numF) if numF else {}
    desc = capi.make_desc(ts, model, F, **REGS, **kw)
    shapes = capi.member_shapes(model, ts.num_users, ts.num_items, ts.num_conditions, F, numF or 10)
    g = oracle.JavaRandom(F + 3)
    ref = {k: g.uniform(shp) for k, shp in shapes.items()}
    for k in ("P", "Q"):
        ref[k] *= 2.0 / np.sqrt(F)  # keep P[u].Q[j] in the rating range for any F
    if "cf_lcs" in ref:
        ref["cf_lcs"] *= 2.0 / np.sqrt(numF)  # similarities of ~1
    if "c_mcs" in ref:
        ref["c_mcs"] /= np.sqrt(4)  # CAMF_MCS.java:44-48: U(0, upbound)
    got = {k: v.copy() for k, v in ref.items()}
    lr = capi.f32(0.002)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)
        for _ in range(3):
            lg, lo = eng.epoch(lr), oracle.epoch(desc, ref, lr)
            np.testing.assert_allclose(lg, lo, rtol=LOSS_RTOL)
        pred = eng.predict(test["u"], test["j"], test["ctx"], bound=False)
        predb = eng.predict(test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0)
        eng.download(got)
    assert_bit_identical(ref, got)
    assert np.array_equal(pred, oracle.predict(desc, ref, test["u"], test["j"], test["ctx"]))
    assert np.array_equal(predb, oracle.predict(desc, ref, test["u"], test["j"], test["ctx"], bound=True, min_rate=1.0, max_rate=5.0))
    with pytest.raises(capi.CarsError):
        capi.Engine(capi.make_desc(ts, model, F, mode=capi.FAST, **REGS, **kw), keepalive=ts)


@pytest.mark.parametrize("F,order", [(10, "user_sorted"), (7, "shuffled"), (64, "user_sorted"), (130, "shuffled")])
def test_svdpp_bit_identical(oracle, cars_lib, F, order):
    """SVD++ (baseline/cf/SVDPlusPlus.java) on the one-warp serial kernel: P, Q, both biases and Y bit-identical to the oracle
    (itself bit-identical to the executed bytecode) -- a user's items are taken ascending (train.getColumns(u)) whatever the
    order of the rating stream; users with 1 .. 60 items, so the 32-items-at-a-time prediction loop takes both paths."""
    ts, test = synth.make_training_set(60, 90, None, 2400, seed=F, order=order, holdout=0.1)
    test["ctx"] = None
    desc = capi.make_desc(ts, capi.SVDPP, F, **REGS)
    ref = init_arrays(oracle, capi.SVDPP, ts, F, 3)
    got = {k: v.copy() for k, v in ref.items()}
    lr = capi.f32(0.01)
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)
        for _ in range(3):
            lg, lo = eng.epoch(lr), oracle.epoch(desc, ref, lr)
            np.testing.assert_allclose(lg, lo, rtol=LOSS_RTOL)
        pred = eng.predict(test["u"], test["j"], None, bound=True, min_rate=1.0, max_rate=5.0)
        eng.download(got)
    assert_bit_identical(ref, got)
    assert np.bincount(ts.u, minlength=60).max() > 32
    assert np.array_equal(pred, oracle.predict(desc, ref, test["u"], test["j"], None, bound=True, min_rate=1.0, max_rate=5.0))
    with pytest.raises(capi.CarsError):
        capi.Engine(capi.make_desc(ts, capi.SVDPP, F, mode=capi.FAST, **REGS), keepalive=ts)
