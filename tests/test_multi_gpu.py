"""Two ranks over NCCL on two B200s: the sharded CUDA path against the single-process statement of the
sharded semantics (tests/test_sharding.block_jacobi_reference, CPU oracle).  Skipped on a one-GPU box."""
import os

import numpy as np
import pytest

from carskit_b200 import capi, recommender, sharding
from tests.test_sharding import _free_port, _problem, block_jacobi_reference

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, model_name, F, epochs, out_dir, combine):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ts, test, init = _problem(model_name, F)
        shard, lo = sharding.shard_training_set(ts, rank, world)
        hi = lo + shard.num_users
        conf = {"num.factors": str(F), "num.max.iter": str(epochs)}
        rec = recommender.getRecommender(model_name)(shard, sharding.shard_test_set(test, lo, hi), conf=conf,
                                                     device=rank, world=world, combine=combine)
        local = {k: (sharding.shard_user_rows(v, lo, hi) if k in ("P", "user_bias", "uc_bias") else v.copy())
                 for k, v in init.items()}
        rec.initModel(init=local)
        rec.keep_engine = True
        rec.buildModel()
        m = rec.evalRatings()
        rec.close_engine()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(rec.iter_losses), rmse=m["RMSE"], **rec.model)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model_name,combine", [("camf_ci", "mean"), ("camf_cu", "mean"), ("camf_cuci", "mean"), ("biasedmf", "sum")])
def test_two_gpus_match_block_jacobi_reference(oracle, cars_lib, tmp_path, model_name, combine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    F, epochs, world = 8, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), model_name, F, epochs, str(tmp_path), combine), nprocs=world, join=True)
    shards, locals_, item, losses = block_jacobi_reference(oracle, model_name, F, epochs, world, combine)
    for g in range(world):
        got = np.load(tmp_path / f"rank{g}.npz")
        for k, v in locals_[g].items():
            assert np.array_equal(got[k], v), k  # user side: trained locally, bit-identical
        for k, v in item.items():
            assert np.array_equal(got[k], v), k  # item block: old + (d0 + d1), order-independent for two ranks
        np.testing.assert_allclose(got["losses"], losses, rtol=1e-11)


def _fm_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle_py as orc
    from tests.test_fm_oracle import fm_inputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ts, test, prob, arrs = fm_inputs(orc, 400, 150, [4, 8], 40000, 8, seed=21, holdout=0.1)
        shard = sharding.shard_rows(ts, rank, world)
        rec = recommender.FM(shard, test, conf={"num.factors": "8", "num.max.iter": "3", "FM": "-lw 0.01 -lf 0.02"},
                             device=rank, world=world)
        rec.initModel(init=arrs)
        rec.keep_engine = True
        rec.buildModel()
        calls = rec.exchange.calls
        m = rec.evalRatings()
        rec.close_engine()
        np.savez(os.path.join(out_dir, f"fm{rank}.npz"), losses=np.array(rec.iter_losses), rmse=m["RMSE"], calls=calls,
                 **rec.model)
    finally:
        dist.destroy_process_group()


def test_fm_row_shards_match_the_single_process_sweep(oracle, cars_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from tests.test_fm_oracle import clone, fm_inputs
    mp.spawn(_fm_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ts, test, prob, arrs = fm_inputs(oracle, 400, 150, [4, 8], 40000, 8, seed=21, holdout=0.1)
    ref = clone(arrs)
    e, Q = oracle.fm_prepare(prob, ref)
    ref_losses = [oracle.fm_iteration(prob, ref, e, Q, closed_den=True) for _ in range(3)]
    a, b = np.load(tmp_path / "fm0.npz"), np.load(tmp_path / "fm1.npz")
    for name in ("w0", "w", "V"):
        assert np.array_equal(a[name], b[name]), name  # every rank computes the same coordinates
        np.testing.assert_allclose(a[name], ref[name], rtol=1e-9, atol=1e-12, err_msg=name)
    np.testing.assert_allclose(a["losses"], ref_losses, rtol=1e-10)
    assert int(a["calls"]) == 3 * (3 * (1 + 8) + 1) and a["rmse"] == b["rmse"]


# ---- ONE handle, N GPUs, one process (cars_desc.num_gpus; what a JVM can use) -----------------------------------------
@pytest.mark.parametrize("model_name,combine", [("camf_ci", "mean"), ("camf_cu", "mean"), ("camf_cuci", "mean"), ("biasedmf", "sum")])
def test_single_process_handle_over_two_gpus_matches_block_jacobi_reference(oracle, cars_lib, model_name, combine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tests.golden.make_golden import REGS
    F, epochs, world = 8, 3, 2
    model = capi.MODEL_NAMES[model_name]
    ts, test, init = _problem(model_name, F)
    got = {k: v.copy() for k, v in init.items()}
    desc = capi.make_desc(ts, model, F, gpu_ids=[0, 1], combine=capi.COMBINE_MEAN if combine == "mean" else capi.COMBINE_SUM, **REGS)
    lr, last, losses = capi.f32(0.02), 0.0, []
    with capi.Engine(desc, keepalive=ts) as eng:
        eng.upload(got)  # the caller's FULL arrays; user-side rows go to the GPU that owns them
        for it in range(1, epochs + 1):
            loss = eng.epoch(lr)
            losses.append(loss)
            if it > 1:
                lr = lr * 1.05 if abs(last) > abs(loss) else lr * 0.5
            last = loss
        pred = eng.predict(test["u"], test["j"], test.get("ctx"), bound=True, min_rate=1.0, max_rate=5.0)
        eng.download(got)
        st = eng.stats()
    shards, locals_, item, ref_losses = block_jacobi_reference(oracle, model_name, F, epochs, world, combine)
    for g, ((sh, lo), loc) in enumerate(zip(shards, locals_)):
        for k, v in loc.items():
            assert np.array_equal(got[k][lo:lo + sh.num_users], v), (g, k)
    for k, v in item.items():
        assert np.array_equal(got[k], v), k
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-11)
    assert st.num_gpus == 2 and st.nnz == ts.nnz and st.exchange_ms > 0
    # predictions routed by user equal the oracle's on the combined model
    full = {**item}
    for k in ("P", "user_bias", "uc_bias"):
        if k in got:
            full[k] = got[k]
    want = oracle.predict(capi.make_desc(ts, model, F, **REGS), full, test["u"], test["j"], test.get("ctx"), bound=True,
                          min_rate=1.0, max_rate=5.0)
    assert np.array_equal(pred, want)


def test_single_process_multi_gpu_touched_combine_and_fast_mode(oracle, cars_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from carskit_b200 import synth
    from tests.golden.make_golden import REGS, init_arrays
    # items split in two halves, each rated by the users of ONE shard only: with COMBINE_TOUCHED every item keeps the
    # full step of the one shard that trained it, i.e. the result equals two independent single-GPU runs
    ts, _ = synth.make_training_set(400, 200, [3, 2], 20000, seed=5)
    half = ts.u < 200
    ts.j[half] = ts.j[half] % 100
    ts.j[~half] = 100 + ts.j[~half] % 100
    key = np.lexsort((ts.ctx, ts.j, ts.u))
    ts = capi.TrainingSet(num_users=400, num_items=200, u=ts.u[key], j=ts.j[key], r=ts.r[key], ctx=ts.ctx[key],
                          num_conditions=ts.num_conditions, num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                          global_mean=ts.global_mean)
    keep = np.ones(ts.nnz, dtype=bool)
    keep[1:] = (ts.u[1:] != ts.u[:-1]) | (ts.j[1:] != ts.j[:-1]) | (ts.ctx[1:] != ts.ctx[:-1])
    ts = capi.TrainingSet(num_users=400, num_items=200, u=ts.u[keep], j=ts.j[keep], r=ts.r[keep], ctx=ts.ctx[keep],
                          num_conditions=ts.num_conditions, num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                          global_mean=ts.global_mean)
    init = init_arrays(oracle, capi.CAMF_CI, ts, 8, 3)
    got = {k: v.copy() for k, v in init.items()}
    with capi.Engine(capi.make_desc(ts, capi.CAMF_CI, 8, gpu_ids=[0, 1], combine=capi.COMBINE_TOUCHED, **REGS), keepalive=ts) as eng:
        eng.upload(got)
        losses = [eng.epoch(capi.f32(0.02)) for _ in range(2)]
        eng.download(got)
    ref = {k: v.copy() for k, v in init.items()}
    desc = capi.make_desc(ts, capi.CAMF_CI, 8, **REGS)
    ref_losses = [oracle.epoch(desc, ref, capi.f32(0.02)) for _ in range(2)]
    # disjoint item sets: the sharded run IS the serial run, up to the rounding of  old + (new - old)  in the exchange
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-10, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-10)
    # FAST mode on two GPUs (CAMF_C too: its shared condBias joins the item block)
    for model in (capi.CAMF_CI, capi.CAMF_C):
        arrs = init_arrays(oracle, model, ts, 8, 4)
        with capi.Engine(capi.make_desc(ts, model, 8, mode=capi.FAST, gpu_ids=[0, 1], **REGS), keepalive=ts) as eng:
            eng.upload(arrs)
            ls = [eng.epoch(capi.f32(0.02)) for _ in range(4)]
            eng.download(arrs)
        assert all(np.isfinite(ls)) and ls[-1] < ls[0]
        assert all(np.all(np.isfinite(v)) for v in arrs.values())


def test_plain_c_client_over_two_gpus(cars_lib, tmp_path):
    # examples/c_client.c --gpus 2: the single-process multi-GPU path from C99, no Python, no torch in the process
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "carskit_b200")
    exe = str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "c_client.c"), "-L", libdir, "-lcarskit_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    one = subprocess.run([exe], capture_output=True, text=True)
    two = subprocess.run([exe, "--gpus", "2"], capture_output=True, text=True)
    assert one.returncode == 0 and two.returncode == 0, one.stdout + two.stdout + two.stderr
    assert "gpus = 1" in one.stdout and "gpus = 2" in two.stdout
    l1 = [float(x.split("=")[1]) for x in one.stdout.splitlines() if x.startswith("iter ")]
    l2 = [float(x.split("=")[1]) for x in two.stdout.splitlines() if x.startswith("iter ")]
    assert len(l1) == len(l2) == 3 and all(np.isfinite(l2)) and l2[2] < l2[0]
    assert abs(l2[0] - l1[0]) < 0.05 * l1[0]  # same ratings, same initial model; the shards only see each other's items next epoch
