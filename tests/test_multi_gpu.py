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
