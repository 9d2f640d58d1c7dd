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


@pytest.mark.parametrize("model_name,combine", [("camf_ci", "mean"), ("camf_cu", "mean"), ("biasedmf", "sum")])
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
