"""N > 1 path on CPU: user-range sharding and the per-epoch item-block exchange, world_size 2 over gloo.

The CUDA engine cannot run here, so a stand-in engine backed by the CPU oracle implements the three
sharded entry points (item_block_doubles / epoch_sharded_begin / epoch_sharded_finish) with the same
contract as include/carskit_b200.h; what is under test is the host logic that a GPU run uses unchanged:
sharding.py (ranges, shards, ItemBlockExchange) and recommender.train_epoch()/evalRatings() at world > 1.
"""
import ctypes
import os
import socket

import numpy as np
import pytest

from carskit_b200 import capi, recommender, sharding, synth

REGS = dict(reg_u=capi.f32(1e-4), reg_i=capi.f32(1e-4), reg_b=capi.f32(1e-4), reg_c=capi.f32(1e-3))
ITEM_MEMBERS = ("Q", "item_bias", "ic_bias")


def test_user_ranges_partition():
    for users, world in ((10, 1), (10, 3), (7, 8), (1_000_000, 8), (97, 4)):
        r = [sharding.user_range(users, g, world) for g in range(world)]
        assert r[0][0] == 0 and r[-1][1] == users
        assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.user_range(10, 3, 3)


def test_shards_keep_reference_order_and_cover_everything():
    ts, test = synth.make_training_set(101, 37, [3, 4], 5000, seed=3, order="shuffled", holdout=0.1)
    seen = 0
    for g in range(4):
        sh, lo = sharding.shard_training_set(ts, g, 4)
        lo2, hi = sharding.user_range(ts.num_users, g, 4)
        assert lo == lo2 and sh.num_users == hi - lo
        idx = np.nonzero((ts.u >= lo) & (ts.u < hi))[0]
        assert np.array_equal(sh.u + lo, ts.u[idx]) and np.array_equal(sh.j, ts.j[idx])
        assert np.array_equal(sh.ctx, ts.ctx[idx]) and np.array_equal(sh.r, ts.r[idx])
        assert sh.global_mean == ts.global_mean and sh.num_items == ts.num_items
        t = sharding.shard_test_set(test, lo, hi)
        assert np.all((t["u"] >= 0) & (t["u"] < sh.num_users))
        seen += sh.nnz
    assert seen == ts.nnz


class OracleEngine:
    """Stand-in for capi.Engine on a machine without a GPU (tests only)."""

    def __init__(self, desc, model, arrs, oracle):
        self.desc, self.model, self.arrs, self.oracle = desc, model, arrs, oracle
        self.members = [k for k in ITEM_MEMBERS if k in arrs]
        self.old = None
        self.loss = None

    def item_block_doubles(self):
        return int(sum(self.arrs[k].size for k in self.members))

    def _view(self, ptr):
        n = self.item_block_doubles()
        return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))

    def _pack(self):
        return np.concatenate([self.arrs[k].reshape(-1) for k in self.members])

    def epoch_sharded_begin(self, lrate, ptr):
        self.old = self._pack()
        self.loss = self.oracle.epoch(self.desc, self.arrs, lrate)
        self._view(ptr)[:] = self._pack() - self.old

    def epoch_sharded_finish(self, ptr, scale=1.0):
        new = self.old + scale * self._view(ptr)
        off = 0
        for k in self.members:
            a = self.arrs[k]
            a.reshape(-1)[:] = new[off:off + a.size]
            off += a.size
        return self.loss

    def upload(self, arrs):
        pass

    def download(self, arrs):
        pass

    def stats(self):
        return None

    def eval_ratings(self, u, j, ctx, r, lo, hi):
        sa, ss, _ = self.oracle.eval_ratings(self.desc, self.arrs, u, j, ctx, r, lo, hi)
        return sa, ss

    def close(self):
        pass


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, model_name, F, epochs, out_dir, combine):
    import torch
    import torch.distributed as dist
    from oracle import oracle_py as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts, test, init = _problem(model_name, F)
        shard, lo = sharding.shard_training_set(ts, rank, world)
        hi = lo + shard.num_users
        Rec = recommender.getRecommender(model_name)

        class CpuRec(Rec):  # same orchestration, oracle-backed engine, CPU tensors for the exchange
            def _new_engine(self):
                return OracleEngine(self._desc(), self.MODEL, self.model, orc)

            def _exchange_device(self):
                return torch.device("cpu")

            def _eval_engine(self):
                return self.engine if self.engine is not None else OracleEngine(self._desc(), self.MODEL, self.model, orc)

        conf = {"num.factors": str(F), "num.max.iter": str(epochs)}
        rec = CpuRec(shard, sharding.shard_test_set(test, lo, hi), conf=conf, world=world, combine=combine)
        local = {k: (sharding.shard_user_rows(v, lo, hi) if k in ("P", "user_bias", "uc_bias") else v.copy())
                 for k, v in init.items()}
        rec.initModel(init=local)
        rec.keep_engine = True
        rec.buildModel()
        m = rec.evalRatings()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(rec.iter_losses), rmse=m["RMSE"],
                 mae=m["MAE"], lo=lo, **rec.model)
    finally:
        dist.destroy_process_group()


def _problem(model_name, F):
    from oracle import oracle_py as orc
    model = capi.MODEL_NAMES[model_name]
    dims = [3, 4] if model in (capi.CAMF_CI, capi.CAMF_CU, capi.CAMF_CUCI) else None
    ts, test = synth.make_training_set(120, 40, dims, 6000, seed=11, order="user_sorted", holdout=0.1)
    g = orc.JavaRandom(5)
    shapes = capi.member_shapes(model, ts.num_users, ts.num_items, ts.num_conditions, F)
    init = {k: (g.uniform(s) if k in ("ic_bias", "uc_bias") else g.gaussian(s)) for k, s in shapes.items()}
    return ts, test, init


def block_jacobi_reference(oracle, model_name, F, epochs, world, combine="mean"):
    """Single-process statement of the sharded semantics, driven like buildModel()."""
    model = capi.MODEL_NAMES[model_name]
    ts, test, init = _problem(model_name, F)
    shards = [sharding.shard_training_set(ts, g, world) for g in range(world)]
    user_side = ("P", "user_bias", "uc_bias")
    item = {k: v.copy() for k, v in init.items() if k not in user_side}
    locals_ = [{k: sharding.shard_user_rows(v, lo, lo + sh.num_users) for k, v in init.items() if k in user_side}
               for sh, lo in shards]
    lr, last, losses = capi.f32(0.02), 0.0, []
    scale = 1.0 / world if combine == "mean" else 1.0
    for it in range(1, epochs + 1):
        deltas, loss = [], 0.0
        for (sh, lo), loc in zip(shards, locals_):
            arrs = {**loc, **{k: v.copy() for k, v in item.items()}}
            loss += oracle.epoch(capi.make_desc(sh, model, F, **REGS), arrs, lr)
            deltas.append({k: arrs[k] - item[k] for k in item})
        for k in item:
            s = deltas[0][k]
            for d in deltas[1:]:
                s = s + d[k]
            item[k] = item[k] + scale * s
        losses.append(loss)
        if it > 1:  # bold driver, IterativeRecommender.java:216-229
            lr = lr * 1.05 if abs(last) > abs(loss) else lr * 0.5
        last = loss
    return shards, locals_, item, losses


@pytest.mark.parametrize("model_name,combine", [("camf_ci", "mean"), ("camf_cu", "mean"), ("camf_cuci", "mean"), ("biasedmf", "sum")])
def test_two_rank_gloo_matches_block_jacobi_reference(oracle, tmp_path, model_name, combine):
    import torch.multiprocessing as mp
    F, epochs, world = 8, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), model_name, F, epochs, str(tmp_path), combine), nprocs=world, join=True)
    shards, locals_, item, losses = block_jacobi_reference(oracle, model_name, F, epochs, world, combine)
    for g in range(world):
        got = np.load(tmp_path / f"rank{g}.npz")
        # the sum of two deltas is order-independent, so world = 2 is bit-exact; the loss is a sum of two terms
        for k, v in locals_[g].items():
            assert np.array_equal(got[k], v), k
        for k, v in item.items():
            assert np.array_equal(got[k], v), k
        np.testing.assert_allclose(got["losses"], losses, rtol=1e-14)
    assert np.load(tmp_path / "rank0.npz")["rmse"] == np.load(tmp_path / "rank1.npz")["rmse"]
