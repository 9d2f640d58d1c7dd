"""User-range sharding of the SGD hot path across the GPUs of one box.

The reference is single-process, so this layer has no Java counterpart; it defines the N > 1 semantics:

  * rank g of `world` owns the contiguous user range [lo, hi) (users are the rows of P): its ratings, its
    rows of P, userBias (and ucBias for CAMF_CU) -- trained locally, never communicated;
  * the item block (Q, itemBias, icBias) is replicated; every epoch each rank trains against its copy and
    the ranks combine   block <- old + scale * sum_g (new_g - old)   with ONE all-reduce (NCCL over NVLink).
    scale = 1/world ("mean", the default) averages the ranks' item blocks -- stable for any shard size;
    scale = 1 ("sum") adds the ranks' steps, which is only safe while one epoch moves an item row a little
    (with ~1000 ratings per item and shard every rank nearly converges the row on its own, and the summed
    step overshoots world-fold: measured to diverge on the bench workload at world = 2);
  * within a rank the epoch is the EXACT serial-equivalent engine epoch; across ranks it is a block-Jacobi
    step (not serial-equivalent -- DESIGN.md "Multi-GPU" states the tolerance the tests use).

Nothing here computes an update: the delta / apply passes are CUDA kernels inside libcarskit_b200.so
(cars_epoch_sharded_begin / _finish); this module only moves a device buffer through torch.distributed.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .capi import TrainingSet


def user_range(num_users: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ranges: the first (num_users % world) ranks get one extra user."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(num_users, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_training_set(ts: TrainingSet, rank: int, world: int) -> Tuple[TrainingSet, int]:
    """The ratings whose user lies in this rank's range, in the reference's iteration order, with user ids
    rebased to the range (local row of P = u - lo).  Returns (shard, lo)."""
    lo, hi = user_range(ts.num_users, rank, world)
    keep = (ts.u >= lo) & (ts.u < hi)
    shard = TrainingSet(num_users=hi - lo, num_items=ts.num_items, u=ts.u[keep] - lo, j=ts.j[keep], r=ts.r[keep],
                        ctx=None if ts.ctx is None else ts.ctx[keep], num_conditions=ts.num_conditions,
                        num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                        global_mean=ts.global_mean,  # globalMean is a property of the WHOLE training matrix
                        rating_scale=ts.rating_scale, num_context_dims=ts.num_context_dims)
    return shard, lo


def shard_test_set(test: Optional[dict], lo: int, hi: int) -> Optional[dict]:
    if test is None:
        return None
    keep = (test["u"] >= lo) & (test["u"] < hi)
    return {"u": test["u"][keep] - lo, "j": test["j"][keep],
            "ctx": None if test.get("ctx") is None else test["ctx"][keep], "r": test["r"][keep]}


def shard_user_rows(arr: np.ndarray, lo: int, hi: int) -> np.ndarray:
    return np.ascontiguousarray(arr[lo:hi])


def row_range(nnz: int, rank: int, world: int) -> Tuple[int, int]:
    """FM shards ROWS (ratings) by contiguous range of the reference order."""
    return user_range(nnz, rank, world)


def shard_rows(ts: TrainingSet, rank: int, world: int) -> TrainingSet:
    lo, hi = row_range(ts.nnz, rank, world)
    return TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=ts.u[lo:hi], j=ts.j[lo:hi], r=ts.r[lo:hi],
                       ctx=None if ts.ctx is None else ts.ctx[lo:hi], num_conditions=ts.num_conditions,
                       num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                       global_mean=ts.global_mean, rating_scale=ts.rating_scale, num_context_dims=ts.num_context_dims)


class CoordinateSumExchange:
    """FM: the all-reduce of the per-coordinate (numerator, denominator) sums between the row shards.  The engine
    calls back 3 * (1 + k) + 1 times per ALS iteration with a slice of the device buffer owned here."""

    def __init__(self, engine, device, group=None, torch_stream=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.torch_stream = torch, dist, group, torch_stream
        self.buf = torch.zeros(engine.exchange_doubles(), dtype=torch.float64, device=device)
        self.base = self.buf.data_ptr()
        self.calls = 0
        self.bytes = 0

    def allreduce(self, dev_ptr: int, count: int):
        off = (dev_ptr - self.base) // 8
        view = self.buf[off:off + count]
        if self.torch_stream is not None:
            with self.torch.cuda.stream(self.torch_stream):
                self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += count * 8

    def iteration(self, engine) -> float:
        return engine.iteration_sharded(self.base, self.allreduce)


class ItemBlockExchange:
    """Per-epoch exchange of the item block.  `device` is a torch device; the delta buffer lives there
    (CUDA for the engine; CPU tensors + gloo exercise the same orchestration in the CPU tests)."""

    def __init__(self, engine, device, group=None, combine: str = "mean", torch_stream=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        if combine not in ("mean", "sum"):
            raise ValueError(combine)
        self.scale = 1.0 / dist.get_world_size(group) if combine == "mean" else 1.0
        # the engine's CUDA stream as a torch stream: collectives issued under it are ordered with the kernels
        self.torch_stream = torch_stream
        self.n = engine.item_block_doubles()
        self.delta = torch.zeros(self.n, dtype=torch.float64, device=device)
        self.scalar = torch.zeros(1, dtype=torch.float64, device=device)
        self.bytes_per_epoch = self.n * 8

    def epoch(self, engine, lrate: float) -> float:
        """One sharded epoch; returns the GLOBAL loss (sum of the ranks' losses: the reference's loss is a sum
        over ratings, CAMF_CI.java:91-124)."""
        ptr = self.delta.data_ptr()
        engine.epoch_sharded_begin(lrate, ptr)
        with self._on_stream():
            self.dist.all_reduce(self.delta, op=self.dist.ReduceOp.SUM, group=self.group)
        local = engine.epoch_sharded_finish(ptr, self.scale)
        with self._on_stream():
            self.scalar[0] = local
            self.dist.all_reduce(self.scalar, op=self.dist.ReduceOp.SUM, group=self.group)
            return float(self.scalar.item())

    def _on_stream(self):
        import contextlib
        if self.torch_stream is None:
            return contextlib.nullcontext()
        return self.torch.cuda.stream(self.torch_stream)

    def sum_scalars(self, *vals: float):
        with self._on_stream():
            t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.delta.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            return [float(x) for x in t.tolist()]
