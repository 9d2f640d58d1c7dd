"""Synthetic (users, items, contexts, nnz) training sets shaped like BASELINE.json's configs.

Harness code (tests + bench): builds the flattened arrays a Java buildModel() would hand to the C ABI,
in the reference's iteration order: user-item pair id ascending, then context id ascending
(CAMF_CI.java:80; pair ids as DataDAO assigns them to a file sorted by user then item).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .capi import TrainingSet


def context_table(dims: Sequence[int]) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """All combinations of one condition per dimension.  Context id = mixed-radix number of its
    conditions; condition ids are numbered dimension by dimension like the one-hot header columns
    (DataDAO.java:202-215).  Returns (ctx_ptr, ctx_cond, num_contexts, num_conditions)."""
    dims = [int(d) for d in dims]
    D = len(dims)
    num_ctx = int(np.prod(dims))
    offs = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
    ids = np.arange(num_ctx, dtype=np.int64)
    cond = np.empty((num_ctx, D), dtype=np.int32)
    rem = ids.copy()
    for k, d in enumerate(dims):
        cond[:, k] = (rem % d + offs[k]).astype(np.int32)
        rem //= d
    ctx_ptr = (np.arange(num_ctx + 1, dtype=np.int64) * D).astype(np.int32)
    return ctx_ptr, cond.reshape(-1).copy(), num_ctx, int(sum(dims))


def make_training_set(num_users: int, num_items: int, dims: Optional[Sequence[int]], nnz: int, seed: int,
                      order: str = "user_sorted", item_zipf: float = 0.0, rating_levels: int = 5,
                      holdout: float = 0.0, planted_rank: int = 0, planted_noise: float = 0.5):
    """Unique (u, j, ctx) triples with u uniform, j uniform (or Zipf(item_zipf)), ctx uniform, ratings
    uniform in {1..rating_levels} -- or, with planted_rank > 0, ratings that CAN be learnt:
    round(clip(mid + b_u + b_j + <p_u, q_j> + b_cond(ctx) + N(0, planted_noise))) from a hidden model of that rank,
    so that a held-out RMSE says something about the trained model (used by the convergence comparisons of the
    non serial-equivalent modes: FAST and multi-GPU).

    order = "user_sorted": pair ids follow (u, j) order, i.e. a ratings file sorted by user then item.
    order = "shuffled":    pair ids in random first-appearance order (an unsorted ratings file).
    Returns (train TrainingSet, test dict or None)."""
    rng = np.random.default_rng(seed)
    u = rng.integers(0, num_users, size=nnz, dtype=np.int64)
    if item_zipf > 0:
        w = 1.0 / np.power(np.arange(1, num_items + 1, dtype=np.float64), item_zipf)
        j = rng.choice(num_items, size=nnz, p=w / w.sum()).astype(np.int64)
    else:
        j = rng.integers(0, num_items, size=nnz, dtype=np.int64)
    if dims is not None and float(np.prod([float(d) for d in dims])) > max(4.0 * nnz, 1e6):
        # far more combinations than ratings (Frappe: 7*7*2*3*2*9*80*233 = 98.6 M for 96 K rows): like DataDAO,
        # only the contexts that OCCUR get an id (ids in lexicographic order of their condition tuples)
        # a pool of about nnz/5 distinct contexts, reused by the ratings (Frappe: 18.6 K contexts for 96 K rows)
        offs = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
        pool = max(1000, nnz // 5)
        rows = np.stack([rng.integers(0, d, size=pool, dtype=np.int64) for d in dims], axis=1)
        uniq = np.unique(rows, axis=0)
        num_ctx, num_cond = int(uniq.shape[0]), int(sum(dims))
        c = rng.integers(0, num_ctx, size=nnz, dtype=np.int64)
        ctx_cond = (uniq + offs[None, :]).astype(np.int32).reshape(-1).copy()
        ctx_ptr = (np.arange(num_ctx + 1, dtype=np.int64) * len(dims)).astype(np.int32)
    elif dims is not None:
        ctx_ptr, ctx_cond, num_ctx, num_cond = context_table(dims)
        c = rng.integers(0, num_ctx, size=nnz, dtype=np.int64)
    else:
        ctx_ptr = ctx_cond = None
        num_ctx, num_cond = 1, 0
        c = np.zeros(nnz, dtype=np.int64)
    pair = u * num_items + j
    if order == "shuffled":
        # pair id = rank of first appearance in the (random) generation order
        uniq, first = np.unique(pair, return_index=True)
        rank_of = np.empty(uniq.shape[0], dtype=np.int64)
        rank_of[np.argsort(first, kind="stable")] = np.arange(uniq.shape[0])
        ui = rank_of[np.searchsorted(uniq, pair)]
    elif order == "user_sorted":
        ui = pair
    else:
        raise ValueError(order)
    key = ui * num_ctx + c
    if order == "user_sorted":
        # the key encodes (u, j, ctx): sort + dedupe + decode (no argsort; 100 M keys in seconds)
        del u, j, c, pair, ui
        key.sort()
        if key.shape[0] > 1:
            key = key[np.concatenate([[True], key[1:] != key[:-1]])]
        pr, c = np.divmod(key, num_ctx)
        u, j = np.divmod(pr, num_items)
        del key, pr
        u, j, c = u.astype(np.int32), j.astype(np.int32), c.astype(np.int32)
    else:
        key, idx = np.unique(key, return_index=True)  # CRS order; duplicates: keep one
        u, j, c = u[idx].astype(np.int32), j[idx].astype(np.int32), c[idx].astype(np.int32)
    if planted_rank > 0:
        k = int(planted_rank)
        prng = np.random.default_rng(seed + 104729)
        hp = prng.standard_normal((num_users, k)) * (0.8 / np.sqrt(k))
        hq = prng.standard_normal((num_items, k)) * (1.0)
        hbu, hbj = 0.4 * prng.standard_normal(num_users), 0.4 * prng.standard_normal(num_items)
        hbc = 0.3 * prng.standard_normal(max(num_cond, 1))
        val = (rating_levels + 1) / 2.0 + hbu[u] + hbj[j] + np.einsum("nk,nk->n", hp[u], hq[j])
        if dims is not None:
            D = int(ctx_ptr[1] - ctx_ptr[0])
            val += hbc[ctx_cond.reshape(-1, D)[c]].sum(axis=1)
        val += planted_noise * prng.standard_normal(u.shape[0])
        r = np.clip(np.rint(val), 1, rating_levels).astype(np.float64)
    else:
        r = rng.integers(1, rating_levels + 1, size=u.shape[0]).astype(np.float64)

    test = None
    if holdout > 0:
        mask = rng.random(u.shape[0]) < holdout
        test = {"u": u[mask].copy(), "j": j[mask].copy(), "ctx": c[mask].copy() if dims is not None else None,
                "r": r[mask].copy()}
        keep = ~mask
        u, j, c, r = u[keep], j[keep], c[keep], r[keep]
    # SparseMatrix.getGlobalAvg: sequential sum / count of non-zeros (structure/SparseMatrix.java:49-56)
    gm = float(np.cumsum(r)[-1] / np.count_nonzero(r)) if r.shape[0] else 0.0
    ts = TrainingSet(num_users=num_users, num_items=num_items, u=u, j=j, r=r,
                     ctx=c if dims is not None else None, num_conditions=num_cond,
                     num_contexts=num_ctx if dims is not None else 0, ctx_ptr=ctx_ptr, ctx_cond=ctx_cond,
                     global_mean=gm)
    if dims is not None:
        # the first condition of every dimension plays the "dim:na" role (rateDao.getEmptyContextConditions(); CAMF_ICS)
        ts.empty_conditions = np.concatenate([[0], np.cumsum([int(d) for d in dims])[:-1]]).astype(np.int32)
        ts.num_context_dims = len(dims)
    return ts, test
