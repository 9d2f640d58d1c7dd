"""Binary-format CARSKit ratings file -> the flattened arrays the C ABI takes.

This is the input contract of the hot path stated as code: the id and layout rules of
carskit.data.processor.DataDAO.readData (src/carskit/data/processor/DataDAO.java:199-354) and of the CRS
`{user-item pair} x {context}` rating matrix it builds (librec SparseMatrix: rows ascending, columns sorted
ascending inside a row, zero entries absent).  Host logic only -- nothing here computes an update.

Rules restated (DataDAO.java line numbers):
  * header `User, Item, Rating, dim:cond, ...` split on [tab,]+ (:201); condition id = column index - 3 (:209);
    dimension id = order of first appearance of the text before ':' (:207-208); conditions ending ':na' are the
    "empty" ones (:214-215);
  * data lines split on ',' (:226); user / item inner ids by first appearance (:238-242); the user-item PAIR id
    by first appearance of "row,col" (:266-268); the context id by first appearance of the comma-joined list of
    condition ids whose flag is 1 (:279-330);
  * the same (pair, context) again overwrites the rating (Table.put, :342);
  * iteration order of `for (MatrixEntry me : trainMatrix)` = pair id ascending, context id ascending.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from .capi import TrainingSet


@dataclass
class RatingDao:
    """What the recommenders read from `rateDao` (DataDAO getters cited per field)."""
    userIds: Dict[str, int] = field(default_factory=dict)          # getUserIds
    itemIds: Dict[str, int] = field(default_factory=dict)          # getItemIds
    uiIds: Dict[str, int] = field(default_factory=dict)            # "row,col" -> pair id
    ctxIds: Dict[str, int] = field(default_factory=dict)           # "c1,c2,.." -> context id (getContextId, :945)
    condIds: Dict[str, int] = field(default_factory=dict)          # "dim:cond" -> condition id
    dimIds: Dict[str, int] = field(default_factory=dict)
    uiUserIds: List[int] = field(default_factory=list)             # getUserIdFromUI (:1038)
    uiItemIds: List[int] = field(default_factory=list)             # getItemIdFromUI (:1042)
    contextConditionsList: List[List[int]] = field(default_factory=list)  # getContextConditionsList (:1035)
    EmptyContextConditions: List[int] = field(default_factory=list)
    ratingScale: List[float] = field(default_factory=list)

    def numUsers(self): return len(self.userIds)
    def numItems(self): return len(self.itemIds)
    def numUserItems(self): return len(self.uiIds)
    def numContexts(self): return len(self.ctxIds)
    def numConditions(self): return len(self.condIds)
    def numContextDims(self): return len(self.dimIds)


def read_binary_csv(path: str) -> Tuple[TrainingSet, RatingDao]:
    dao = RatingDao()
    table: Dict[Tuple[int, int], float] = {}
    scale = set()
    with open(path) as f:
        header = re.split(r"[\t,]+", f.readline().strip())
        for i in range(3, len(header)):
            context = header[i].strip()
            dim = context.split(":")[0].strip()
            dao.dimIds.setdefault(dim, len(dao.dimIds))
            dao.condIds[context] = i - 3
            if context.endswith(":na"):
                dao.EmptyContextConditions.append(i - 3)
        for line in f:
            if line.endswith("\n"):
                line = line[:-1]
            if not line.strip():
                continue
            data = line.strip().split(",")
            user, item, rate = data[0], data[1], float(data[2])
            scale.add(rate)
            row = dao.userIds.setdefault(user, len(dao.userIds))
            col = dao.itemIds.setdefault(item, len(dao.itemIds))
            key = f"{row},{col}"
            if key not in dao.uiIds:
                dao.uiIds[key] = len(dao.uiIds)
                dao.uiUserIds.append(row)
                dao.uiItemIds.append(col)
            uic = dao.uiIds[key]
            conds = [i - 3 for i in range(3, len(data)) if int(data[i].strip()) == 1]
            ctx = ",".join(str(c) for c in conds)
            if ctx not in dao.ctxIds:
                dao.ctxIds[ctx] = len(dao.ctxIds)
                dao.contextConditionsList.append(conds)
            table[(uic, dao.ctxIds[ctx])] = rate  # Table.put: the last duplicate wins
    dao.ratingScale = sorted(scale)
    entries = sorted((k for k, v in table.items() if v != 0.0))  # CRS order; zero entries are not stored
    ui = np.array([e[0] for e in entries], dtype=np.int64)
    c = np.array([e[1] for e in entries], dtype=np.int32)
    r = np.array([table[e] for e in entries], dtype=np.float64)
    uu = np.asarray(dao.uiUserIds, dtype=np.int32)[ui] if len(entries) else np.empty(0, np.int32)
    jj = np.asarray(dao.uiItemIds, dtype=np.int32)[ui] if len(entries) else np.empty(0, np.int32)
    ctx_ptr = np.zeros(dao.numContexts() + 1, dtype=np.int32)
    flat: List[int] = []
    for k, conds in enumerate(dao.contextConditionsList):
        flat.extend(conds)
        ctx_ptr[k + 1] = len(flat)
    total = 0.0
    for v in r.tolist():  # SparseMatrix.getGlobalAvg: sequential sum / count of non-zeros (SparseMatrix.java:49-56)
        total += v
    ts = TrainingSet(num_users=dao.numUsers(), num_items=dao.numItems(), u=uu, j=jj, r=r, ctx=c,
                     num_conditions=dao.numConditions(), num_contexts=dao.numContexts(), ctx_ptr=ctx_ptr,
                     ctx_cond=np.asarray(flat, dtype=np.int32), global_mean=total / len(r) if len(r) else 0.0)
    ts.pair_ids = ui  # the CRS row (user-item pair id) of every entry, for callers that need it
    return ts, dao


# ------------------------------------------------------------------------------------------------------
# k-fold assignment: carskit.data.processor.DataSplitter (src/carskit/data/processor/DataSplitter.java)
# ------------------------------------------------------------------------------------------------------
_JR_MULT = np.uint64(0x5DEECE66D)
_JR_ADD = np.uint64(0xB)
_JR_MASK = np.uint64((1 << 48) - 1)


def java_random_doubles(seed: int, n: int) -> np.ndarray:
    """The first n values of `new java.util.Random(seed).nextDouble()` -- what happy.coding.math.Randoms.uniform()
    returns after Randoms.seed(seed) (CARSKit.java:174).  The 48-bit LCG
    s' = s * 0x5DEECE66D + 0xB is jumped ahead in closed form (s_k = A_k s_0 + C_k, all arithmetic wrapping
    mod 2^64 and masked to 48 bits), so the whole sequence is vectorised:
    nextDouble = ((next(26) << 27) + next(27)) * 2^-53, two LCG steps per double."""
    if n == 0:
        return np.empty(0, dtype=np.float64)
    s0 = np.uint64((int(seed) ^ 0x5DEECE66D) & ((1 << 48) - 1))
    steps = 2 * n
    with np.errstate(over="ignore"):
        a_pow = np.empty(steps + 1, dtype=np.uint64)  # A_k = a^k
        a_pow[0] = 1
        a_pow[1:] = _JR_MULT
        np.multiply.accumulate(a_pow, out=a_pow)
        c_sum = np.add.accumulate(a_pow[:-1]) * _JR_ADD  # C_k = c * (a^0 + .. + a^(k-1)), k = 1..steps
        states = (a_pow[1:] * s0 + c_sum) & _JR_MASK
    hi = (states[0::2] >> np.uint64(48 - 26)).astype(np.int64)
    lo = (states[1::2] >> np.uint64(48 - 27)).astype(np.int64)
    return ((hi << 27) + lo).astype(np.float64) * (1.0 / (1 << 53))


class DataSplitter:
    """`new DataSplitter(rateMatrix, kfold)` + getKthFold(k) (DataSplitter.java:46-50, 68-133) over the flattened
    CRS entries of a TrainingSet (entry f = the f-th (user-item pair, context) cell in iteration order).

    splitFolds: rdm[i] = Randoms.uniform(); fold[i] = (int)(i / (numRates / numFold)) + 1 (:108-118);
    Sortor.quickSort(rdm, fold, ..) sorts rdm ascending carrying fold (:120) -- an argsort, as long as no two
    draws are equal (checked); the f-th CRS entry then gets fold[f] (:125-132).  getKthFold(k): the entries
    labelled k are the TEST set, the others the training set, order kept (reshape drops the zeroed cells).
    `seed` = evaluation.setup's --rand-seed; the reference seeds the generator right before the splitter draws
    (CARSKit.java:174 precedes :393), so the draws are the first numRates doubles of Random(seed)."""

    def __init__(self, ts: TrainingSet, kfold: int, seed: int):
        self.ts = ts
        n = ts.nnz
        self.numFold = min(int(kfold), n) if n else int(kfold)
        if self.numFold < 1:
            raise ValueError("kfold must be > 0")
        rdm = java_random_doubles(seed, n)
        indv = (n + 0.0) / self.numFold
        fold = (np.arange(n, dtype=np.float64) / indv).astype(np.int32) + 1
        order = np.argsort(rdm, kind="stable")
        srt = rdm[order]
        ties = np.nonzero(srt[1:] == srt[:-1])[0]
        if ties.size and np.any(fold[order[ties]] != fold[order[ties + 1]]):
            # equal keys: the reference's Lomuto partition order decides; not reproduced (p ~ n^2 / 2^54)
            raise NotImplementedError("two equal random draws carry different fold labels")
        self.assign = fold[order]  # assignMatrix in CRS order

    def getKthFold(self, k: int):
        """Returns (train TrainingSet, test dict) for fold k (1-based), or None when k is out of range (:69-70)."""
        if k > self.numFold or k < 1:
            return None
        ts = self.ts
        test_mask = self.assign == k
        keep = ~test_mask
        has_ctx = ts.ctx is not None
        r = ts.r[keep]
        total = 0.0
        for v in r.tolist():  # SparseMatrix.getGlobalAvg of the TRAINING matrix (Recommender.java:265)
            total += v
        train = TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=ts.u[keep], j=ts.j[keep], r=r,
                            ctx=ts.ctx[keep] if has_ctx else None, num_conditions=ts.num_conditions,
                            num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr, ctx_cond=ts.ctx_cond,
                            global_mean=total / len(r) if len(r) else 0.0)
        test = {"u": ts.u[test_mask].copy(), "j": ts.j[test_mask].copy(),
                "ctx": ts.ctx[test_mask].copy() if has_ctx else None, "r": ts.r[test_mask].copy()}
        return train, test
