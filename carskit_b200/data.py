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
