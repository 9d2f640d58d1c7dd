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
from typing import Dict, List, Optional, Tuple

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
                     ctx_cond=np.asarray(flat, dtype=np.int32), global_mean=total / len(r) if len(r) else 0.0,
                     rating_scale=(dao.ratingScale[0], dao.ratingScale[-1]) if dao.ratingScale else None,
                     num_context_dims=dao.numContextDims())
    ts.pair_ids = ui  # the CRS row (user-item pair id) of every entry, for callers that need it
    ts.empty_conditions = np.asarray(dao.EmptyContextConditions, dtype=np.int32)
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
                            global_mean=total / len(r) if len(r) else 0.0, rating_scale=ts.rating_scale,
                            num_context_dims=ts.num_context_dims)
        test = {"u": ts.u[test_mask].copy(), "j": ts.j[test_mask].copy(),
                "ctx": ts.ctx[test_mask].copy() if has_ctx else None, "r": ts.r[test_mask].copy()}
        if getattr(ts, "pair_ids", None) is not None:  # the CRS rows (user-item pair ids) travel with the entries
            train.pair_ids = ts.pair_ids[keep]
            test["pair_ids"] = ts.pair_ids[test_mask].copy()
        return train, test


# ------------------------------------------------------------------------------------------------------
# Format conversion: carskit.data.processor.DataTransformer (src/carskit/data/processor/DataTransformer.java)
# ------------------------------------------------------------------------------------------------------
def java_string_hash(s: str) -> int:
    """String.hashCode(): h = 31 * h + c over the UTF-16 code units, in 32-bit two's complement."""
    h = 0
    units = s.encode("utf-16-be", "surrogatepass")
    for i in range(0, len(units), 2):
        h = (31 * h + ((units[i] << 8) | units[i + 1])) & 0xFFFFFFFF
    return h


def java_hashmap_key_order(keys, jdk: int = 7) -> List[str]:
    """Iteration order of a java.util.HashMap<String, ?> filled by put() in the given order -- the order in which
    DataTransformer writes the converted rating lines (`for (String key : newlines.keySet())`,
    DataTransformer.java:276).  It depends on the JDK the reference runs on:

    jdk = 7 (the files under the reference's sampleData/ were written by one: tests/test_data.py reproduces them
      byte for byte): hash = h ^ (h >>> 20) ^ (h >>> 12), then ^ (>>> 7) ^ (>>> 4); a new entry goes to the HEAD
      of its bucket; the table doubles when size >= 0.75 * capacity AND the target bucket is occupied, and the
      transfer re-inserts every entry at the head of its new bucket (reversing each bucket).
    jdk = 8: hash = h ^ (h >>> 16); entries are appended, a resize keeps their relative order; doubles while
      size > 0.75 * capacity.  A bucket that would be treeified (>= 8 entries at capacity >= 64) iterates in an
      order this function does not reproduce; it raises instead."""
    uniq, seen = [], set()
    for k in keys:
        if k not in seen:
            seen.add(k)
            uniq.append(k)
    if jdk == 7:
        cap, table = 16, [[] for _ in range(16)]  # a bucket is listed from its head
        for size, k in enumerate(uniq):
            h = java_string_hash(k)
            h ^= (h >> 20) ^ (h >> 12)
            h = (h ^ (h >> 7) ^ (h >> 4)) & 0xFFFFFFFF
            if size >= int(cap * 0.75) and table[h & (cap - 1)]:
                cap *= 2
                new = [[] for _ in range(cap)]
                for b in table:
                    for kk, hh in b:
                        new[hh & (cap - 1)].insert(0, (kk, hh))
                table = new
            table[h & (cap - 1)].insert(0, (k, h))
        return [kk for b in table for kk, _ in b]
    if jdk != 8:
        raise ValueError("jdk must be 7 or 8")
    cap = 16
    while len(uniq) > 0.75 * cap:
        cap *= 2
    buckets: Dict[int, List[str]] = {}
    for k in uniq:
        h = java_string_hash(k)
        buckets.setdefault((h ^ (h >> 16)) & (cap - 1), []).append(k)
    if cap >= 64 and any(len(b) >= 8 for b in buckets.values()):
        raise NotImplementedError("a HashMap bucket reaches the treeify threshold")
    out: List[str] = []
    for b in sorted(buckets):
        out.extend(buckets[b])
    return out


def _detect_format(path: str) -> int:
    """1 binary / 2 loose / 3 compact, by the header like the reference's users declare it in setting.conf."""
    with open(path) as f:
        header = [h.strip().lower() for h in f.readline().rstrip("\r\n").split(",")]
    if len(header) == 5 and header[3] == "dimension" and header[4] == "condition":
        return 2
    if len(header) > 3 and all(":" in h for h in header[3:]):
        return 1
    return 3


def _read_lines(path: str) -> List[str]:
    with open(path) as f:
        return [ln.rstrip("\r\n") for ln in f if ln.rstrip("\r\n") != ""]  # BufferedReader.readLine()


class _Conditions:
    """dim -> conditions: a Guava TreeMultimap (both sorted, DataTransformer.java:60) or LinkedHashMultimap
    (insertion order, :163/:203/:243)."""

    def __init__(self, sorted_: bool):
        self.sorted = sorted_
        self.map: Dict[str, List[str]] = {}

    def put(self, dim: str, cond: str):
        conds = self.map.setdefault(dim, [])
        if cond not in conds:
            conds.append(cond)

    def dims(self) -> List[str]:
        return sorted(self.map) if self.sorted else list(self.map)

    def conds(self, dim: str) -> List[str]:
        return sorted(self.map[dim]) if self.sorted else list(self.map[dim])


def _scan_conditions(path: str, fmt: int, conditions: _Conditions):
    lines = _read_lines(path)
    header = lines[0].split(",")
    if fmt == 1:      # getConditionsFromBinaryData (:94-103)
        for h in header[3:]:
            d, c = h.split(":")[0:2]
            conditions.put(d.strip().lower(), c.strip().lower())
    elif fmt == 2:    # getConditionsFromLooseData (:105-117)
        for ln in lines[1:]:
            s = ln.split(",")
            conditions.put(s[3].strip().lower(), s[4].strip().lower() or "na")
    else:             # getConditionsFromCompactData (:119-138)
        dims = [h.strip().lower() for h in header[3:]]
        for ln in lines[1:]:
            s = ln.split(",")
            for i, d in enumerate(dims):
                conditions.put(d, s[3 + i].strip().lower() or "na")


def _to_binary(path: str, fmt: int, is_test: bool, conditions: Optional[_Conditions], jdk: int = 7) -> str:
    """TransformationFrom{Binary,Loose,Compact}ToBinary + PublishNewRatingFiles (:150-297): the text of the new
    binary-format file."""
    lines = _read_lines(path)
    header = lines[0].split(",")
    if conditions is None:
        conditions = _Conditions(sorted_=False)
    newlines: Dict[str, Dict[str, str]] = {}
    order: List[str] = []

    def put_line(key: str, ctx: Dict[str, str]):
        if key not in newlines:
            order.append(key)
        newlines[key] = ctx

    if fmt == 1:
        for ln in lines[1:]:
            s = ln.split(",")
            ctx: Dict[str, str] = {}
            for i in range(3, len(header)):
                if int(s[i].strip()) == 0:
                    continue
                d, c = [t.strip().lower() for t in header[i].split(":")[0:2]]
                ctx[d] = c
                if not is_test:
                    conditions.put(d, c)
            put_line(ln, ctx)
    elif fmt == 2:
        for ln in lines[1:]:
            s = ln.split(",")
            key = ",".join(t.strip().lower() for t in s[0:3])
            cond = s[4].strip().lower() or "na"
            if not is_test:
                conditions.put(s[3].strip().lower(), cond)
            if key in newlines:
                newlines[key][s[3].strip().lower()] = cond
            else:
                put_line(key, {s[3].strip().lower(): cond})
    else:
        dims = [h.strip().lower() for h in header[3:]]
        for ln in lines[1:]:
            s = ln.split(",")
            ctx = {}
            for i, d in enumerate(dims):
                cond = s[3 + i].strip().lower() or "na"
                ctx[d] = cond
                if not is_test:
                    conditions.put(d, cond)
            put_line(ln, ctx)
    out = ["User, Item, Rating" + "".join(f", {d}:{c}" for d in conditions.dims() for c in conditions.conds(d))]
    for key in java_hashmap_key_order(order, jdk):
        ctx = newlines[key]
        flags: List[str] = []
        for d in conditions.dims():
            mine = ctx.get(d)
            is_na = mine is None or mine == "na"
            done = False
            for c in conditions.conds(d):
                if fmt == 2:   # isLoose (:262-279)
                    if is_na:
                        flags.append("1" if c == "na" else "0")
                    elif done:
                        flags.append("0")
                    else:
                        done = c == mine
                        flags.append("1" if done else "0")
                else:          # :280-285 (a rating line without this dimension is a NullPointerException there)
                    if mine is None:
                        raise ValueError(f"rating line '{key}' has no condition for dimension '{d}'")
                    flags.append("1" if mine == c else "0")
        s = key.split(",")
        if len(s) > 3:
            key = ",".join(t.strip().lower() for t in s[0:3])
        out.append(key + "," + ",".join(flags))
    return "\n".join(out) + "\n"


def transform_to_binary(train_path: str, test_path: Optional[str] = None, jdk: int = 7):
    """DataTransformer.run() (:299-395).  One file: converted on its own -- dimensions / conditions in order of
    first appearance, no `na` columns added.  Train + test: the conditions of BOTH files, sorted, plus an `na`
    condition per dimension (getConditions, :57-92), shape the header of both outputs.  Returns the text of
    train.csv (and of test.csv).  The ORDER of the rating lines is the iteration order of a java.util.HashMap and so
    depends on the JVM (`jdk`, see java_hashmap_key_order); it matters downstream because user / item / pair /
    context ids are handed out by first appearance (DataDAO.java:238-330)."""
    f_train = _detect_format(train_path)
    if test_path is None:
        if f_train == 1:
            with open(train_path) as f:
                return f.read()  # FileIO.copyFile (:304)
        return _to_binary(train_path, f_train, False, None, jdk)
    f_test = _detect_format(test_path)
    conditions = _Conditions(sorted_=True)
    _scan_conditions(train_path, f_train, conditions)
    _scan_conditions(test_path, f_test, conditions)
    for d in conditions.dims():
        if "na" not in conditions.map[d]:
            conditions.put(d, "na")
    return _to_binary(train_path, f_train, False, conditions, jdk), _to_binary(test_path, f_test, True, conditions, jdk)


def to_traditional(ts: TrainingSet) -> TrainingSet:
    """DataDAO.toTraditionalSparseMatrix (DataDAO.java:1241-1257): the 2-D user x item `train` matrix that PMF /
    BiasedMF iterate (Recommender.java:252): one entry per user-item PAIR that has ratings, its value the mean of
    the pair's ratings over contexts (SparseVector.mean: sequential sum / count), in CRS order (user ascending,
    item ascending).  Needs ts.pair_ids (read_binary_csv and DataSplitter provide them).  globalMean stays the one
    of the {pair x context} matrix (Recommender.java:265 takes it from trainMatrix, not from `train`)."""
    pair = np.asarray(ts.pair_ids, dtype=np.int64)
    if pair.shape[0] != ts.nnz:
        raise ValueError("pair_ids must list the user-item pair id of every entry")
    sums: Dict[int, List[float]] = {}
    owner: Dict[int, Tuple[int, int]] = {}
    for p, u, j, r in zip(pair.tolist(), ts.u.tolist(), ts.j.tolist(), ts.r.tolist()):
        acc = sums.setdefault(p, [0.0, 0])
        acc[0] += r
        acc[1] += 1
        owner[p] = (u, j)
    cells: Dict[Tuple[int, int], float] = {}
    for p in sorted(sums):                      # `for (int uiid : sm.rows())`; Table.put: a later pair of the same
        cells[owner[p]] = sums[p][0] / sums[p][1]  # (user, item) would overwrite -- pair ids are unique per (u, i)
    keys = sorted(cells)
    out = TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=np.array([k[0] for k in keys], dtype=np.int32),
                      j=np.array([k[1] for k in keys], dtype=np.int32), r=np.array([cells[k] for k in keys], dtype=np.float64),
                      ctx=None, global_mean=ts.global_mean, rating_scale=ts.rating_scale,
                      num_context_dims=ts.num_context_dims)
    return out
