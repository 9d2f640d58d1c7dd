"""Host-side mirror of the reference's recommender classes for the SGD hot path.

The reference is Java and no JVM exists in this image, so the host side above the C ABI is written
here with the reference's own names, argument meaning and call order:

    Recommender.execute()              src/carskit/generic/Recommender.java:319-357
      initModel() -> buildModel() -> evalRatings()
    IterativeRecommender               src/carskit/generic/IterativeRecommender.java:36-108 (hyper-parameters),
      isConverged(iter)                :145-199
      updateLRate(iter)                :216-229
    PMF / BiasedMF                     src/carskit/alg/baseline/cf/{PMF,BiasedMF}.java
    CAMF_C / CAMF_CI / CAMF_CU / CAMF_CUCI  src/carskit/alg/cars/adaptation/dependent/dev/*.java

Only buildModel() differs from the reference: instead of the per-rating Java loop it flattens the
containers once and drives `cars_epoch` (include/carskit_b200.h) once per iteration, keeping
isConverged()/updateLRate() on the host exactly as a Java subclass would.  Nothing in
this module computes a rating update; without libcarskit_b200.so and an sm_100 device it raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

from . import capi
from .capi import TrainingSet


def _f32(x) -> float:
    return float(np.float32(x))


class LineConfiger:
    """`happy.coding.io.LineConfiger`: "main -opt v1 -flag ..." split on [,\\t ]; the first token that is
    not an option is the main parameter."""

    def __init__(self, line: str):
        toks = [t for t in line.replace(",", " ").replace("\t", " ").split(" ") if t]
        self.main: Optional[str] = None
        self.opts: Dict[str, list] = {}
        cur = None
        for t in toks:
            if t.startswith("-") and not _is_number(t):
                cur = t
                self.opts.setdefault(cur, [])
            elif cur is None:
                if self.main is None:
                    self.main = t
            else:
                self.opts[cur].append(t)

    def getMainParam(self) -> Optional[str]:
        return self.main

    def contains(self, key: str) -> bool:
        return key in self.opts

    def getFloat(self, key: str, default: float) -> float:
        v = self.opts.get(key)
        return _f32(v[0]) if v else _f32(default)

    def getString(self, key: str, default: Optional[str] = None) -> Optional[str]:
        v = self.opts.get(key)
        return v[0] if v else default


def java_hashset_order(ids_in_insertion_order) -> np.ndarray:
    """Iteration order of a java.util.HashSet<Integer> filled by add() in the given order (JDK 8 HashMap: table
    of 16 doubling while size > 0.75 * capacity, bucket = (h ^ (h >>> 16)) & (capacity - 1) with h = the int
    itself, buckets walked in index order, entries of one bucket in insertion order -- resizes preserve that).
    evalRankings() iterates candItems = rateDao.getItemList(trainMatrix) in this order (Recommender.java:704,
    DataDAO.java:1210-1218), and ties of the stable sort keep it, so it is part of the ranking contract."""
    seen, uniq = set(), []
    for v in ids_in_insertion_order:
        v = int(v)
        if v not in seen:
            seen.add(v)
            uniq.append(v)
    cap = 16
    while len(uniq) > 0.75 * cap:
        cap *= 2
    buckets: Dict[int, list] = {}
    for v in uniq:
        h = v & 0xFFFFFFFF
        buckets.setdefault((h ^ (h >> 16)) & (cap - 1), []).append(v)
    out = []
    for b in sorted(buckets):
        out.extend(buckets[b])
    return np.asarray(out, dtype=np.int32)


def _is_number(t: str) -> bool:
    try:
        float(t)
        return True
    except ValueError:
        return False


DEFAULT_CONF = {  # setting.conf:52-59
    "num.factors": "10",
    "num.max.iter": "100",
    "learn.rate": "2e-2 -max -1 -bold-driver",
    "reg.lambda": "0.0001 -c 0.001",
    "evaluation.setup": "test-set --test-view all",
}


class IterativeRecommender:
    """Hyper-parameters are Java *floats* widened to double at every use
    (IterativeRecommender.java:36-49): 0.02f -> 0.019999999552965164."""

    MODEL = capi.PMF
    algoName = "IterativeRecommender"
    GAUSSIAN_CONTEXT_BIAS = False  # icBias / ucBias ~ U(0,1) (CAMF_CI.java:58-59); CAMF_CUCI draws Gaussians

    def __init__(self, trainMatrix: TrainingSet, testMatrix: Optional[dict] = None, fold: int = -1,
                 conf: Optional[Dict[str, str]] = None, device: int = 0, stream: int = 0, world: int = 1,
                 group=None, combine: str = "mean", mode: str = "exact", tuning: Optional[str] = None, gpu_ids=None):
        """`world` > 1: this process is one rank of a user-range-sharded job (sharding.py); trainMatrix /
        testMatrix are THIS rank's shard and torch.distributed is initialised.  The engine then runs on
        torch's current CUDA stream so the all-reduce is ordered with the kernels."""
        cf = dict(DEFAULT_CONF)
        cf.update(conf or {})
        self.cf = cf
        self.trainMatrix, self.testMatrix, self.fold = trainMatrix, testMatrix, fold
        self.device, self.stream = device, stream
        self.world, self.group, self.exchange, self.combine = world, group, None, combine
        # engine.mode: "exact" (serial-equivalent, the default) or "fast" (hogwild, include/carskit_b200.h cars_mode);
        # no reference counterpart -- a B200 subclass would read it from the algorithm's option line
        self.mode = cf.get("engine.mode", mode).lower()
        if self.mode not in ("exact", "fast"):
            raise ValueError(f"engine.mode must be exact or fast, not {self.mode}")
        self.fastMaxConc = float(cf.get("engine.fast.max.conc", 0.0))
        self.tuning = tuning
        # gpu_ids with more than one entry: ONE engine handle drives those GPUs from this process (cars_desc.num_gpus);
        # `world` > 1 is the other route (one process per GPU, torch.distributed owns the collective)
        self.gpu_ids = list(gpu_ids) if gpu_ids is not None else None
        self._torch_stream = None
        self.numUsers, self.numItems = trainMatrix.num_users, trainMatrix.num_items
        self.numConditions = trainMatrix.num_conditions
        self.globalMean = trainMatrix.global_mean  # Recommender.java:265

        lc = LineConfiger(cf["learn.rate"])  # :83-90
        self.initLRate = _f32(lc.getMainParam())
        self.maxLRate = lc.getFloat("-max", -1)
        self.isBoldDriver = lc.contains("-bold-driver")
        self.decay = lc.getFloat("-decay", -1)
        ro = LineConfiger(cf["reg.lambda"])  # :92-99
        self.reg = _f32(ro.getMainParam())
        self.regU = ro.getFloat("-u", self.reg)
        self.regI = ro.getFloat("-i", self.reg)
        self.regB = ro.getFloat("-b", self.reg)
        self.regC = ro.getFloat("-c", self.reg)
        self.numFactors = int(cf["num.factors"])  # :101
        self.numIters = int(cf["num.max.iter"])  # :102
        ev = LineConfiger(cf["evaluation.setup"])
        self.earlyStopMeasure = ev.getString("--early-stop")  # Recommender.java:221-229
        # minRate / maxRate = first / last of rateDao.getRatingScale() (Recommender.java:196-200): they come with the
        # data (TrainingSet.rating_scale, filled by data.read_binary_csv).  Arrays that did not come through the
        # loader fall back to the scale of the training ratings themselves; rating.min / rating.max override both.
        scale = getattr(trainMatrix, "rating_scale", None)
        if scale is None and trainMatrix.nnz:
            scale = (float(trainMatrix.r.min()), float(trainMatrix.r.max()))
        if scale is None:
            scale = (1.0, 5.0)
        self.minRate = float(cf.get("rating.min", scale[0]))
        self.maxRate = float(cf.get("rating.max", scale[1]))
        self.initMean, self.initStd = 0.0, 0.1  # Recommender.java:203-204

        self.lRate = float(self.initLRate)  # :106
        self.loss = 0.0
        self.last_loss = 0.0
        self.measure = 0.0
        self.last_measure = 0.0
        self.verbose = False
        self.model: Dict[str, np.ndarray] = {}
        self.engine: Optional[capi.Engine] = None
        self.iter_losses = []
        self.measures: Dict[str, float] = {}

    # ---- model members ---------------------------------------------------------------------------
    def initModel(self, init: Optional[Dict[str, np.ndarray]] = None, seed: int = 0):
        """IterativeRecommender.initModel (:231-247) + the subclass' initModel.  `init` hands over arrays
        produced elsewhere (what a JNI subclass receives from Java).  Otherwise P, Q and the bias
        vectors are drawn N(initMean, initStd) and icBias/ucBias U(0,1) (CAMF_CI.java:55-60); the
        reference's own generator is wall-clock seeded, so any generator is faithful."""
        shapes = capi.member_shapes(self.MODEL, self.numUsers, self.numItems, self.numConditions, self.numFactors,
                                    getattr(self, "numContextFactors", 10))
        if init is not None:
            for k, s in shapes.items():
                a = np.ascontiguousarray(init[k], dtype=np.float64)
                if a.shape != tuple(s):
                    raise ValueError(f"{k}: shape {a.shape} != {s}")
                self.model[k] = a
            return
        rng = np.random.default_rng(seed)
        for k, s in shapes.items():
            if k in ("ic_bias", "uc_bias") and not self.GAUSSIAN_CONTEXT_BIAS:
                self.model[k] = rng.random(s)
            else:
                self.model[k] = self.initMean + self.initStd * rng.standard_normal(s)

    # ---- epoch control (host, unchanged semantics) --------------------------------------------------
    def updateLRate(self, iter: int):
        """IterativeRecommender.java:216-229."""
        if self.lRate <= 0:
            return
        if self.isBoldDriver and iter > 1:
            self.lRate = self.lRate * 1.05 if abs(self.last_loss) > abs(self.loss) else self.lRate * 0.5
        elif 0 < self.decay < 1:
            self.lRate *= self.decay
        if self.maxLRate > 0 and self.lRate > self.maxLRate:
            self.lRate = self.maxLRate

    def isConverged(self, iter: int) -> bool:
        """IterativeRecommender.java:145-199 (the debug print is kept behind `verbose`)."""
        delta_loss = _f32(self.last_loss - self.loss)
        if self.earlyStopMeasure is not None:
            if self.earlyStopMeasure.lower() == "loss":
                self.measure, self.last_measure = self.loss, self.last_loss
            else:
                self.measure = self.evalRatings()[self.earlyStopMeasure.upper()]
        delta_measure = _f32(self.last_measure - self.measure)
        if self.verbose:
            print(f"{self.algoName} iter {iter}: loss = {_f32(self.loss)}, delta_loss = {delta_loss}, "
                  f"learn_rate = {_f32(self.lRate)}")
        if math.isnan(self.loss) or math.isinf(self.loss):
            # the reference calls System.exit(-1) here (:181-184); a library raises instead
            raise FloatingPointError("Loss = NaN or Infinity: current settings does not fit the recommender!")
        cond1 = abs(self.loss) < 1e-5
        cond2 = (delta_measure > 0) and (delta_measure < 1e-5)
        converged = cond1 or cond2
        if not converged:
            self.updateLRate(iter)
        self.last_loss = self.loss
        self.last_measure = self.measure
        return converged

    # ---- the hot path ---------------------------------------------------------------------------------
    def _desc(self):
        return capi.make_desc(self.trainMatrix, self.MODEL, self.numFactors, device=self.device,
                              reg_u=self.regU, reg_i=self.regI, reg_b=self.regB, reg_c=self.regC,
                              stream=self.stream, mode=capi.FAST if self.mode == "fast" else capi.EXACT,
                              fast_max_conc=self.fastMaxConc, tuning=self.tuning, gpu_ids=self.gpu_ids,
                              num_context_factors=getattr(self, "numContextFactors", 10),
                              combine={"mean": capi.COMBINE_MEAN, "sum": capi.COMBINE_SUM, "touched": capi.COMBINE_TOUCHED}[self.combine])

    def open_engine(self) -> capi.Engine:
        """cars_create + cars_upload: what buildModel() does before its first iteration."""
        if not self.model:
            raise RuntimeError("buildModel before initModel")
        import time
        t0 = time.perf_counter()
        eng = self._new_engine()
        t1 = time.perf_counter()
        try:
            eng.upload(self.model)
            self.open_seconds = {"create": t1 - t0, "upload": time.perf_counter() - t1}
            if self.world > 1:
                from .sharding import ItemBlockExchange
                self.exchange = ItemBlockExchange(eng, self._exchange_device(), self.group, self.combine,
                                                  self._torch_stream)
        except Exception:
            eng.close()
            raise
        self.engine = eng
        return eng

    def _new_engine(self):
        if self.world > 1:  # the all-reduce must be stream-ordered with the kernels
            import torch
            if self.stream:
                self._torch_stream = torch.cuda.ExternalStream(self.stream, device=self.device)
            else:
                self._torch_stream = torch.cuda.Stream(device=self.device)
                self.stream = self._torch_stream.cuda_stream
        return capi.Engine(self._desc(), keepalive=self.trainMatrix)

    def _exchange_device(self):
        import torch
        return torch.device("cuda", self.device)

    def train_epoch(self, iter: int) -> bool:
        """One iteration of the `for (int iter = 1; ...)` loop (CAMF_CI.java:77-128): the per-rating pass on
        the device, then isConverged(iter) on the host.  Returns isConverged's verdict."""
        if self.world > 1:
            self.loss = self.exchange.epoch(self.engine, self.lRate)
        else:
            self.loss = self.engine.epoch(self.lRate)
        self.iter_losses.append(self.loss)
        return self.isConverged(iter)

    def close_engine(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
        self.exchange = None

    def buildModel(self):
        """Replaces the per-rating loop of buildModel() (CAMF_CI.java:74-131 and siblings): flatten once,
        one cars_epoch per iteration, isConverged() on the host, copy the model back."""
        import time
        self.iter_losses = []
        t0 = time.perf_counter()
        eng = self.open_engine()
        t1 = time.perf_counter()
        try:
            for it in range(1, self.numIters + 1):
                if self.train_epoch(it):
                    break
            t2 = time.perf_counter()
            eng.download(self.model)
            t3 = time.perf_counter()
            self.stats = eng.stats()
        finally:
            if not getattr(self, "keep_engine", False):
                self.close_engine()
        # where the wall time of one buildModel() went (seconds): create + upload, the epoch loop, download, destroy
        self.phase_seconds = {"open": t1 - t0, "epochs": t2 - t1, "download": t3 - t2, "close": time.perf_counter() - t3}
        self.phase_seconds.update(getattr(self, "open_seconds", {}))

    # ---- consumers of the trained model ------------------------------------------------------------------
    def _eval_engine(self) -> capi.Engine:
        if self.engine is not None:
            return self.engine
        eng = capi.Engine(self._desc_for_predict(), keepalive=self.trainMatrix)
        eng.upload(self.model)
        return eng

    def _desc_for_predict(self):
        # an engine without training ratings: predict()/evalRatings() only need the model
        ts = self.trainMatrix
        empty = TrainingSet(num_users=ts.num_users, num_items=ts.num_items, u=np.empty(0, np.int32),
                            j=np.empty(0, np.int32), r=np.empty(0, np.float64),
                            ctx=None if ts.ctx is None else np.empty(0, np.int32),
                            num_conditions=ts.num_conditions, num_contexts=ts.num_contexts, ctx_ptr=ts.ctx_ptr,
                            ctx_cond=ts.ctx_cond, global_mean=ts.global_mean)
        self._empty_keep = empty
        return capi.make_desc(empty, self.MODEL, self.numFactors, device=self.device, reg_u=self.regU,
                              reg_i=self.regI, reg_b=self.regB, reg_c=self.regC)

    def predict(self, u, j, c=None, bound: bool = False) -> np.ndarray:
        """Recommender.predict(u, j, c, bound) (Recommender.java:306-317), batched."""
        eng = self._eval_engine()
        try:
            return eng.predict(u, j, c, bound=bound, min_rate=self.minRate, max_rate=self.maxRate)
        finally:
            if eng is not self.engine:
                eng.close()

    def _allreduce_sum(self, vals):
        if self.world <= 1:
            return list(vals)
        import torch
        import torch.distributed as dist
        import contextlib
        ctx = torch.cuda.stream(self._torch_stream) if self._torch_stream is not None else contextlib.nullcontext()
        with ctx:
            t = torch.tensor(list(vals), dtype=torch.float64, device=self._exchange_device())
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return [float(x) for x in t.tolist()]

    def evalRatings(self) -> Dict[str, float]:
        """Recommender.evalRatings (:504-594): MAE / RMSE over testMatrix with bounded predictions.
        Sharded runs sum |err|, err^2 and the count over the ranks (each rank holds its users' test ratings)."""
        t = self.testMatrix
        if t is None:
            return {"MAE": float("nan"), "RMSE": float("nan")}
        n = len(t["u"])
        sa = ss = 0.0
        if n:
            eng = self._eval_engine()
            try:
                sa, ss = eng.eval_ratings(t["u"], t["j"], t.get("ctx"), t["r"], self.minRate, self.maxRate)
            finally:
                if eng is not self.engine:
                    eng.close()
        sa, ss, n = self._allreduce_sum([sa, ss, float(n)])
        if n == 0:
            return {"MAE": float("nan"), "RMSE": float("nan")}
        return {"MAE": sa / n, "RMSE": math.sqrt(ss / n)}

    def evalRankings(self, numRecs: int = 10, binThold: float = -1.0):
        """The scoring half of Recommender.evalRankings (:738-824) on the device: for every test user u and every
        context c in which u has a positive test rating (rating > binThold, DataDAO.java:1114-1139), all
        candidate items (items of trainMatrix, in the reference's HashSet order) that u has not rated in c in
        the training set are scored, filtered by score > binThold, stably sorted and cut to numRecs.
        Returns a list of dicts {u, c, ranked (item ids), scores, kept, correct (positive test items that are
        candidates)} in (u, c) order; the measures (Measures.PrecAt ..., :852-858) are left to the caller."""
        t, tr = self.testMatrix, self.trainMatrix
        if t is None:
            return []
        cand = java_hashset_order(tr.j)  # getItemList iterates the rows of trainMatrix (DataDAO.java:1213)
        cand_set = set(int(x) for x in cand)
        tc = t.get("ctx")
        tctx = np.zeros(len(t["u"]), dtype=np.int32) if tc is None else np.asarray(tc)
        pos: Dict[tuple, list] = {}
        for u, j, c, r in zip(t["u"].tolist(), t["j"].tolist(), tctx.tolist(), np.asarray(t["r"]).tolist()):
            if r > binThold:
                pos.setdefault((u, c), []).append(j)
        trctx = np.zeros(tr.nnz, dtype=np.int32) if tr.ctx is None else tr.ctx
        rated: Dict[tuple, list] = {}
        for u, j, c in zip(tr.u.tolist(), tr.j.tolist(), trctx.tolist()):
            if (u, c) in pos:
                rated.setdefault((u, c), []).append(j)
        queries = []
        for (u, c) in sorted(pos):
            correct = [j for j in pos[(u, c)] if j in cand_set]
            if correct:  # :781-782 `continue` when no positive item is a candidate
                queries.append((u, c, correct))
        if not queries:
            return []
        qu = np.array([q[0] for q in queries], dtype=np.int32)
        qc = np.array([q[1] for q in queries], dtype=np.int32)
        rptr = np.zeros(len(queries) + 1, dtype=np.int64)
        ritems = []
        for i, (u, c, _) in enumerate(queries):
            ritems.extend(rated.get((u, c), []))
            rptr[i + 1] = len(ritems)
        eng = self._eval_engine()
        try:
            items, scores, count, kept = eng.rank_topn(qu, None if tr.ctx is None else qc, cand, rptr,
                                                       np.asarray(ritems, dtype=np.int32), binThold, numRecs)
        finally:
            if eng is not self.engine:
                eng.close()
        out = []
        for i, (u, c, correct) in enumerate(queries):
            n = int(count[i])
            if n == 0:
                continue  # :817-818 no recommendations available
            out.append({"u": u, "c": c, "ranked": items[i, :n].tolist(), "scores": scores[i, :n].tolist(),
                        "kept": int(kept[i]), "correct": correct})
        return out

    def execute(self, init: Optional[Dict[str, np.ndarray]] = None, seed: int = 0) -> Dict[str, float]:
        """Recommender.execute (:319-357): initModel -> buildModel -> evalRatings."""
        self.initModel(init, seed)
        self.keep_engine = True
        try:
            self.buildModel()
            self.measures = self.evalRatings()
        finally:
            self.keep_engine = False
            if self.engine is not None:
                self.engine.close()
                self.engine = None
        return self.measures


class PMF(IterativeRecommender):
    MODEL, algoName = capi.PMF, "PMF"


class BiasedMF(IterativeRecommender):
    MODEL, algoName = capi.BIASEDMF, "BiasedMF"


class SVDPlusPlus(IterativeRecommender):
    """carskit.alg.baseline.cf.SVDPlusPlus (SVDPlusPlus.java): BiasedMF plus the implicit-feedback factors Y of every item the
    user rated.  Every rating rewrites Y[k] for all of the user's items, so the engine runs it as ONE chain (one warp):
    bit-identical to the reference, meant for the small data sets the dependency structure allows."""
    MODEL, algoName = capi.SVDPP, "SVD++"


class CAMF_C(IterativeRecommender):
    MODEL, algoName = capi.CAMF_C, "CAMF_C"


class CAMF_CI(IterativeRecommender):
    MODEL, algoName = capi.CAMF_CI, "CAMF_CI"


class CAMF_CU(IterativeRecommender):
    MODEL, algoName = capi.CAMF_CU, "CAMF_CU"


class CAMF_CUCI(IterativeRecommender):
    """CAMF_CUCI.java: item-context AND user-context deviations; its Guava tables hold one Gaussian cell per
    (user | item, condition) (:58-64), passed to the engine as dense arrays."""
    MODEL, algoName = capi.CAMF_CUCI, "CAMF_CUCI"
    GAUSSIAN_CONTEXT_BIAS = True


class CAMF_ICS(IterativeRecommender):
    """carskit.alg.cars.adaptation.dependent.sim.CAMF_ICS (CAMF_ICS.java): independent context similarity -- the rating is
    P[u].Q[j] times the learnt similarity of every condition to its dimension's "na" condition.  A top-N model (isRankingPred,
    :29): P, Q ~ U(0, 1), every similarity starts at 1 (:40-48).  EXACT mode, one warp (one chain through ccMatrix_ICS)."""
    MODEL, algoName = capi.CAMF_ICS, "CAMF_ICS"

    def initModel(self, init=None, seed: int = 0):
        if init is not None:
            return super().initModel(init, seed)
        rng = np.random.default_rng(seed)
        C = self.numConditions
        self.model = {"P": rng.random((self.numUsers, self.numFactors)), "Q": rng.random((self.numItems, self.numFactors)),
                      "cc_sim": np.ones((C, C))}


class CAMF_LCS(IterativeRecommender):
    """carskit.alg.cars.adaptation.dependent.sim.CAMF_LCS (CAMF_LCS.java): latent context similarity -- every condition has a
    vector of `-f` factors (default 10, :38), sim(condition, "na") = their dot product.  P, Q and the vectors ~ U(0, 1)."""
    MODEL, algoName = capi.CAMF_LCS, "CAMF_LCS"

    def __init__(self, trainMatrix, testMatrix=None, fold=-1, conf=None, device=0, stream=0, **kw):
        super().__init__(trainMatrix, testMatrix, fold, conf, device, stream, **kw)
        self.numContextFactors = int(LineConfiger(self.cf.get("CAMF_LCS", "-f 10")).getFloat("-f", 10))  # algoOptions.getInt("-f", 10)

    def initModel(self, init=None, seed: int = 0):
        if init is not None:
            return super().initModel(init, seed)
        rng = np.random.default_rng(seed)
        self.model = {"P": rng.random((self.numUsers, self.numFactors)), "Q": rng.random((self.numItems, self.numFactors)),
                      "cf_lcs": rng.random((self.numConditions, self.numContextFactors))}


class CAMF_MCS(IterativeRecommender):
    """carskit.alg.cars.adaptation.dependent.sim.CAMF_MCS (CAMF_MCS.java): multidimensional context similarity -- every
    condition is a position on its dimension's axis, sim = 1 - distance(context, all-"na" context); positions start in
    U(0, 1 / sqrt(numContextDims)) (:44-48) and are clamped there; the reference scales this model's loss by 0.05 (:158)."""
    MODEL, algoName = capi.CAMF_MCS, "CAMF_MCS"

    def initModel(self, init=None, seed: int = 0):
        if init is not None:
            return super().initModel(init, seed)
        rng = np.random.default_rng(seed)
        dims = max(1, len(self.trainMatrix.empty_conditions))
        self.model = {"P": rng.random((self.numUsers, self.numFactors)), "Q": rng.random((self.numItems, self.numFactors)),
                      "c_mcs": rng.random(self.numConditions) / np.sqrt(dims)}


class FM(IterativeRecommender):
    """carskit.alg.cars.adaptation.dependent.FM (FM.java): ALS factorization machine over the one-hot features
    (user, item, context).  `FM=-lw <f> -lf <f>` in the configuration (FM.java:53-54); learn.rate and
    isConverged() are not used by the reference's buildModel() (:148-219 runs exactly numIters iterations)."""
    MODEL, algoName = capi.FM, "FM"

    def __init__(self, trainMatrix, testMatrix=None, fold=-1, conf=None, device=0, stream=0, **kw):
        super().__init__(trainMatrix, testMatrix, fold, conf, device, stream, **kw)
        opts = LineConfiger(self.cf.get("FM", "-lw 0.01 -lf 0.02"))
        self.regLw, self.regLf = opts.getFloat("-lw", 0.0), opts.getFloat("-lf", 0.0)
        # rateDao.numContextDims(): one condition per dimension in every context (DataDAO.java:281-290)
        ts = trainMatrix
        # (TrainingSet.num_context_dims when the data came through the loader; the length of the first context otherwise)
        self.numContextDims = int(self.cf.get("num.context.dims", 0)) or int(getattr(ts, "num_context_dims", 0)) or (
            int(ts.ctx_ptr[1] - ts.ctx_ptr[0]) if ts.ctx_ptr is not None and len(ts.ctx_ptr) > 1 else 1)
        self.p = self.numUsers + self.numItems + self.numConditions
        self.k = self.numFactors
        self.globalSize = 0

    def initModel(self, init=None, seed: int = 0):
        """FM.initModel (FM.java:57-74): w0 = 0, w ~ U(0,1) (DenseVector.init()), V ~ N(0, 0.1)."""
        if init is not None:
            self.model = {"w0": np.ascontiguousarray(init["w0"], dtype=np.float64).reshape(1),
                          "w": np.ascontiguousarray(init["w"], dtype=np.float64),
                          "V": np.ascontiguousarray(init["V"], dtype=np.float64)}
            if self.model["w"].shape != (self.p,) or self.model["V"].shape != (self.p, self.k):
                raise ValueError("FM arrays: w must be [p], V [p x k]")
            return
        rng = np.random.default_rng(seed)
        self.model = {"w0": np.zeros(1), "w": rng.random(self.p),
                      "V": self.initMean + self.initStd * rng.standard_normal((self.p, self.k))}

    def _desc(self):
        return capi.make_desc(self.trainMatrix, capi.FM, self.k, device=self.device, reg_lw=self.regLw,
                              reg_lf=self.regLf, num_context_dims=self.numContextDims, stream=self.stream,
                              global_nnz=self.globalSize, tuning=self.tuning)

    def _new_engine(self):
        if self.world > 1:  # the coordinate-sum all-reduces must be stream-ordered with the kernels
            import torch
            if self.stream:
                self._torch_stream = torch.cuda.ExternalStream(self.stream, device=self.device)
            else:
                self._torch_stream = torch.cuda.Stream(device=self.device)
                self.stream = self._torch_stream.cuda_stream
        return capi.FmEngine(self._desc(), keepalive=self.trainMatrix)

    def open_engine(self):
        """world > 1: trainMatrix holds THIS rank's contiguous range of rows (sharding.shard_rows); w0, w and V
        are replicated; `size` in FM.java's denominators is the global row count."""
        if not self.model:
            raise RuntimeError("buildModel before initModel")
        self.globalSize = 0
        if self.world > 1:
            self.globalSize = int(self._allreduce_sum([float(self.trainMatrix.nnz)])[0])
        eng = self._new_engine()
        try:
            eng.upload(self.model)
            eng.prepare()  # FM.java:118-146
            if self.world > 1:
                from .sharding import CoordinateSumExchange
                self.exchange = CoordinateSumExchange(eng, self._exchange_device(), self.group, self._torch_stream)
        except Exception:
            eng.close()
            raise
        self.engine = eng
        return eng

    def train_epoch(self, iter: int) -> bool:
        self.loss = self.exchange.iteration(self.engine) if self.world > 1 else self.engine.iteration()
        self.iter_losses.append(self.loss)
        return False  # FM.java never calls isConverged()

    def _eval_engine(self):
        if self.engine is not None:
            return self.engine
        eng = capi.FmEngine(self._desc_for_predict(), keepalive=self.trainMatrix)
        eng.upload(self.model)
        return eng

    def _desc_for_predict(self):
        d = super()._desc_for_predict()
        return capi.make_desc(self._empty_keep, capi.FM, self.k, device=self.device, reg_lw=self.regLw,
                              reg_lf=self.regLf, num_context_dims=self.numContextDims)

    def evalRatings(self):
        t = self.testMatrix
        if t is None or len(t["u"]) == 0:
            return {"MAE": float("nan"), "RMSE": float("nan")}
        eng = self._eval_engine()
        try:
            pred = eng.predict(t["u"], t["j"], t["ctx"], bound=True, min_rate=self.minRate, max_rate=self.maxRate)
        finally:
            if eng is not self.engine:
                eng.close()
        err = np.abs(np.asarray(t["r"], dtype=np.float64) - pred)  # Recommender.java:518-545
        sa, ss = 0.0, 0.0
        for e in err.tolist():
            sa += e
            ss += e * e
        return {"MAE": sa / len(err), "RMSE": math.sqrt(ss / len(err))}


def runCrossValidation(rateMatrix: TrainingSet, name: str, conf: Optional[Dict[str, str]] = None, kFold: int = 5,
                       rand_seed: int = 1, parallel: bool = True, devices=(0,), inits=None):
    """CARSKit.runCrossValidation (src/carskit/main/CARSKit.java:387-423) over this path: `DataSplitter ds = new
    DataSplitter(rateMatrix, kFold)`, one recommender per fold on its own thread (`-p on`, the default of
    setting.conf:39; `-p off` joins each thread before the next starts), then the average of every measure,
    accumulated in fold order as `val + measure / kFold` (:415-421).

    Each fold owns a handle (the C ABI is re-entrant per handle) on `devices[fold % len(devices)]`; the ctypes
    calls release the GIL, so the folds' epochs overlap on the device like the Java threads overlap on the cores.
    `inits[i]` hands fold i+1 its initial model (the reference's own generator is wall-clock seeded); otherwise
    fold i+1 draws from seed `rand_seed + i + 1`.  Returns (average measures, the per-fold recommenders)."""
    import threading
    from .data import DataSplitter
    ds = DataSplitter(rateMatrix, kFold, rand_seed)
    Rec = getRecommender(name)
    algos, threads, errors = [], [], []

    def run(algo, i):
        try:
            algo.execute(init=None if inits is None else inits[i], seed=rand_seed + i + 1)
        except BaseException as e:  # surfaced after the join, like Recommender.run() logging a failed fold (:1162-1171)
            errors.append((i + 1, e))

    for i in range(ds.numFold):
        train, test = ds.getKthFold(i + 1)
        if Rec.MODEL in (capi.PMF, capi.BIASEDMF) and getattr(train, "pair_ids", None) is not None:
            # the 2-D models iterate `train = rateDao.toTraditionalSparseMatrix(trainMatrix)` (Recommender.java:252)
            from .data import to_traditional
            train = to_traditional(train)
        algo = Rec(train, test, fold=i + 1, conf=conf, device=devices[i % len(devices)])
        algos.append(algo)
        t = threading.Thread(target=run, args=(algo, i))
        threads.append(t)
        t.start()
        if not parallel:
            t.join()
    if parallel:
        for t in threads:
            t.join()
    if errors:
        raise RuntimeError(f"fold {errors[0][0]} failed: {errors[0][1]}") from errors[0][1]
    avg: Dict[str, float] = {}
    for algo in algos:
        for m, v in algo.measures.items():
            avg[m] = avg.get(m, 0.0) + v / ds.numFold
    return avg, algos


def getRecommender(name: str):
    """The `switch` of CARSKit.getRecommender (src/carskit/main/CARSKit.java:429-705) for this path."""
    table = {"pmf": PMF, "biasedmf": BiasedMF, "camf_c": CAMF_C, "camf_ci": CAMF_CI, "camf_cu": CAMF_CU, "fm": FM,
             "camf_cuci": CAMF_CUCI, "camf_ics": CAMF_ICS, "camf_lcs": CAMF_LCS, "camf_mcs": CAMF_MCS, "svdpp": SVDPlusPlus, "svd++": SVDPlusPlus}
    try:
        return table[name.lower()]
    except KeyError:
        raise ValueError(f"recommender '{name}' is not on the B200 hot path") from None
