"""Runs the REFERENCE'S OWN BYTECODE for the hot path on the mini JVM -- TEST INFRASTRUCTURE (parity pin).

    carskit/alg/**/{PMF,BiasedMF,CAMF_C,CAMF_CI,CAMF_CU,CAMF_CUCI}.buildModel()/predict()   jar/CARSKit-v0.4.0.jar
    carskit/generic/{IterativeRecommender.isConverged,updateLRate, Recommender.predict(IIIZ)}  "
    librec/data/{DenseMatrix.get/set/add/rowMult, DenseVector.get/add, SparseMatrix.iterator,
                 SparseMatrix$MatrixIterator, SparseMatrix$SparseMatrixEntry}                 lib/librec-v1.4-alpha.jar

are all INTERPRETED from the class files under /root/reference (never copied into this repo).  Supplied by this
harness, because they are containers / IO rather than arithmetic:  DataDAO.getUserIdFromUI / getItemIdFromUI
(HashMap<Integer,Integer> lookups, DataDAO.java:1038-1046), ContextRecommender.getConditions (BiMap lookup +
String.split, ContextRecommender.java:53-61 -- returns the condition ids in header-column order), Guava's
HashBasedTable get / put (CAMF_CUCI), boxed Integer / Double, java.util.List / Iterator, Math.abs / Math.pow.

The model arrays go in and come out as numpy arrays with the shapes of carskit_b200.capi.member_shapes, so the output
can be compared directly with the CPU oracle and with the CUDA engine.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np

from carskit_b200 import capi

from .classfile import Jar
from .minijvm import F32, HostIterator, JObject, MiniJVM

REFERENCE = "/root/reference"
JARS = [os.path.join(REFERENCE, "jar", "CARSKit-v0.4.0.jar"), os.path.join(REFERENCE, "lib", "librec-v1.4-alpha.jar")]
DEV = "carskit/alg/cars/adaptation/dependent/dev/"
CLASS_OF = {capi.PMF: "carskit/alg/baseline/cf/PMF", capi.BIASEDMF: "carskit/alg/baseline/cf/BiasedMF",
            capi.SVDPP: "carskit/alg/baseline/cf/SVDPlusPlus",
            capi.CAMF_C: DEV + "CAMF_C", capi.CAMF_CI: DEV + "CAMF_CI", capi.CAMF_CU: DEV + "CAMF_CU",
            capi.CAMF_CUCI: DEV + "CAMF_CUCI", capi.CAMF_ICS: "carskit/alg/cars/adaptation/dependent/sim/CAMF_ICS",
            capi.CAMF_LCS: "carskit/alg/cars/adaptation/dependent/sim/CAMF_LCS",
            capi.CAMF_MCS: "carskit/alg/cars/adaptation/dependent/sim/CAMF_MCS"}
REC = "carskit/generic/Recommender"
ITER = "carskit/generic/IterativeRecommender"


def available() -> bool:
    return all(os.path.exists(p) for p in JARS)


class HostTable:
    """com.google.common.collect.HashBasedTable<Integer, Integer, Double>: get / put / contains / size / rowKeySet / row(r).keySet()
    (CAMF_CUCI.java:58-64, 96-112; CAMF_ICS.java:76-105; librec SymmMatrix).  rowKeySet order = insertion order: the cells a
    rating updates are distinct, so the order Guava's HashMap would yield does not change any value."""
    jclass = "com/google/common/collect/HashBasedTable"

    def __init__(self, dense: Optional[np.ndarray] = None, lower_triangle: bool = False):
        self.d = {}
        if dense is not None:
            for r in range(dense.shape[0]):
                for c in range(dense.shape[1]):
                    if not lower_triangle or r >= c:
                        self.d[(r, c)] = float(dense[r, c])

    def row_keys(self):
        return list(dict.fromkeys(r for r, _ in self.d))

    def row(self, r):
        return HostRow([c for (rr, c) in self.d if rr == r])

    def to_dense(self, shape) -> np.ndarray:
        out = np.zeros(shape)
        for (r, c), v in self.d.items():
            out[r, c] = v
        return out


class HostRow:
    jclass = "java/util/Map"

    def __init__(self, keys):
        self.keys = keys


def dense_matrix(a: np.ndarray) -> JObject:
    return JObject("librec/data/DenseMatrix", data=[[float(x) for x in row] for row in a], numRows=int(a.shape[0]),
                   numColumns=int(a.shape[1]))


def dense_vector(a: np.ndarray) -> JObject:
    return JObject("librec/data/DenseVector", data=[float(x) for x in a], size=int(a.shape[0]))


def crs_matrix(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, num_rows: int, num_cols: int) -> JObject:
    """librec.data.SparseMatrix in CRS form: entries arrive sorted by (row, col)."""
    row_ptr = np.zeros(num_rows + 1, dtype=np.int64)
    np.add.at(row_ptr, rows.astype(np.int64) + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return JObject("carskit/data/structure/SparseMatrix", rowPtr=[int(x) for x in row_ptr], colInd=[int(x) for x in cols],
                   rowData=[float(x) for x in vals], numRows=int(num_rows), numColumns=int(num_cols))


class ReferenceRun:
    def __init__(self, model: int, ts: capi.TrainingSet, arrays: Dict[str, np.ndarray], F: int, *, lrate: float = 0.02,
                 reg_u: float = 1e-4, reg_i: float = 1e-4, reg_b: float = 1e-4, reg_c: float = 1e-3, bold_driver: bool = True,
                 decay: float = -1.0, max_lrate: float = -1.0, min_rate: float = 1.0, max_rate: float = 5.0):
        self.model, self.ts, self.F = model, ts, F
        self.jar = Jar(*JARS)
        self.jvm = jvm = MiniJVM(self.jar)
        self.cls = CLASS_OF[model]
        has_ctx = model not in (capi.PMF, capi.BIASEDMF, capi.SVDPP)
        # ---- trainMatrix / train ------------------------------------------------------------------------------
        if has_ctx:
            # rows = user-item pair ids in first-appearance order of the CRS stream (entries of one pair are adjacent)
            new_pair = np.ones(ts.nnz, dtype=bool)
            if ts.nnz > 1:
                new_pair[1:] = (ts.u[1:] != ts.u[:-1]) | (ts.j[1:] != ts.j[:-1])
            ui = np.cumsum(new_pair) - 1
            self.ui_user = [int(x) for x in ts.u[new_pair]]
            self.ui_item = [int(x) for x in ts.j[new_pair]]
            mat = crs_matrix(ui, ts.ctx, ts.r, int(ui[-1]) + 1 if ts.nnz else 0, ts.num_contexts)
        else:
            mat = crs_matrix(ts.u, ts.j, ts.r, ts.num_users, ts.num_items)
        self.conds: List[List[int]] = []
        if has_ctx:
            for c in range(ts.num_contexts):
                self.conds.append([int(x) for x in ts.ctx_cond[ts.ctx_ptr[c]:ts.ctx_ptr[c + 1]]])
        # ---- natives --------------------------------------------------------------------------------------------
        dao = JObject("carskit/data/processor/DataDAO")
        jvm.natives[("carskit/data/processor/DataDAO", "getUserIdFromUI")] = lambda vm, a: self.ui_user[a[1]]
        jvm.natives[("carskit/data/processor/DataDAO", "getItemIdFromUI")] = lambda vm, a: self.ui_item[a[1]]
        jvm.natives[("carskit/generic/ContextRecommender", "getConditions")] = lambda vm, a: self.conds[a[1]]
        for tcls in ("com/google/common/collect/Table", "com/google/common/collect/HashBasedTable"):
            jvm.natives[(tcls, "get")] = lambda vm, a: a[0].d.get((a[1], a[2]))
            jvm.natives[(tcls, "put")] = self._table_put
            jvm.natives[(tcls, "contains")] = lambda vm, a: int((a[1], a[2]) in a[0].d)
            jvm.natives[(tcls, "size")] = lambda vm, a: len(a[0].d)
            jvm.natives[(tcls, "rowKeySet")] = lambda vm, a: a[0].row_keys()
            jvm.natives[(tcls, "row")] = lambda vm, a: a[0].row(a[1])
            jvm.natives[(tcls, "create")] = lambda vm, a: HostTable()
        jvm.natives[("java/util/Map", "keySet")] = lambda vm, a: a[0].keys
        jvm.natives[("java/util/Set", "iterator")] = lambda vm, a: HostIterator(a[0])
        jvm.natives[("java/util/ArrayList", "get")] = lambda vm, a: a[0][a[1]]
        # ---- statics (IterativeRecommender.java:36-49, Recommender.java:196-204) ---------------------------------
        for name, v in (("regU", reg_u), ("regI", reg_i), ("regB", reg_b), ("regC", reg_c), ("decay", decay),
                        ("maxLRate", max_lrate), ("initLRate", lrate)):
            jvm.set_static(ITER, name, F32(v))
        jvm.set_static(ITER, "numFactors", int(F))
        jvm.set_static(ITER, "isBoldDriver", int(bold_driver))
        jvm.set_static(REC, "rateDao", dao)
        jvm.set_static(REC, "verbose", 0)
        jvm.set_static(REC, "earlyStopMeasure", None)
        jvm.set_static(REC, "minRate", float(min_rate))
        jvm.set_static(REC, "maxRate", float(max_rate))
        jvm.set_static(REC, "numUsers", ts.num_users)
        jvm.set_static(REC, "numItems", ts.num_items)
        # ---- the recommender instance ------------------------------------------------------------------------------
        rec = JObject(self.cls)
        rec.f["trainMatrix" if has_ctx else "train"] = mat
        rec.f["globalMean"] = float(ts.global_mean)
        rec.f["lRate"] = float(F32(lrate))  # `lRate = initLRate` (IterativeRecommender.java:106): float widened to double
        rec.f["loss"] = 0.0
        rec.f["last_loss"] = 0.0
        rec.f["measure"] = 0.0
        rec.f["last_measure"] = 0.0
        rec.f["isUserSplitting"] = 0
        rec.f["isItemSplitting"] = 0
        rec.f["P"] = dense_matrix(arrays["P"])
        rec.f["Q"] = dense_matrix(arrays["Q"])
        if "user_bias" in arrays:
            rec.f["userBias"] = dense_vector(arrays["user_bias"])
        if "item_bias" in arrays:
            rec.f["itemBias"] = dense_vector(arrays["item_bias"])
        if "cond_bias" in arrays:
            rec.f["condBias"] = dense_vector(arrays["cond_bias"])
        if model == capi.CAMF_ICS:
            # ccMatrix_ICS: librec SymmMatrix (interpreted) over a Guava table holding the (max, min) cells (CAMF_ICS.java:45-48)
            rec.f["ccMatrix_ICS"] = JObject("librec/data/SymmMatrix", dim=int(ts.num_conditions),
                                            data=HostTable(arrays["cc_sim"], lower_triangle=True))
            jvm.set_static("carskit/generic/ContextRecommender", "EmptyContextConditions", [int(x) for x in ts.empty_conditions])
        if model == capi.SVDPP:  # SVDPlusPlus.java:45-52: Y, and userItemsCache = train.rowColumnsCache(): getColumns(u), ascending
            rec.f["Y"] = dense_matrix(arrays["Y"])
            items = [sorted(int(x) for x in ts.j[ts.u == uu]) for uu in range(ts.num_users)]
            rec.f["userItemsCache"] = JObject("com/google/common/cache/LoadingCache", items=items)
            jvm.natives[("com/google/common/cache/LoadingCache", "get")] = lambda vm, a: a[0].f["items"][a[1]]
        if model == capi.CAMF_LCS:  # CAMF_LCS.java:36-41: numF = `-f`, cfMatrix_LCS [numConditions x numF]
            rec.f["cfMatrix_LCS"] = dense_matrix(arrays["cf_lcs"])
            rec.f["numF"] = int(arrays["cf_lcs"].shape[1])
            jvm.set_static("carskit/generic/ContextRecommender", "EmptyContextConditions", [int(x) for x in ts.empty_conditions])
        if model == capi.CAMF_MCS:  # CAMF_MCS.java:39-49: upbound = 1 / sqrt(numContextDims), lowbound = 1 / 10^100
            import math
            rec.f["cVector_MCS"] = dense_vector(arrays["c_mcs"])
            rec.f["upbound"] = 1.0 / math.sqrt(len(ts.empty_conditions))
            rec.f["lowbound"] = 1.0 / math.pow(10, 100)
            jvm.set_static("carskit/generic/ContextRecommender", "EmptyContextConditions", [int(x) for x in ts.empty_conditions])
        if model == capi.CAMF_CUCI:
            rec.f["icBias"] = HostTable(arrays["ic_bias"])
            rec.f["ucBias"] = HostTable(arrays["uc_bias"])
        else:
            if "ic_bias" in arrays:
                rec.f["icBias"] = dense_matrix(arrays["ic_bias"])
            if "uc_bias" in arrays:
                rec.f["ucBias"] = dense_matrix(arrays["uc_bias"])
        self.rec = rec
        self.shapes = {k: v.shape for k, v in arrays.items()}
        self.losses: List[float] = []
        self.lrates: List[float] = []

    @staticmethod
    def _table_put(vm, a):
        old = a[0].d.get((a[1], a[2]))
        a[0].d[(a[1], a[2])] = a[3]
        return old

    # ---- execution -----------------------------------------------------------------------------------------------
    def build_model(self, num_iters: int):
        """buildModel() as javac compiled it, isConverged()/updateLRate() included; records loss and lRate per iteration
        through a hook on isConverged (which the bytecode calls once per iteration)."""
        jvm = self.jvm
        jvm.set_static(ITER, "numIters", int(num_iters))
        cf, m = self.jar.find_method(ITER, "isConverged", "(I)Z")

        def hooked(vm, a):
            self.losses.append(a[0].f["loss"])
            self.lrates.append(a[0].f["lRate"])  # the rate this iteration USED
            return vm.run(cf, m, a)

        jvm.natives[(self.cls, "isConverged")] = hooked
        jvm.call_virtual(self.rec, self.cls, "buildModel", "()V", [])
        return self

    def predict(self, u, j, c, bound: bool) -> np.ndarray:
        out = np.empty(len(u))
        for k in range(len(u)):
            out[k] = self.jvm.call_virtual(self.rec, self.cls, "predict", "(IIIZ)D",
                                           [int(u[k]), int(j[k]), int(c[k]) if c is not None else 0, int(bound)])
        return out

    def arrays(self) -> Dict[str, np.ndarray]:
        f = self.rec.f
        out = {"P": np.array(f["P"].f["data"], dtype=np.float64).reshape(self.shapes["P"]),
               "Q": np.array(f["Q"].f["data"], dtype=np.float64).reshape(self.shapes["Q"])}
        for key, field in (("user_bias", "userBias"), ("item_bias", "itemBias"), ("cond_bias", "condBias")):
            if key in self.shapes:
                out[key] = np.array(f[field].f["data"], dtype=np.float64)
        if "cc_sim" in self.shapes:
            C = self.shapes["cc_sim"][0]
            cc = np.zeros((C, C))
            for (r, c), v in f["ccMatrix_ICS"].f["data"].d.items():
                cc[r, c] = cc[c, r] = v
            out["cc_sim"] = cc
        if "Y" in self.shapes:
            out["Y"] = np.array(f["Y"].f["data"], dtype=np.float64).reshape(self.shapes["Y"])
        if "cf_lcs" in self.shapes:
            out["cf_lcs"] = np.array(f["cfMatrix_LCS"].f["data"], dtype=np.float64).reshape(self.shapes["cf_lcs"])
        if "c_mcs" in self.shapes:
            out["c_mcs"] = np.array(f["cVector_MCS"].f["data"], dtype=np.float64)
        for key, field in (("ic_bias", "icBias"), ("uc_bias", "ucBias")):
            if key in self.shapes:
                o = f[field]
                out[key] = o.to_dense(self.shapes[key]) if isinstance(o, HostTable) else \
                    np.array(o.f["data"], dtype=np.float64).reshape(self.shapes[key])
        return out


class ReferenceFM:
    """carskit/alg/cars/adaptation/dependent/FM.buildModel() / predict() from the reference's bytecode (ALS with cached
    residuals, FM.java:115-220).  Natives: HashBasedTable create / put / get (the `fvalues` cache, FM.java:121-142),
    DataDAO lookups, numContextDims()."""
    CLS = "carskit/alg/cars/adaptation/dependent/FM"

    def __init__(self, ts: capi.TrainingSet, arrays: Dict[str, np.ndarray], k: int, num_context_dims: int, reg_lw: float,
                 reg_lf: float, min_rate: float = 1.0, max_rate: float = 5.0):
        self.ts, self.k = ts, k
        self.jar = Jar(*JARS)
        self.jvm = jvm = MiniJVM(self.jar)
        new_pair = np.ones(ts.nnz, dtype=bool)
        if ts.nnz > 1:
            new_pair[1:] = (ts.u[1:] != ts.u[:-1]) | (ts.j[1:] != ts.j[:-1])
        ui = np.cumsum(new_pair) - 1
        ui_user, ui_item = [int(x) for x in ts.u[new_pair]], [int(x) for x in ts.j[new_pair]]
        mat = crs_matrix(ui, ts.ctx, ts.r, int(ui[-1]) + 1 if ts.nnz else 0, ts.num_contexts)
        dao = JObject("carskit/data/processor/DataDAO")
        jvm.natives[("carskit/data/processor/DataDAO", "getUserIdFromUI")] = lambda vm, a: ui_user[a[1]]
        jvm.natives[("carskit/data/processor/DataDAO", "getItemIdFromUI")] = lambda vm, a: ui_item[a[1]]
        jvm.natives[("carskit/data/processor/DataDAO", "numContextDims")] = lambda vm, a: int(num_context_dims)

        class Table:
            jclass = "com/google/common/collect/HashBasedTable"

            def __init__(self):
                self.d = {}
        jvm.natives[("com/google/common/collect/HashBasedTable", "create")] = lambda vm, a: Table()
        jvm.natives[("com/google/common/collect/HashBasedTable", "put")] = lambda vm, a: a[0].d.__setitem__((a[1], a[2]), a[3])
        jvm.natives[("com/google/common/collect/HashBasedTable", "get")] = lambda vm, a: a[0].d.get((a[1], a[2]))
        jvm.set_static(REC, "rateDao", dao)
        jvm.set_static(REC, "minRate", float(min_rate))
        jvm.set_static(REC, "maxRate", float(max_rate))
        jvm.set_static("carskit/generic/ContextRecommender", "numConditions", int(ts.num_conditions))
        p = ts.num_users + ts.num_items + ts.num_conditions
        rec = JObject(self.CLS, trainMatrix=mat, numUsers=int(ts.num_users), numItems=int(ts.num_items), p=int(p), k=int(k),
                      size=int(ts.nnz), w0=float(arrays["w0"][0]), regLw=F32(reg_lw), regLf=F32(reg_lf), loss=0.0,
                      globalMean=float(ts.global_mean))
        w = dense_vector(arrays["w"]); w.cls = "carskit/data/structure/DenseVector"
        V = dense_matrix(arrays["V"]); V.cls = "carskit/data/structure/DenseMatrix"
        Q = dense_matrix(np.zeros((ts.nnz, k))); Q.cls = "carskit/data/structure/DenseMatrix"
        rec.f.update(w=w, V=V, Q=Q)
        self.rec, self.p = rec, p

    def build_model(self, num_iters: int):
        self.jvm.set_static(ITER, "numIters", int(num_iters))
        self.jvm.call_virtual(self.rec, self.CLS, "buildModel", "()V", [])
        return self

    def arrays(self):
        f = self.rec.f
        return {"w0": np.array([f["w0"]]), "w": np.array(f["w"].f["data"]), "V": np.array(f["V"].f["data"]).reshape(self.p, self.k)}

    def predict(self, u, j, c, bound: bool) -> np.ndarray:
        return np.array([self.jvm.call_virtual(self.rec, self.CLS, "predict", "(IIIZ)D", [int(a), int(b), int(d), int(bound)])
                         for a, b, d in zip(u, j, c)])
