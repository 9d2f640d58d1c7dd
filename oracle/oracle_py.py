"""ctypes binding of oracle/libcars_oracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (carskit_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from carskit_b200.capi import CarsDesc, CarsModelArrays, make_arrays, _ptr_f64, _ptr_i32

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcars_oracle.so")


class JRandom(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("next_next_gaussian", C.c_double), ("have_next_next_gaussian", C.c_int)]


class EpochState(C.Structure):
    _fields_ = [("lRate", C.c_double), ("loss", C.c_double), ("last_loss", C.c_double),
                ("measure", C.c_double), ("last_measure", C.c_double),
                ("maxLRate", C.c_float), ("decay", C.c_float),
                ("isBoldDriver", C.c_int32), ("early_stop", C.c_int32)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("cars_oracle.cpp", "fm_oracle.cpp")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "carskit_b200.h"))  # cars_desc is shared with the product
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    f64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.oracle_jr_seed.argtypes = [C.POINTER(JRandom), C.c_int64]
    L.oracle_jr_seed.restype = None
    L.oracle_jr_next_int.argtypes = [C.POINTER(JRandom)]
    L.oracle_jr_next_int.restype = C.c_int32
    L.oracle_jr_next_int_bound.argtypes = [C.POINTER(JRandom), C.c_int32]
    L.oracle_jr_next_int_bound.restype = C.c_int32
    L.oracle_jr_next_double.argtypes = [C.POINTER(JRandom)]
    L.oracle_jr_next_double.restype = C.c_double
    L.oracle_jr_next_gaussian.argtypes = [C.POINTER(JRandom)]
    L.oracle_jr_next_gaussian.restype = C.c_double
    L.oracle_init_gaussian.argtypes = [C.POINTER(JRandom), f64p, C.c_int64, C.c_double, C.c_double]
    L.oracle_init_gaussian.restype = None
    L.oracle_init_uniform.argtypes = [C.POINTER(JRandom), f64p, C.c_int64]
    L.oracle_init_uniform.restype = None
    L.oracle_epoch.argtypes = [C.POINTER(CarsDesc), C.POINTER(CarsModelArrays), C.c_double]
    L.oracle_epoch.restype = C.c_double
    L.oracle_predict.argtypes = [C.POINTER(CarsDesc), C.POINTER(CarsModelArrays), C.c_int64, i32p, i32p, i32p,
                                 C.c_int32, C.c_double, C.c_double, f64p]
    L.oracle_predict.restype = C.c_int
    L.oracle_eval_ratings.argtypes = [C.POINTER(CarsDesc), C.POINTER(CarsModelArrays), C.c_int64, i32p, i32p, i32p,
                                      f64p, C.c_double, C.c_double, f64p, f64p, C.POINTER(C.c_int64)]
    L.oracle_eval_ratings.restype = C.c_int
    L.oracle_update_lrate.argtypes = [C.POINTER(EpochState), C.c_int]
    L.oracle_update_lrate.restype = None
    L.oracle_is_converged.argtypes = [C.POINTER(EpochState), C.c_int]
    L.oracle_is_converged.restype = C.c_int
    L.oracle_build_model.argtypes = [C.POINTER(CarsDesc), C.POINTER(CarsModelArrays), C.POINTER(EpochState), C.c_int, f64p]
    L.oracle_build_model.restype = C.c_int
    L.oracle_global_mean.argtypes = [f64p, C.c_int64]
    L.oracle_global_mean.restype = C.c_double
    L.oracle_to_traditional.argtypes = [C.c_int64, i32p, f64p, C.c_int32, i32p, i32p, i32p, i32p, f64p]
    L.oracle_to_traditional.restype = C.c_int64
    L.oracle_selftest_no_fma.argtypes = []
    L.oracle_selftest_no_fma.restype = C.c_int
    L.oracle_version.argtypes = []
    L.oracle_version.restype = C.c_char_p
    _lib = L
    return L


class JavaRandom:
    """java.util.Random clone."""

    def __init__(self, seed: int):
        self.g = JRandom()
        lib().oracle_jr_seed(C.byref(self.g), seed)

    def next_int(self) -> int:
        return lib().oracle_jr_next_int(C.byref(self.g))

    def next_int_bound(self, bound: int) -> int:
        return lib().oracle_jr_next_int_bound(C.byref(self.g), bound)

    def next_double(self) -> float:
        return lib().oracle_jr_next_double(C.byref(self.g))

    def next_gaussian(self) -> float:
        return lib().oracle_jr_next_gaussian(C.byref(self.g))

    def gaussian(self, shape, mean=0.0, sigma=0.1) -> np.ndarray:
        a = np.empty(shape, dtype=np.float64)
        lib().oracle_init_gaussian(C.byref(self.g), _ptr_f64(a), a.size, mean, sigma)
        return a

    def uniform(self, shape) -> np.ndarray:
        a = np.empty(shape, dtype=np.float64)
        lib().oracle_init_uniform(C.byref(self.g), _ptr_f64(a), a.size)
        return a


def epoch(desc: CarsDesc, arrs: dict, lrate: float) -> float:
    a = make_arrays(arrs)
    loss = lib().oracle_epoch(C.byref(desc), C.byref(a), lrate)
    cc = arrs.get("cc_sim")
    if cc is not None:  # librec's SymmMatrix holds one cell per unordered pair (max, min): show it on both sides
        low = np.tril(cc)
        cc[...] = low + np.tril(cc, -1).T
    return loss


def predict(desc: CarsDesc, arrs: dict, u, j, ctx=None, bound=False, min_rate=0.0, max_rate=0.0) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.int32)
    j = np.ascontiguousarray(j, dtype=np.int32)
    ctx = None if ctx is None else np.ascontiguousarray(ctx, dtype=np.int32)
    out = np.empty(u.shape[0], dtype=np.float64)
    a = make_arrays(arrs)
    lib().oracle_predict(C.byref(desc), C.byref(a), u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                         1 if bound else 0, min_rate, max_rate, _ptr_f64(out))
    return out


def eval_ratings(desc: CarsDesc, arrs: dict, u, j, ctx, r, min_rate, max_rate):
    u = np.ascontiguousarray(u, dtype=np.int32)
    j = np.ascontiguousarray(j, dtype=np.int32)
    ctx = None if ctx is None else np.ascontiguousarray(ctx, dtype=np.int32)
    r = np.ascontiguousarray(r, dtype=np.float64)
    sa, ss, cnt = C.c_double(), C.c_double(), C.c_int64()
    a = make_arrays(arrs)
    lib().oracle_eval_ratings(C.byref(desc), C.byref(a), u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                              _ptr_f64(r), min_rate, max_rate, C.byref(sa), C.byref(ss), C.byref(cnt))
    return sa.value, ss.value, cnt.value


def rank_topn(desc: CarsDesc, arrs: dict, qu, qc, cand, rated_ptr, rated_items, bin_thold: float, num_recs: int):
    L = lib()
    i64p = C.POINTER(C.c_int64)
    L.oracle_rank_topn.argtypes = [C.POINTER(CarsDesc), C.POINTER(CarsModelArrays), C.c_int64, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), i64p, C.POINTER(C.c_int32),
                                   C.c_double, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.oracle_rank_topn.restype = C.c_int
    qu = np.ascontiguousarray(qu, dtype=np.int32)
    qc = None if qc is None else np.ascontiguousarray(qc, dtype=np.int32)
    cand = np.ascontiguousarray(cand, dtype=np.int32)
    nq = qu.shape[0]
    items = np.full((nq, num_recs), -1, dtype=np.int32)
    scores = np.zeros((nq, num_recs))
    count = np.zeros(nq, dtype=np.int32)
    kept = np.zeros(nq, dtype=np.int32)
    rp = None if rated_ptr is None else np.ascontiguousarray(rated_ptr, dtype=np.int64)
    ri = None if rated_items is None else np.ascontiguousarray(rated_items, dtype=np.int32)
    a = make_arrays(arrs)
    L.oracle_rank_topn(C.byref(desc), C.byref(a), nq, _ptr_i32(qu), _ptr_i32(qc), cand.shape[0], _ptr_i32(cand),
                       None if rp is None else rp.ctypes.data_as(i64p), _ptr_i32(ri), bin_thold, num_recs,
                       _ptr_i32(items), _ptr_f64(scores), _ptr_i32(count), _ptr_i32(kept))
    return items, scores, count, kept


def new_state(lrate: float, bold_driver=True, decay=-1.0, max_lrate=-1.0, early_stop=0) -> EpochState:
    s = EpochState()
    s.lRate = lrate
    s.loss = s.last_loss = s.measure = s.last_measure = 0.0
    s.maxLRate, s.decay = max_lrate, decay
    s.isBoldDriver = 1 if bold_driver else 0
    s.early_stop = early_stop
    return s


def build_model(desc: CarsDesc, arrs: dict, state: EpochState, num_iters: int):
    losses = np.zeros(num_iters, dtype=np.float64)
    a = make_arrays(arrs)
    n = lib().oracle_build_model(C.byref(desc), C.byref(a), C.byref(state), num_iters, _ptr_f64(losses))
    return n, losses[: max(n, 0)]


def global_mean(r: np.ndarray) -> float:
    r = np.ascontiguousarray(r, dtype=np.float64)
    return lib().oracle_global_mean(_ptr_f64(r), r.shape[0])


# ---------------------------------------------------------------------------------------------------
# FM (fm_oracle.cpp)
# ---------------------------------------------------------------------------------------------------
class FmProblem(C.Structure):
    _fields_ = [("num_users", C.c_int32), ("num_items", C.c_int32), ("num_conditions", C.c_int32),
                ("num_context_dims", C.c_int32), ("k", C.c_int32), ("pad", C.c_int32), ("size", C.c_int64),
                ("u", C.POINTER(C.c_int32)), ("j", C.POINTER(C.c_int32)), ("ctx", C.POINTER(C.c_int32)),
                ("r", C.POINTER(C.c_double)), ("reg_lw", C.c_float), ("reg_lf", C.c_float)]


class FmModel(C.Structure):
    _fields_ = [("w0", C.POINTER(C.c_double)), ("w", C.POINTER(C.c_double)), ("V", C.POINTER(C.c_double))]


_fm_ready = False


def _fm_lib():
    global _fm_ready
    L = lib()
    if not _fm_ready:
        f64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        P, M = C.POINTER(FmProblem), C.POINTER(FmModel)
        L.oracle_fm_dense_predict.argtypes = [P, M, C.c_int, C.c_int, C.c_int]
        L.oracle_fm_dense_predict.restype = C.c_double
        L.oracle_fm_dense_build.argtypes = [P, M, C.c_int, f64p, f64p, f64p]
        L.oracle_fm_dense_build.restype = C.c_double
        L.oracle_fm_predict.argtypes = [P, M, C.c_int, C.c_int, C.c_int]
        L.oracle_fm_predict.restype = C.c_double
        L.oracle_fm_predict_batch.argtypes = [P, M, C.c_int64, i32p, i32p, i32p, C.c_int32, C.c_double, C.c_double, f64p]
        L.oracle_fm_predict_batch.restype = C.c_int
        L.oracle_fm_prepare.argtypes = [P, M, f64p, f64p]
        L.oracle_fm_prepare.restype = None
        L.oracle_fm_iteration.argtypes = [P, M, f64p, f64p, C.c_int]
        L.oracle_fm_iteration.restype = C.c_double
        _fm_ready = True
    return L


def fm_problem(ts, k: int, num_context_dims: int, reg_lw: float, reg_lf: float) -> FmProblem:
    """ts: carskit_b200.capi.TrainingSet (kept alive by the caller)."""
    p = FmProblem()
    p.num_users, p.num_items, p.num_conditions = ts.num_users, ts.num_items, ts.num_conditions
    p.num_context_dims, p.k, p.size = num_context_dims, k, ts.nnz
    p.u, p.j, p.ctx, p.r = _ptr_i32(ts.u), _ptr_i32(ts.j), _ptr_i32(ts.ctx), _ptr_f64(ts.r)
    p.reg_lw, p.reg_lf = reg_lw, reg_lf
    return p


def _fm_model(arrs: dict) -> FmModel:
    m = FmModel()
    m.w0, m.w, m.V = _ptr_f64(arrs["w0"]), _ptr_f64(arrs["w"]), _ptr_f64(arrs["V"])
    return m


def fm_dense_build(prob: FmProblem, arrs: dict, num_iters: int):
    """Literal FM.buildModel(); returns (loss, errors, Q)."""
    p = prob.num_users + prob.num_items + prob.num_conditions
    errors = np.zeros(prob.size)
    Q = np.zeros((prob.size, prob.k))
    fv = np.zeros((prob.size, p))
    m = _fm_model(arrs)
    loss = _fm_lib().oracle_fm_dense_build(C.byref(prob), C.byref(m), num_iters, _ptr_f64(errors), _ptr_f64(Q), _ptr_f64(fv))
    return loss, errors, Q


def fm_dense_predict(prob: FmProblem, arrs: dict, u: int, j: int, c: int) -> float:
    m = _fm_model(arrs)
    return _fm_lib().oracle_fm_dense_predict(C.byref(prob), C.byref(m), u, j, c)


def fm_predict(prob: FmProblem, arrs: dict, u, j, ctx, bound=False, lo=0.0, hi=0.0) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.int32)
    j = np.ascontiguousarray(j, dtype=np.int32)
    ctx = np.ascontiguousarray(ctx, dtype=np.int32)
    out = np.empty(u.shape[0])
    m = _fm_model(arrs)
    _fm_lib().oracle_fm_predict_batch(C.byref(prob), C.byref(m), u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                                      1 if bound else 0, lo, hi, _ptr_f64(out))
    return out


def fm_prepare(prob: FmProblem, arrs: dict):
    errors = np.zeros(prob.size)
    Q = np.zeros((prob.size, prob.k))
    m = _fm_model(arrs)
    _fm_lib().oracle_fm_prepare(C.byref(prob), C.byref(m), _ptr_f64(errors), _ptr_f64(Q))
    return errors, Q


def fm_iteration(prob: FmProblem, arrs: dict, errors: np.ndarray, Q: np.ndarray, closed_den: bool) -> float:
    m = _fm_model(arrs)
    return _fm_lib().oracle_fm_iteration(C.byref(prob), C.byref(m), _ptr_f64(errors), _ptr_f64(Q), 1 if closed_den else 0)
