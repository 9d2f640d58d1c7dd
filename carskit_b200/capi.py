"""ctypes binding of include/carskit_b200.h (the C ABI a JNI shim would bind).

Nothing here computes: it marshals numpy arrays into the plain-pointer ABI and raises when the CUDA
library is missing or a call fails.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

ABI_VERSION = 3
PMF, BIASEDMF, CAMF_C, CAMF_CI, CAMF_CU, FM, CAMF_CUCI, CAMF_ICS, CAMF_LCS, CAMF_MCS, SVDPP = range(11)
EXACT, FAST = 0, 1
SCHED_FLAGGED, SCHED_WAVEFRONT, SCHED_DATAFLOW = 0, 1, 2
COMBINE_MEAN, COMBINE_SUM, COMBINE_TOUCHED = 0, 1, 2
MODEL_NAMES = {"pmf": PMF, "biasedmf": BIASEDMF, "camf_c": CAMF_C, "camf_ci": CAMF_CI, "camf_cu": CAMF_CU, "fm": FM,
               "camf_cuci": CAMF_CUCI, "camf_ics": CAMF_ICS, "camf_lcs": CAMF_LCS, "camf_mcs": CAMF_MCS, "svdpp": SVDPP}

_HERE = os.path.dirname(os.path.abspath(__file__))
# CARSKIT_B200_LIB selects another build of the same ABI (e.g. the developer build with stage tracing)
LIB_PATH = os.environ.get("CARSKIT_B200_LIB") or os.path.join(_HERE, "libcarskit_b200.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class CarsDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("mode", C.c_int32), ("device", C.c_int32),
        ("num_users", C.c_int32), ("num_items", C.c_int32), ("num_conditions", C.c_int32),
        ("num_contexts", C.c_int32), ("num_factors", C.c_int32), ("schedule", C.c_int32),
        ("nnz", C.c_int64),
        ("u", _i32p), ("j", _i32p), ("ctx", _i32p), ("r", _f64p), ("ctx_ptr", _i32p), ("ctx_cond", _i32p),
        ("global_mean", C.c_double),
        ("reg_u", C.c_double), ("reg_i", C.c_double), ("reg_b", C.c_double), ("reg_c", C.c_double),
        ("reg_lw", C.c_double), ("reg_lf", C.c_double),
        ("num_context_dims", C.c_int32), ("num_gpus", C.c_int32), ("global_nnz", C.c_int64),
        ("stream", C.c_void_p), ("gpu_ids", _i32p), ("fast_max_conc", C.c_double), ("tuning", C.c_char_p),
        ("combine", C.c_int32), ("num_empty_conditions", C.c_int32), ("empty_conditions", _i32p),
        ("num_context_factors", C.c_int32),
    ]


class CarsModelArrays(C.Structure):
    _fields_ = [(n, _f64p) for n in ("P", "Q", "user_bias", "item_bias", "cond_bias", "ic_bias", "uc_bias", "cc_sim", "cf_lcs",
                                     "c_mcs", "Y")]


class CarsStats(C.Structure):
    _fields_ = [
        ("nnz", C.c_int64), ("num_levels", C.c_int64), ("max_level_size", C.c_int64),
        ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("schedule_ms", C.c_double), ("last_epoch_ms", C.c_double),
        ("grid_ctas", C.c_int32), ("block_threads", C.c_int32), ("sm_count", C.c_int32), ("reserved", C.c_int32),
        ("schedule_copy_ms", C.c_double), ("schedule_levels_ms", C.c_double), ("schedule_pack_ms", C.c_double),
        ("fast_min_item_scale", C.c_double), ("fast_min_cond_scale", C.c_double), ("max_item_degree", C.c_int64),
        ("fast_hot_rows", C.c_int32), ("num_gpus", C.c_int32), ("exchange_ms", C.c_double),
    ]


# every symbol include/carskit_b200.h declares
EXPORTS = [
    "cars_create", "cars_upload", "cars_epoch", "cars_epoch_begin", "cars_epoch_wait", "cars_download",
    "cars_predict", "cars_eval_ratings", "cars_destroy", "cars_last_error", "cars_get_stats",
    "cars_get_stream", "cars_version", "cars_item_block_doubles", "cars_epoch_sharded_begin",
    "cars_epoch_sharded_finish", "cars_fm_create", "cars_fm_upload", "cars_fm_prepare", "cars_fm_iteration",
    "cars_fm_download", "cars_fm_predict", "cars_fm_get_stats", "cars_fm_last_error", "cars_fm_destroy",
    "cars_fm_exchange_doubles", "cars_fm_iteration_sharded", "cars_fm_get_stream", "cars_rank_topn",
    "cars_device_count", "cars_dataset_read_binary_csv", "cars_dataset_from_arrays", "cars_dataset_save", "cars_dataset_load",
    "cars_dataset_kfold", "cars_dataset_get_view", "cars_dataset_free", "cars_dataset_last_error",
]

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)


class CarsFmArrays(C.Structure):
    _fields_ = [("w0", _f64p), ("w", _f64p), ("V", _f64p)]


class CarsFmStats(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("p", C.c_int64), ("pieces", C.c_int64), ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("last_iteration_ms", C.c_double)]


class CarsDatasetView(C.Structure):
    _fields_ = [("num_users", C.c_int32), ("num_items", C.c_int32), ("num_pairs", C.c_int32), ("num_contexts", C.c_int32),
                ("num_conditions", C.c_int32), ("num_context_dims", C.c_int32), ("nnz", C.c_int64),
                ("u", _i32p), ("j", _i32p), ("ctx", _i32p), ("pair", _i32p), ("r", _f64p), ("ctx_ptr", _i32p), ("ctx_cond", _i32p),
                ("global_mean", C.c_double), ("min_rate", C.c_double), ("max_rate", C.c_double),
                ("num_empty_conditions", C.c_int32), ("empty_conditions", _i32p)]


class CarsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"carskit_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). carskit_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    H = C.c_void_p
    lib.cars_create.argtypes = [C.POINTER(CarsDesc), C.POINTER(H)]
    lib.cars_create.restype = C.c_int
    lib.cars_upload.argtypes = [H, C.POINTER(CarsModelArrays)]
    lib.cars_upload.restype = C.c_int
    lib.cars_download.argtypes = [H, C.POINTER(CarsModelArrays)]
    lib.cars_download.restype = C.c_int
    lib.cars_epoch.argtypes = [H, C.c_double, _f64p]
    lib.cars_epoch.restype = C.c_int
    lib.cars_epoch_begin.argtypes = [H, C.c_double]
    lib.cars_epoch_begin.restype = C.c_int
    lib.cars_epoch_wait.argtypes = [H, _f64p]
    lib.cars_epoch_wait.restype = C.c_int
    lib.cars_predict.argtypes = [H, C.c_int64, _i32p, _i32p, _i32p, C.c_int32, C.c_double, C.c_double, _f64p]
    lib.cars_predict.restype = C.c_int
    lib.cars_eval_ratings.argtypes = [H, C.c_int64, _i32p, _i32p, _i32p, _f64p, C.c_double, C.c_double, _f64p, _f64p]
    lib.cars_eval_ratings.restype = C.c_int
    lib.cars_destroy.argtypes = [H]
    lib.cars_destroy.restype = None
    lib.cars_last_error.argtypes = [H]
    lib.cars_last_error.restype = C.c_char_p
    lib.cars_get_stats.argtypes = [H, C.POINTER(CarsStats)]
    lib.cars_get_stats.restype = C.c_int
    lib.cars_get_stream.argtypes = [H]
    lib.cars_get_stream.restype = C.c_void_p
    lib.cars_rank_topn.argtypes = [H, C.c_int64, _i32p, _i32p, C.c_int32, _i32p, C.POINTER(C.c_int64), _i32p, C.c_double,
                                   C.c_int32, _i32p, _f64p, _i32p, _i32p]
    lib.cars_rank_topn.restype = C.c_int
    lib.cars_item_block_doubles.argtypes = [H, C.POINTER(C.c_int64)]
    lib.cars_item_block_doubles.restype = C.c_int
    lib.cars_epoch_sharded_begin.argtypes = [H, C.c_double, C.c_void_p]
    lib.cars_epoch_sharded_begin.restype = C.c_int
    lib.cars_epoch_sharded_finish.argtypes = [H, C.c_void_p, C.c_double, _f64p]
    lib.cars_epoch_sharded_finish.restype = C.c_int
    lib.cars_fm_create.argtypes = [C.POINTER(CarsDesc), C.POINTER(H)]
    lib.cars_fm_create.restype = C.c_int
    lib.cars_fm_upload.argtypes = [H, C.POINTER(CarsFmArrays)]
    lib.cars_fm_upload.restype = C.c_int
    lib.cars_fm_download.argtypes = [H, C.POINTER(CarsFmArrays)]
    lib.cars_fm_download.restype = C.c_int
    lib.cars_fm_prepare.argtypes = [H]
    lib.cars_fm_prepare.restype = C.c_int
    lib.cars_fm_iteration.argtypes = [H, _f64p]
    lib.cars_fm_iteration.restype = C.c_int
    lib.cars_fm_predict.argtypes = [H, C.c_int64, _i32p, _i32p, _i32p, C.c_int32, C.c_double, C.c_double, _f64p]
    lib.cars_fm_predict.restype = C.c_int
    lib.cars_fm_get_stats.argtypes = [H, C.POINTER(CarsFmStats)]
    lib.cars_fm_get_stats.restype = C.c_int
    lib.cars_fm_last_error.argtypes = [H]
    lib.cars_fm_last_error.restype = C.c_char_p
    lib.cars_fm_exchange_doubles.argtypes = [H, C.POINTER(C.c_int64)]
    lib.cars_fm_exchange_doubles.restype = C.c_int
    lib.cars_fm_iteration_sharded.argtypes = [H, C.c_void_p, ALLREDUCE_FN, C.c_void_p, _f64p]
    lib.cars_fm_iteration_sharded.restype = C.c_int
    lib.cars_fm_get_stream.argtypes = [H]
    lib.cars_fm_get_stream.restype = C.c_void_p
    lib.cars_fm_destroy.argtypes = [H]
    lib.cars_fm_destroy.restype = None
    lib.cars_version.argtypes = []
    lib.cars_version.restype = C.c_char_p
    lib.cars_device_count.argtypes = []
    lib.cars_device_count.restype = C.c_int
    lib.cars_dataset_read_binary_csv.argtypes = [C.c_char_p, C.POINTER(H)]
    lib.cars_dataset_read_binary_csv.restype = C.c_int
    lib.cars_dataset_from_arrays.argtypes = [C.c_int32] * 5 + [C.c_int64, _i32p, _i32p, _i32p, _f64p, _i32p, _i32p, C.POINTER(H)]
    lib.cars_dataset_from_arrays.restype = C.c_int
    lib.cars_dataset_save.argtypes = [H, C.c_char_p]
    lib.cars_dataset_save.restype = C.c_int
    lib.cars_dataset_load.argtypes = [C.c_char_p, C.POINTER(H)]
    lib.cars_dataset_load.restype = C.c_int
    lib.cars_dataset_kfold.argtypes = [H, C.c_int32, C.c_int64, C.c_int32, C.POINTER(H), C.POINTER(H)]
    lib.cars_dataset_kfold.restype = C.c_int
    lib.cars_dataset_get_view.argtypes = [H, C.POINTER(CarsDatasetView)]
    lib.cars_dataset_get_view.restype = C.c_int
    lib.cars_dataset_free.argtypes = [H]
    lib.cars_dataset_free.restype = None
    lib.cars_dataset_last_error.argtypes = []
    lib.cars_dataset_last_error.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib


def _ptr_i32(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i32p)


def _ptr_f64(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_f64p)


def f32(x: float) -> float:
    """Java `float` field widened to double: 1e-4f -> 9.999999747378752e-05 (IterativeRecommender.java:40)."""
    return float(np.float32(x))


@dataclass
class TrainingSet:
    """Flattened view of what buildModel() iterates: ratings in the reference's iteration order
    (CRS order of trainMatrix, CAMF_CI.java:80) plus the context -> conditions table
    (rateDao.getContextConditionsList(), DataDAO.java:1035) in CSR form."""
    num_users: int
    num_items: int
    u: np.ndarray
    j: np.ndarray
    r: np.ndarray
    ctx: Optional[np.ndarray] = None
    num_conditions: int = 0
    num_contexts: int = 0
    ctx_ptr: Optional[np.ndarray] = None
    ctx_cond: Optional[np.ndarray] = None
    global_mean: float = 0.0
    # rateDao.getRatingScale() of the WHOLE data set -> (minRate, maxRate) (Recommender.java:196-200) and
    # rateDao.numContextDims() (FM.java:86); None / 0 when the arrays did not come through the loader
    rating_scale: Optional[tuple] = None
    num_context_dims: int = 0
    # rateDao.getEmptyContextConditions(): the "dim:na" condition id of every context dimension (CAMF_ICS)
    empty_conditions: Optional[np.ndarray] = None

    def __post_init__(self):
        self.u = np.ascontiguousarray(self.u, dtype=np.int32)
        self.j = np.ascontiguousarray(self.j, dtype=np.int32)
        self.r = np.ascontiguousarray(self.r, dtype=np.float64)
        if self.ctx is not None:
            self.ctx = np.ascontiguousarray(self.ctx, dtype=np.int32)
            self.ctx_ptr = np.ascontiguousarray(self.ctx_ptr, dtype=np.int32)
            self.ctx_cond = np.ascontiguousarray(self.ctx_cond, dtype=np.int32)

    @property
    def nnz(self) -> int:
        return int(self.u.shape[0])


def make_desc(ts: TrainingSet, model: int, num_factors: int, *, mode: int = EXACT, device: int = 0,
              reg_u: float = 0.0, reg_i: float = 0.0, reg_b: float = 0.0, reg_c: float = 0.0,
              reg_lw: float = 0.0, reg_lf: float = 0.0, num_context_dims: int = 0,
              stream: int = 0, schedule: int = SCHED_FLAGGED, global_nnz: int = 0, fast_max_conc: float = 0.0,
              tuning: Optional[str] = None, gpu_ids=None, combine: int = COMBINE_MEAN, num_context_factors: int = 10) -> CarsDesc:
    """Fill a cars_desc.  The reg_* values must already be float-widened (use f32()).
    `tuning`: developer knobs "key=value;..." (csrc/tuning.h); `gpu_ids`: N > 1 CUDA ordinals for ONE handle that
    drives N GPUs from this process (users sharded by range, item block combined with NCCL inside cars_epoch)."""
    d = CarsDesc()
    d.fast_max_conc = fast_max_conc
    d.tuning = tuning.encode() if tuning else None
    d.combine = combine
    if gpu_ids is not None and len(gpu_ids) > 1:
        ids = np.ascontiguousarray(gpu_ids, dtype=np.int32)
        d._gpu_ids_keep = ids  # the descriptor must keep the array alive
        d.gpu_ids = _ptr_i32(ids)
        d.num_gpus = len(ids)
    d.abi_version = ABI_VERSION
    d.model, d.mode, d.device = model, mode, device
    d.schedule = schedule
    d.num_users, d.num_items = ts.num_users, ts.num_items
    d.num_conditions, d.num_contexts = ts.num_conditions, ts.num_contexts
    d.num_factors = num_factors
    d.nnz = ts.nnz
    d.u, d.j, d.r = _ptr_i32(ts.u), _ptr_i32(ts.j), _ptr_f64(ts.r)
    use_ctx = model in (CAMF_C, CAMF_CI, CAMF_CU, CAMF_CUCI, FM, CAMF_ICS, CAMF_LCS, CAMF_MCS) and ts.ctx is not None
    if model in (CAMF_ICS, CAMF_LCS, CAMF_MCS):
        ec = getattr(ts, "empty_conditions", None)
        if ec is None:
            raise ValueError("CAMF_ICS / LCS / MCS need TrainingSet.empty_conditions (rateDao.getEmptyContextConditions())")
        ec = np.ascontiguousarray(ec, dtype=np.int32)
        d._empty_keep = ec
        d.empty_conditions = _ptr_i32(ec)
        d.num_empty_conditions = len(ec)
        d.num_context_factors = num_context_factors  # CAMF_LCS `-f`
        if not num_context_dims:  # CAMF_MCS: upbound = 1 / sqrt(rateDao.numContextDims())
            num_context_dims = int(getattr(ts, "num_context_dims", 0)) or len(ec)
    d.ctx = _ptr_i32(ts.ctx) if use_ctx else None
    d.ctx_ptr = _ptr_i32(ts.ctx_ptr) if use_ctx else None
    d.ctx_cond = _ptr_i32(ts.ctx_cond) if use_ctx else None
    d.global_mean = ts.global_mean
    d.reg_u, d.reg_i, d.reg_b, d.reg_c, d.reg_lw, d.reg_lf = reg_u, reg_i, reg_b, reg_c, reg_lw, reg_lf
    d.num_context_dims = num_context_dims
    d.global_nnz = global_nnz
    d.stream = stream or None
    return d


MODEL_MEMBERS = {
    PMF: ("P", "Q"),
    BIASEDMF: ("P", "Q", "user_bias", "item_bias"),
    CAMF_C: ("P", "Q", "user_bias", "item_bias", "cond_bias"),
    CAMF_CI: ("P", "Q", "user_bias", "ic_bias"),
    CAMF_CU: ("P", "Q", "item_bias", "uc_bias"),
    CAMF_CUCI: ("P", "Q", "ic_bias", "uc_bias"),
    CAMF_ICS: ("P", "Q", "cc_sim"),
    CAMF_LCS: ("P", "Q", "cf_lcs"),
    CAMF_MCS: ("P", "Q", "c_mcs"),
    SVDPP: ("P", "Q", "user_bias", "item_bias", "Y"),
}


def member_shapes(model: int, num_users: int, num_items: int, num_conditions: int, F: int, num_context_factors: int = 10):
    all_shapes = {
        "P": (num_users, F), "Q": (num_items, F), "user_bias": (num_users,), "item_bias": (num_items,),
        "cond_bias": (num_conditions,), "ic_bias": (num_items, num_conditions),
        "uc_bias": (num_users, num_conditions), "cc_sim": (num_conditions, num_conditions),
        "cf_lcs": (num_conditions, num_context_factors), "c_mcs": (num_conditions,), "Y": (num_items, F),
    }
    return {k: all_shapes[k] for k in MODEL_MEMBERS[model]}


def make_arrays(arrs: dict) -> CarsModelArrays:
    a = CarsModelArrays()
    for name, _ in CarsModelArrays._fields_:
        v = arrs.get(name)
        setattr(a, name, _ptr_f64(v) if v is not None else None)
    return a


class Engine:
    """Thin RAII wrapper: cars_create .. cars_destroy."""

    def __init__(self, desc: CarsDesc, keepalive=None):
        self.lib = load_library()
        self._keep = keepalive
        self.h = C.c_void_p()
        rc = self.lib.cars_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise CarsError(rc, self.lib.cars_last_error(None).decode())
        self.model = desc.model

    def _check(self, rc: int):
        if rc != 0:
            raise CarsError(rc, self.lib.cars_last_error(self.h).decode())

    def upload(self, arrs: dict):
        a = make_arrays(arrs)
        self._check(self.lib.cars_upload(self.h, C.byref(a)))

    def download(self, arrs: dict):
        a = make_arrays(arrs)
        self._check(self.lib.cars_download(self.h, C.byref(a)))

    def epoch(self, lrate: float) -> float:
        loss = C.c_double()
        self._check(self.lib.cars_epoch(self.h, lrate, C.byref(loss)))
        return loss.value

    def epoch_begin(self, lrate: float):
        self._check(self.lib.cars_epoch_begin(self.h, lrate))

    def epoch_wait(self) -> float:
        loss = C.c_double()
        self._check(self.lib.cars_epoch_wait(self.h, C.byref(loss)))
        return loss.value

    def item_block_doubles(self) -> int:
        n = C.c_int64()
        self._check(self.lib.cars_item_block_doubles(self.h, C.byref(n)))
        return n.value

    def epoch_sharded_begin(self, lrate: float, dev_delta_ptr: int):
        """dev_delta_ptr: device address of item_block_doubles() doubles owned by the caller."""
        self._check(self.lib.cars_epoch_sharded_begin(self.h, lrate, C.c_void_p(dev_delta_ptr)))

    def epoch_sharded_finish(self, dev_delta_ptr: int, scale: float = 1.0) -> float:
        loss = C.c_double()
        self._check(self.lib.cars_epoch_sharded_finish(self.h, C.c_void_p(dev_delta_ptr), scale, C.byref(loss)))
        return loss.value

    def predict(self, u, j, ctx=None, bound=False, min_rate=0.0, max_rate=0.0) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        ctx = None if ctx is None else np.ascontiguousarray(ctx, dtype=np.int32)
        out = np.empty(u.shape[0], dtype=np.float64)
        self._check(self.lib.cars_predict(self.h, u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                                          1 if bound else 0, min_rate, max_rate, _ptr_f64(out)))
        return out

    def rank_topn(self, qu, qc, cand, rated_ptr=None, rated_items=None, bin_thold: float = -1.0, num_recs: int = 10):
        """evalRankings' scoring + top-N cut for a batch of (user, context) queries; see cars_rank_topn.
        Returns (items [nq x num_recs], scores, count, kept)."""
        qu = np.ascontiguousarray(qu, dtype=np.int32)
        qc = None if qc is None else np.ascontiguousarray(qc, dtype=np.int32)
        cand = np.ascontiguousarray(cand, dtype=np.int32)
        nq = qu.shape[0]
        items = np.full((nq, num_recs), -1, dtype=np.int32)
        scores = np.zeros((nq, num_recs), dtype=np.float64)
        count = np.zeros(nq, dtype=np.int32)
        kept = np.zeros(nq, dtype=np.int32)
        rp = None if rated_ptr is None else np.ascontiguousarray(rated_ptr, dtype=np.int64)
        ri = None if rated_items is None else np.ascontiguousarray(rated_items, dtype=np.int32)
        self._check(self.lib.cars_rank_topn(
            self.h, nq, _ptr_i32(qu), _ptr_i32(qc), cand.shape[0], _ptr_i32(cand),
            None if rp is None else rp.ctypes.data_as(C.POINTER(C.c_int64)), _ptr_i32(ri), bin_thold, num_recs,
            _ptr_i32(items), _ptr_f64(scores), _ptr_i32(count), _ptr_i32(kept)))
        return items, scores, count, kept

    def eval_ratings(self, u, j, ctx, r, min_rate, max_rate):
        u = np.ascontiguousarray(u, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        ctx = None if ctx is None else np.ascontiguousarray(ctx, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64)
        sa, ss = C.c_double(), C.c_double()
        self._check(self.lib.cars_eval_ratings(self.h, u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                                               _ptr_f64(r), min_rate, max_rate, C.byref(sa), C.byref(ss)))
        return sa.value, ss.value

    def stats(self) -> CarsStats:
        s = CarsStats()
        self._check(self.lib.cars_get_stats(self.h, C.byref(s)))
        return s

    def stream(self) -> int:
        return int(self.lib.cars_get_stream(self.h) or 0)

    def close(self):
        if self.h:
            self.lib.cars_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_fm_arrays(arrs: dict) -> CarsFmArrays:
    a = CarsFmArrays()
    a.w0, a.w, a.V = _ptr_f64(arrs["w0"]), _ptr_f64(arrs["w"]), _ptr_f64(arrs["V"])
    return a


class FmEngine:
    """cars_fm_create .. cars_fm_destroy (FM.java, ALS)."""

    def __init__(self, desc: CarsDesc, keepalive=None):
        self.lib = load_library()
        self._keep = keepalive
        self.h = C.c_void_p()
        rc = self.lib.cars_fm_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise CarsError(rc, self.lib.cars_fm_last_error(None).decode())

    def _check(self, rc: int):
        if rc != 0:
            raise CarsError(rc, self.lib.cars_fm_last_error(self.h).decode())

    def upload(self, arrs: dict):
        a = make_fm_arrays(arrs)
        self._check(self.lib.cars_fm_upload(self.h, C.byref(a)))

    def download(self, arrs: dict):
        a = make_fm_arrays(arrs)
        self._check(self.lib.cars_fm_download(self.h, C.byref(a)))

    def prepare(self):
        self._check(self.lib.cars_fm_prepare(self.h))

    def iteration(self) -> float:
        loss = C.c_double()
        self._check(self.lib.cars_fm_iteration(self.h, C.byref(loss)))
        return loss.value

    def exchange_doubles(self) -> int:
        n = C.c_int64()
        self._check(self.lib.cars_fm_exchange_doubles(self.h, C.byref(n)))
        return n.value

    def iteration_sharded(self, dev_buf_ptr: int, allreduce) -> float:
        """allreduce(dev_ptr: int, count: int) must sum `count` doubles at device address dev_ptr over the
        ranks, in place, ordered on the handle's stream; exceptions it raises fail the iteration."""
        err = []

        def _cb(user, ptr, count):
            try:
                allreduce(int(ptr), int(count))
                return 0
            except Exception as ex:  # noqa: BLE001 -- reported through the return code
                err.append(ex)
                return 1

        cb = ALLREDUCE_FN(_cb)
        loss = C.c_double()
        rc = self.lib.cars_fm_iteration_sharded(self.h, C.c_void_p(dev_buf_ptr), cb, None, C.byref(loss))
        if err:
            raise err[0]
        self._check(rc)
        return loss.value

    def stream(self) -> int:
        return int(self.lib.cars_fm_get_stream(self.h) or 0)

    def predict(self, u, j, ctx, bound=False, min_rate=0.0, max_rate=0.0) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        ctx = np.ascontiguousarray(ctx, dtype=np.int32)
        out = np.empty(u.shape[0], dtype=np.float64)
        self._check(self.lib.cars_fm_predict(self.h, u.shape[0], _ptr_i32(u), _ptr_i32(j), _ptr_i32(ctx),
                                             1 if bound else 0, min_rate, max_rate, _ptr_f64(out)))
        return out

    def stats(self) -> CarsFmStats:
        s = CarsFmStats()
        self._check(self.lib.cars_fm_get_stats(self.h, C.byref(s)))
        return s

    def close(self):
        if self.h:
            self.lib.cars_fm_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dataset:
    """cars_dataset: ratings held by the native library (csrc/ingest.cpp) -- DataDAO.readData, DataSplitter and the
    columnar file without Python loops.  `training_set()` copies the columns into a TrainingSet for cars_desc."""

    def __init__(self, handle):
        self.lib = load_library()
        self.h = handle

    @classmethod
    def _wrap(cls, lib, rc, h):
        if rc != 0:
            raise CarsError(rc, lib.cars_dataset_last_error().decode())
        return cls(h)

    @classmethod
    def read_binary_csv(cls, path: str) -> "Dataset":
        lib, h = load_library(), C.c_void_p()
        return cls._wrap(lib, lib.cars_dataset_read_binary_csv(path.encode(), C.byref(h)), h)

    @classmethod
    def load(cls, path: str) -> "Dataset":
        lib, h = load_library(), C.c_void_p()
        return cls._wrap(lib, lib.cars_dataset_load(path.encode(), C.byref(h)), h)

    @classmethod
    def from_training_set(cls, ts: TrainingSet, num_context_dims: int = 0) -> "Dataset":
        lib, h = load_library(), C.c_void_p()
        rc = lib.cars_dataset_from_arrays(ts.num_users, ts.num_items, ts.num_conditions, ts.num_contexts, num_context_dims, ts.nnz,
                                          _ptr_i32(ts.u), _ptr_i32(ts.j), _ptr_i32(ts.ctx), _ptr_f64(ts.r), _ptr_i32(ts.ctx_ptr),
                                          _ptr_i32(ts.ctx_cond), C.byref(h))
        return cls._wrap(lib, rc, h)

    def save(self, path: str):
        rc = self.lib.cars_dataset_save(self.h, path.encode())
        if rc != 0:
            raise CarsError(rc, self.lib.cars_dataset_last_error().decode())

    def kfold(self, k: int, seed: int, fold: int):
        tr, te = C.c_void_p(), C.c_void_p()
        rc = self.lib.cars_dataset_kfold(self.h, k, seed, fold, C.byref(tr), C.byref(te))
        if rc != 0:
            raise CarsError(rc, self.lib.cars_dataset_last_error().decode())
        return Dataset(tr), Dataset(te)

    def view(self) -> CarsDatasetView:
        v = CarsDatasetView()
        self.lib.cars_dataset_get_view(self.h, C.byref(v))
        return v

    def training_set(self) -> TrainingSet:
        v = self.view()
        n = v.nnz

        def arr(p, count, dt):
            return np.ctypeslib.as_array(p, shape=(count,)).astype(dt, copy=True) if count and p else np.empty(0, dt)
        has_ctx = bool(v.ctx)
        ts = TrainingSet(num_users=v.num_users, num_items=v.num_items, u=arr(v.u, n, np.int32), j=arr(v.j, n, np.int32),
                         r=arr(v.r, n, np.float64), ctx=arr(v.ctx, n, np.int32) if has_ctx else None,
                         num_conditions=v.num_conditions, num_contexts=v.num_contexts,
                         ctx_ptr=arr(v.ctx_ptr, v.num_contexts + 1, np.int32) if has_ctx else None,
                         ctx_cond=arr(v.ctx_cond, int(v.ctx_ptr[v.num_contexts]) if v.num_contexts else 0, np.int32) if has_ctx else None,
                         global_mean=v.global_mean, rating_scale=(v.min_rate, v.max_rate), num_context_dims=v.num_context_dims)
        ts.pair_ids = arr(v.pair, n, np.int64)
        return ts

    def close(self):
        if self.h:
            self.lib.cars_dataset_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
