"""K5 / K6 measurement: cars_predict on 10 M test ratings and cars_rank_topn on Q queries x 100 K
candidate items (numRecs 10) against a CAMF_CI F = 64 model of BASELINE config 3's shape.  Prints one JSON line per
call with the wall time through the C ABI (host buffers, copies included); the per-kernel device times and DRAM bytes
come from running this script under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
(profiles/r2/launches_eval_kernels.txt), which is what the rooflines in DESIGN.md divide by.

    python scripts/bench_eval_kernels.py [num_predict] [num_queries]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carskit_b200 import capi, synth  # noqa: E402

n_pred = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
U, I, F, dims = 1_000_000, 100_000, 64, [8, 8, 8, 8]
ts, _ = synth.make_training_set(U, I, dims, 1000, seed=1)  # the model is what matters; the training set is a stub
rng = np.random.default_rng(7)
arrs = {"P": 0.1 * rng.standard_normal((U, F)), "Q": 0.1 * rng.standard_normal((I, F)), "user_bias": 0.1 * rng.standard_normal(U),
        "ic_bias": rng.random((I, ts.num_conditions))}
desc = capi.make_desc(ts, capi.CAMF_CI, F, reg_u=1e-4, reg_i=1e-4, reg_b=1e-4, reg_c=1e-3)
with capi.Engine(desc, keepalive=ts) as eng:
    eng.upload(arrs)
    u = rng.integers(0, U, n_pred, dtype=np.int32)
    j = rng.integers(0, I, n_pred, dtype=np.int32)
    c = rng.integers(0, ts.num_contexts, n_pred, dtype=np.int32)
    eng.predict(u[:1000], j[:1000], c[:1000], bound=True, min_rate=1.0, max_rate=5.0)
    samples = []
    for _ in range(3):  # the first full-size call of a process also grows the device-memory pool: median of three, all listed
        t0 = time.perf_counter()
        out = eng.predict(u, j, c, bound=True, min_rate=1.0, max_rate=5.0)
        samples.append(time.perf_counter() - t0)
    dt = sorted(samples)[1]
    B = 2 * F * 8 + 12 + 8 + 8 + 4 * 8
    print(json.dumps({"call": "cars_predict", "queries": n_pred, "seconds_e2e_host_buffers": dt, "seconds_samples": samples,
                      "queries_per_s_e2e": n_pred / dt, "algorithmic_bytes_per_query": B, "checksum": float(out[:1000].sum())}))
    qu = rng.integers(0, U, n_q, dtype=np.int32)
    qc = rng.integers(0, ts.num_contexts, n_q, dtype=np.int32)
    cand = rng.permutation(I).astype(np.int32)
    rptr = np.arange(n_q + 1, dtype=np.int64) * 20
    ritems = rng.integers(0, I, 20 * n_q, dtype=np.int32)
    eng.rank_topn(qu[:64], qc[:64], cand, rptr[:65], ritems[:64 * 20], -1.0, 10)
    samples = []
    for _ in range(3):
        t0 = time.perf_counter()
        items, scores, count, kept = eng.rank_topn(qu, qc, cand, rptr, ritems, -1.0, 10)
        samples.append(time.perf_counter() - t0)
    dt = sorted(samples)[1]
    print(json.dumps({"call": "cars_rank_topn", "queries": n_q, "candidates": I, "num_recs": 10, "seconds_e2e_host_buffers": dt,
                      "seconds_samples": samples, "scores_per_s_e2e": n_q * I / dt, "flop_per_score": 2 * F, "key_bytes_per_score": 16,
                      "checksum": int(items[:100].sum()), "kept_mean": float(kept.mean())}))
