"""Epoch time of the one-chain models (one warp, EXACT): CAMF_C, CAMF_ICS, CAMF_LCS, CAMF_MCS on a Frappe-shaped set, SVD++ on a
DePaulMovie-sized and a MovieLens-100K-sized set.  Prints one line per model (profiles/r2/one_chain_models.txt)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carskit_b200 import capi, synth  # noqa: E402

REGS = dict(reg_u=capi.f32(1e-4), reg_i=capi.f32(1e-4), reg_b=capi.f32(1e-4), reg_c=capi.f32(1e-3))


def run(name, ts, F, epochs=5):
    model = capi.MODEL_NAMES[name]
    rng = np.random.default_rng(1)
    shapes = capi.member_shapes(model, ts.num_users, ts.num_items, ts.num_conditions, F)
    arrs = {k: 0.1 * rng.random(s) for k, s in shapes.items()}
    if "cc_sim" in arrs:
        arrs["cc_sim"][...] = 1.0
    if "cf_lcs" in arrs:
        arrs["cf_lcs"][...] = 0.3 + 0.05 * rng.random(arrs["cf_lcs"].shape)
    with capi.Engine(capi.make_desc(ts, model, F, **REGS), keepalive=ts) as eng:
        eng.upload(arrs)
        eng.epoch(1e-4)
        t0 = time.perf_counter()
        for _ in range(epochs):
            loss = eng.epoch(1e-4)
        dt = (time.perf_counter() - t0) / epochs
        st = eng.stats()
    print(f"{name:9s} F={F:3d} nnz={ts.nnz:7d}  {dt * 1e3:8.2f} ms/epoch  {ts.nnz / dt / 1e6:6.2f} M updates/s  "
          f"kernel {st.last_epoch_ms:8.2f} ms  loss {loss:.6g}", flush=True)


frappe, _ = synth.make_training_set(957, 4082, [7, 7, 2, 3, 2, 9, 80, 233], 96203, seed=1)
for name in ("camf_c", "camf_ics", "camf_lcs", "camf_mcs"):
    run(name, frappe, 10)
depaul, _ = synth.make_training_set(97, 79, None, 1443, seed=2)
run("svdpp", depaul, 10)
ml100k, _ = synth.make_training_set(943, 1682, None, 100000, seed=3)
run("svdpp", ml100k, 10)
run("biasedmf", ml100k, 10)
