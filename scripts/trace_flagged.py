"""Developer tool: where does one turn of the flagged kernel spend its time?
Needs the -DCARS_TRACE build (carskit_b200/libcarskit_b200_trace.so); run on the GPU box:
  CARSKIT_B200_LIB=carskit_b200/libcarskit_b200_trace.so python scripts/trace_flagged.py <workload> <out.json>"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from carskit_b200 import capi, recommender

wl_name = sys.argv[1]
wl = bench.WORKLOADS[wl_name]
ts, model, arrs = bench.make_inputs(wl, 0)
win = 400_000
lo = ts.nnz // 2
os.environ["CARS_TRACE_WINDOW"] = f"{lo}:{win}"
os.environ["CARS_TRACE_OUT"] = "/tmp/trace.bin"
rec = recommender.getRecommender(wl["model"])(ts, None, conf={"num.factors": str(wl["F"]), "num.max.iter": "3"})
rec.initModel(init=arrs)
eng = rec.open_engine()
for it in range(3):
    rec.train_epoch(it + 1)
st = eng.stats()
ms = st.last_epoch_ms
rec.close_engine()
raw = np.fromfile("/tmp/trace.bin", dtype=np.uint64)
t = raw[: win * 8].reshape(win, 8).astype(np.int64)
recs = raw[win * 8:].view(np.int32).reshape(win, 8)  # u j ctx ku kj pad r(2)
u, j, ku, kj = recs[:, 0], recs[:, 1], recs[:, 3], recs[:, 4]
valid = t[:, 5] > 0
clk = 1.965  # GHz
def stat(x):
    x = x[valid]
    return {"mean_us": float(x.mean() / clk / 1e3), "p50_us": float(np.median(x) / clk / 1e3), "p90_us": float(np.percentile(x, 90) / clk / 1e3)}
out = {"workload": wl_name, "epoch_ms": ms, "levels": int(st.num_levels), "groups": int(st.grid_ctas * st.block_threads // 32 * 8),
       "poll_turn": stat(t[:, 0]), "gather": stat(t[:, 1]), "compute_scatter": stat(t[:, 2]), "release": stat(t[:, 3]),
       "tries_mean": float(t[valid, 6].mean()), "tries_p90": float(np.percentile(t[valid, 6], 90))}
# link latency: successor's poll-success time minus predecessor's release time (globaltimer, ns)
key_j = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(j, kj)) if valid[i]}
key_u = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(u, ku)) if valid[i]}
links = []
for i in np.nonzero(valid)[0][:200000]:
    pj = key_j.get((int(j[i]), int(kj[i]) - 1))
    pu = key_u.get((int(u[i]), int(ku[i]) - 1))
    rel = max(t[pj, 5] if pj is not None else 0, t[pu, 5] if pu is not None else 0)
    if rel > 0:
        links.append(t[i, 4] - rel)
links = np.array(links, dtype=np.float64)
if len(links):
    out["link_ns"] = {"n": int(len(links)), "mean": float(links.mean()), "p10": float(np.percentile(links, 10)),
                      "p50": float(np.median(links)), "p90": float(np.percentile(links, 90)),
                      "frac_below_2us": float((links < 2000).mean())}
turn = (t[valid, 5] - t[valid, 4])
out["ready_to_release_ns"] = {"mean": float(turn.mean()), "p50": float(np.median(turn))}
print(json.dumps(out, indent=1))
json.dump(out, open(sys.argv[2], "w"), indent=1)
