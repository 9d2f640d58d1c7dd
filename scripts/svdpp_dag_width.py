"""Measured evidence for DESIGN.md "Out of scope: SVD++" (VERDICT r1 item 9): how wide is the serial-equivalent
dependency DAG of SVDPlusPlus.buildModel() (src/carskit/alg/baseline/cf/SVDPlusPlus.java:55-124) on the reference's own
data sets?

A rating (u, j) reads and writes P[u], Q[j], userBias[u], itemBias[j] AND Y[k] for EVERY item k the user rated
(:107-118).  Two ratings therefore conflict when they share the user, the item, or when their users share ANY rated
item.  level(n) = 1 + the largest level of an earlier conflicting rating (reference order: the 2-D `train` matrix, user
ascending, item ascending); average width = ratings / levels is the parallelism a serial-equivalent schedule could use.
For comparison the same computation for BiasedMF's conflict rule (same user or same item), which is what EXACT mode runs.

    python scripts/svdpp_dag_width.py            # needs /root/reference/context-aware_data_sets/*.zip
"""
import io
import os
import sys
import zipfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SETS = "/root/reference/context-aware_data_sets"


def pairs_depaul():
    z = zipfile.ZipFile(os.path.join(SETS, "Movie_DePaulMovie.zip"))
    name = [n for n in z.namelist() if n.endswith("ratings.txt")][0]
    rows = z.read(name).decode("utf-8", "replace").strip().splitlines()[1:]
    return [(r.split(",")[0], r.split(",")[1]) for r in rows if r.strip()]


def pairs_frappe():
    z = zipfile.ZipFile(os.path.join(SETS, "Mobile_Frappe.zip"))
    name = [n for n in z.namelist() if n.endswith("frappe.csv")][0]
    rows = z.read(name).decode("utf-8", "replace").strip().splitlines()[1:]
    return [(r.split("\t")[0], r.split("\t")[1]) for r in rows if r.strip()]


def widths(pairs):
    users = {u: i for i, u in enumerate(dict.fromkeys(p[0] for p in pairs))}
    items = {j: i for i, j in enumerate(dict.fromkeys(p[1] for p in pairs))}
    uj = sorted({(users[u], items[j]) for u, j in pairs})  # the 2-D train matrix in CRS order
    rated = {}
    for u, j in uj:
        rated.setdefault(u, []).append(j)
    out = {}
    for rule in ("biasedmf", "svdpp"):
        last_u, last_j, last_y = {}, {}, {}
        levels = 0
        for u, j in uj:
            lvl = max(last_u.get(u, 0), last_j.get(j, 0))
            if rule == "svdpp":
                lvl = max([lvl] + [last_y.get(k, 0) for k in rated[u]])
            lvl += 1
            last_u[u] = last_j[j] = lvl
            if rule == "svdpp":
                for k in rated[u]:
                    last_y[k] = lvl
            levels = max(levels, lvl)
        out[rule] = (len(uj), levels, len(uj) / levels)
    return len(users), len(items), out


if __name__ == "__main__":
    print(f"{'data set':14s} {'users':>6s} {'items':>6s} {'ratings':>8s} | {'BiasedMF levels':>15s} {'width':>7s} | {'SVD++ levels':>13s} {'width':>7s}")
    for name, fn in (("DePaulMovie", pairs_depaul), ("Frappe", pairs_frappe)):
        nu, ni, w = widths(fn())
        b, s = w["biasedmf"], w["svdpp"]
        print(f"{name:14s} {nu:6d} {ni:6d} {b[0]:8d} | {b[1]:15d} {b[2]:7.2f} | {s[1]:13d} {s[2]:7.2f}")
