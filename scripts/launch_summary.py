"""Per-kernel totals of an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list:
python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"<.*", "", name)
    a = agg.setdefault(name, {"n": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"].add(r[0])
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if r[-3] == "gpu__time_duration.sum":
        a["ns"] += v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    elif r[-3].startswith("dram__bytes"):
        b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        a["rd" if "read" in r[-3] else "wr"] += b
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'dram rd GB':>11s} {'dram wr GB':>11s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{k[:60]:60s} {len(a['n']):8d} {a['ns'] / 1e6:10.3f} {100 * a['ns'] / tot:6.1f}% {a['rd'] / 1e9:11.3f} {a['wr'] / 1e9:11.3f}")
print(f"{'total':60s} {len(rows):8d} {tot / 1e6:10.3f}")
