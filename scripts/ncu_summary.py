"""Summarise an ncu report (run here, no GPU needed): python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h:75s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {s: 0 for s in stalls}; ts = 0; ti = 0
for r in data:
    ts += int(r[ix['# Samples']]); ti += int(r[ix['Instructions Executed']])
    for s in stalls: tot[s] += int(r[ix[s]])
print(f"samples {ts}  warp-instructions {ti}")
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    print(f"  {s:28s} {100 * v / ts:5.1f}%")
print("top instructions by stall samples:")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    d = {s: int(r[ix[s]]) for s in stalls}; mx = max(d, key=d.get)
    print(f"  {100 * int(r[ix['# Samples']]) / ts:5.1f}%  x{r[ix['Instructions Executed']]:>11s}  {r[ix['Source']].strip()[:64]:64s} {mx}")
