"""Group an ncu --csv launch list (--metrics gpu__time_duration.sum) by (kernel, grid): count, median, total, share.
usage: python scripts/ncu_launch_groups.py launches.csv"""
import csv
import statistics
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
kn, mn, mv, gs = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Grid Size"))
groups = {}
n = 0
for r in rows[hdr + 2:]:
    if len(r) > mv and r[mn] == "gpu__time_duration.sum":
        groups.setdefault((r[kn], r[gs]), []).append(float(r[mv].replace(",", "")) / 1e3)   # us
        n += 1
total = sum(sum(v) for v in groups.values())
print("launches captured:", n)
print()
print(f"{'kernel':88s} {'grid':>14s} {'n':>5s} {'median us':>11s} {'total ms':>10s} {'share':>7s}")
for (k, g), v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:88]:88s} {g:>14s} {len(v):5d} {statistics.median(v):11.2f} {sum(v) / 1e3:10.2f} {100 * sum(v) / total:6.1f}%")
