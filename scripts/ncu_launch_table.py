"""Print the last N launches of an ncu --csv launch list: kernel, grid, block, duration, warp instructions."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
kn, mn, mv, gs, bs, idc = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size", "ID"))
d = {}
for r in rows[hdr + 2:]:
    if len(r) > mv:
        d.setdefault(r[idc], {"k": r[kn].split("(")[0][-48:], "g": r[gs], "b": r[bs]})[r[mn]] = r[mv]
for i in sorted(d, key=int)[-n:]:
    e = d[i]
    print(f"{e['k']:50s} {e['g']:>14s} {e['b']:>14s} {e.get('gpu__time_duration.sum', ''):>10s} ns {e.get('smsp__inst_executed.sum', ''):>10s} inst")
