#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples / executed instructions / shared-memory wavefronts, from an
.ncu-rep taken with --set full --import-source on (kernels compiled with -lineinfo).
usage: python scripts/ncu_hot_lines.py prof.ncu-rep <kernel regex> [top] [samples|inst|smem]"""
import csv
import subprocess
import sys


def main(path, regex, top=40, key="samples"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{regex}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, data, first_fn = None, None, [], None
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            if first_fn is None:
                first_fn = r[1]
            elif r[1] != first_fn and hdr is not None:
                break          # only the first matching kernel instance
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            samp = int(r[hdr.index("Warp Stall Sampling (All Samples)")])
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        try:
            wav = int(r[hdr.index("L1 Wavefronts Shared")])
            ideal = int(r[hdr.index("L1 Wavefronts Shared Ideal")])
        except ValueError:
            wav = ideal = 0
        data.append((samp, inst, cur_file, r[0], r[1].strip()[:100], wav, ideal))
    # the cuda,sass view repeats a source line once per SASS instruction: fold them
    agg = {}
    for d in data:
        a = agg.setdefault((d[2], d[3]), [0, 0, d[2], d[3], d[4], 0, 0])
        a[0] += d[0]; a[1] += d[1]; a[5] += d[5]; a[6] += d[6]
    data = list(agg.values())
    tot, toti, totw = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1, sum(d[5] for d in data) or 1
    print(first_fn)
    print(f"total stall samples {tot}, warp instructions executed {toti}, shared wavefronts {totw} (ideal {sum(d[6] for d in data)})")
    col = {"samples": 0, "inst": 1, "smem": 5}[key]
    for d in sorted(data, key=lambda x: -x[col])[:top]:
        print(f"{100*d[0]/tot:5.1f}% samples  {100*d[1]/toti:5.1f}% inst  {100*d[5]/totw:5.1f}% smem-wavefronts (x{d[5]/max(d[6],1):.2f} ideal)  {d[2]}:{d[3]:>4s}  {d[4]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else "samples")
