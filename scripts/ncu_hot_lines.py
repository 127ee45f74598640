#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples / executed instructions, from an .ncu-rep taken with
--set full --import-source on (kernels compiled with -lineinfo).
usage: python scripts/ncu_hot_lines.py prof.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys


def main(path, regex, top=40):
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
        data.append((samp, inst, cur_file, r[0], r[1].strip()[:100]))
    tot, toti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    print(first_fn)
    print(f"total stall samples {tot}, warp instructions executed {toti}")
    for d in sorted(data, key=lambda x: -x[0])[:top]:
        print(f"{100*d[0]/tot:5.1f}% samples  {100*d[1]/toti:5.1f}% inst  {d[2]}:{d[3]:>4s}  {d[4]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
