#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("==", d["Kernel Name"][:70], d.get("Grid Size"), d.get("Block Size"))
        for w in WANT:
            if w in d:
                print(f"   {w:70s} {d[w]:>18s} {u[w]}")
        stalls = [(k, float(v.replace(",", ""))) for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v]
        for k, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
            print(f"   stall {k[len(STALL):-len('_per_issue_active.ratio')]:40s} {v:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
