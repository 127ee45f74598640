#!/bin/bash
mkdir -p gpurun_out
for L in 450 300 200 140 110; do
  BENCH_C5_SEG_LEN=$L timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/c5_len$L.json 2> gpurun_out/c5_len$L.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c5_len$L.json")); x=d["detail"]
    print("seg len $L: segments", x["num_segments"], "step", round(x["ms_per_step"],4), "kernel", round(x["roofline"]["kernel_ms"],4), "parity", x.get("parity_rel_err"))
except Exception as e: print("len $L failed", e)
PY
done
