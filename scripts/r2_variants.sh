#!/bin/bash
# C5 (N = 2^20) timing of register-capped builds of the warp kernel
mkdir -p gpurun_out
for v in "" mb16 mb20 mb24; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v PDPLQR_CFLAGS=-DPDPLQR_WARP_MINB=${v#mb}; else unset PDPLQR_VARIANT PDPLQR_CFLAGS; fi
  timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/var_${v:-default}.json 2> gpurun_out/var_${v:-default}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_${v:-default}.json")); x=d["detail"]
print("variant ${v:-default}: segments", x["num_segments"], "step", round(x["ms_per_step"],4), "kernel", round(x["roofline"]["kernel_ms"],4), "frac", round(x["roofline"]["frac"],3), "parity", x.get("parity_rel_err"))
PY
done
