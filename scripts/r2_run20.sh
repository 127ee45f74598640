#!/bin/bash
# ADMM update split into a flat relaxation kernel + per-stage row kernel: variants at C4, ADMM parity subset
mkdir -p gpurun_out
for v in "" h1 h1m6 m6 h1m4; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r20_c4_${v:-default}.json 2> gpurun_out/r20_c4_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r20_c4_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "launches", x["gpu_launches"], "parity", x["parity_rel_err"])
except Exception as e: print("   c4 failed", e)
PY
done
unset PDPLQR_VARIANT
timeout 600 python -m pytest tests -m gpu -q -k "admm or mpc" > gpurun_out/r20_pytest.log 2>&1; tail -1 gpurun_out/r20_pytest.log
