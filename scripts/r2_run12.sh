#!/bin/bash
# ADMM update kernel with hoisted row loads: resident-CTA variants at C4 + ADMM parity subset
mkdir -p gpurun_out
for v in "" am6 am5 am4; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r12_c4_${v:-default}.json 2> gpurun_out/r12_c4_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r12_c4_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "parity", x["parity_rel_err"])
except Exception as e: print("   c4 failed", e)
PY
done
unset PDPLQR_VARIANT
timeout 600 python -m pytest tests -m gpu -q -k "admm or mpc" > gpurun_out/r12_pytest.log 2>&1; tail -1 gpurun_out/r12_pytest.log
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r12_c2.json 2> gpurun_out/r12_c2.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r12_c2.json")); x=d["detail"]
    print("c2: step us", round(x["ms_per_step"]*1e3,2), "parity", x["parity_rel_err"], "lat", x.get("latency_vs_N"))
except Exception as e: print("   c2 failed", e)
PY
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r12_pytest_full.log 2>&1; tail -2 gpurun_out/r12_pytest_full.log
