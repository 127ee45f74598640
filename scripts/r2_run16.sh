#!/bin/bash
# rollout kernel ring depth: C5 / C2 timing per variant + parity suite of the default
mkdir -p gpurun_out
for v in "" fd2 fd3 fd6; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  for w in c5 c2; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r16_${w}_${v:-default}.json 2> gpurun_out/r16_${w}_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r16_${w}_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} $w: step", round(x["ms_per_step"],4), "parity", x.get("parity_rel_err"))
except Exception as e: print("${v:-default} $w failed", e)
PY
  done
done
unset PDPLQR_VARIANT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r16_pytest.log 2>&1; tail -2 gpurun_out/r16_pytest.log
