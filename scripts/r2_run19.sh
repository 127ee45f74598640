#!/bin/bash
# shorter dependent chains in the affine sweep and the rollout: parity suite, C4 / C2 / C5 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r19_pytest.log 2>&1; tail -2 gpurun_out/r19_pytest.log
for w in c4 c2 c5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r19_$w.json 2> gpurun_out/r19_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r19_$w.json")); x=d["detail"]
    print("$w: step", round(x["ms_per_step"],4), "affine", x.get("ms_affine_backward"), "e2e", round(x["e2e"]["ms_per_step"],3), "parity", x.get("parity_rel_err"))
except Exception as e: print("$w failed", e)
PY
done
