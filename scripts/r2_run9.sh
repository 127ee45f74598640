#!/bin/bash
# after the affine-kernel fix: constrained subset per variant + C4 timing per variant + phase clocks of the default build
mkdir -p gpurun_out
for v in "" vb vc va1 ve; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -k "constraint or fold or admm or factorization or padded" > gpurun_out/r9_pytest_${v:-default}.log 2>&1
  echo "== variant ${v:-default}: $(tail -1 gpurun_out/r9_pytest_${v:-default}.log)"; grep "^FAILED" gpurun_out/r9_pytest_${v:-default}.log | cut -c1-150 | head -8
  timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r9_c4_${v:-default}.json 2> gpurun_out/r9_c4_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r9_c4_${v:-default}.json")); x=d["detail"]
    print("   c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "parity", x["parity_rel_err"])
except Exception as e: print("   c4 failed", e)
PY
done
PDPLQR_VARIANT=prof timeout 300 python scripts/prof_phases_c4.py > gpurun_out/r9_phases_c4.txt 2>&1; grep -m2 "seg_backward" gpurun_out/r9_phases_c4.txt | cut -c1-400
