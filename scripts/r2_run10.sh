#!/bin/bash
# compile-time tile lists in the symmetric S3: C4 timing per variant, phase clocks, C2 latency, constrained parity subset
mkdir -p gpurun_out
for v in "" vb vnv; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r10_c4_${v:-default}.json 2> gpurun_out/r10_c4_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r10_c4_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "parity", x["parity_rel_err"])
except Exception as e: print("   c4 failed", e)
PY
  timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r10_c2_${v:-default}.json 2> gpurun_out/r10_c2_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r10_c2_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} c2: step us", round(x["ms_per_step"]*1e3,2), "parity", x["parity_rel_err"])
except Exception as e: print("   c2 failed", e)
PY
done
unset PDPLQR_VARIANT
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -k "constraint or fold or admm or factorization or padded or c1 or config" > gpurun_out/r10_pytest.log 2>&1; tail -1 gpurun_out/r10_pytest.log
PDPLQR_VARIANT=prof timeout 300 python scripts/prof_phases_c4.py > gpurun_out/r10_phases_c4.txt 2>&1; grep -m2 "seg_backward" gpurun_out/r10_phases_c4.txt | cut -c1-400
