#!/bin/bash
# symmetric S3 + fold one stage ahead + H mirrored at pack time: parity suite, C4 timing, phase clocks, C5/C2 sanity
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest.log 2>&1; tail -5 gpurun_out/r7_pytest.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r7_c4.json 2> gpurun_out/r7_c4.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r7_c4.json")); x=d["detail"]
    print("c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "parity", x["parity_rel_err"])
except Exception as e: print("c4 failed", e)
PY
PDPLQR_VARIANT=prof timeout 300 python scripts/prof_phases_c4.py > gpurun_out/r7_phases_c4.txt 2>&1; grep -m2 "seg_backward" gpurun_out/r7_phases_c4.txt | cut -c1-400
