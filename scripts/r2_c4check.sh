#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4check_pytest.log 2>&1; tail -4 gpurun_out/c4check_pytest.log
PDPLQR_VARIANT=prof PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS python scripts/prof_phases_c4.py 2>&1 | grep seg_backward | tail -2
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/c4check.json 2> gpurun_out/c4check.err
python - <<PY
import json
d=json.load(open("gpurun_out/c4check.json")); x=d["detail"]
print("c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "launches", x["gpu_launches"], x["parity_rel_err"])
PY
