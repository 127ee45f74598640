#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "admm or mpc" > gpurun_out/admm_pytest.log 2>&1; tail -3 gpurun_out/admm_pytest.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/c4_split.json 2> gpurun_out/c4_split.err
python - <<PY
import json
d=json.load(open("gpurun_out/c4_split.json")); x=d["detail"]
print("c4: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "launches", x["gpu_launches"], x["parity_rel_err"])
PY
