#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "c2_ or solve_device or c1_ or horizon or not_positive or edge" > gpurun_out/c2only_pytest.log 2>&1; tail -3 gpurun_out/c2only_pytest.log
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/c2only.json 2> gpurun_out/c2only.err
python - <<PY
import json
d=json.load(open("gpurun_out/c2only.json")); x=d["detail"]
print("c2 graph:", round(x["latency_us"],1), "us; protocol calls:", round(x["ms_per_step_protocol_calls"]*1e3,1), "us; launches", x["gpu_launches"], {n: round(v["us"],1) for n,v in x["latency_vs_N_us"].items()}, "parity", x.get("parity_rel_err"), "e2e", round(x["e2e"]["ms_per_step"]*1e3,1))
PY
