#!/bin/bash
# blocked Gauss-Jordan in the latency tree: tree parity tests, C2 latency (old / new), phase clocks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x -k "tree or c1 or c2 or latency or horizon or shard or solve_device or multilevel or interface" > gpurun_out/r14_pytest.log 2>&1; tail -3 gpurun_out/r14_pytest.log
for v in "" gj1; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r14_c2_${v:-default}.json 2> gpurun_out/r14_c2_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r14_c2_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default} c2: step us", round(x["ms_per_step"]*1e3,2), "parity", x["parity_rel_err"])
except Exception as e: print("   c2 failed", e)
PY
done
PDPLQR_VARIANT=prof timeout 200 python scripts/prof_c2.py 1024 > gpurun_out/r14_tree_phases.txt 2>&1; grep combine_lat gpurun_out/r14_tree_phases.txt | sort | uniq -c | sort -rn | head -6
