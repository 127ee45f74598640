#!/bin/bash
# PDL restricted to chain -> chain edges: suite (default; PDL forced on every handle + guard bands), C2 A/B, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r15_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r15_pytest.log
PDPLQR_PDL=1 PDPLQR_DEBUG_GUARDS=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r15_pytest_pdl_guards.log 2>&1; echo "pytest pdl+guards rc=$?"; tail -12 gpurun_out/r15_pytest_pdl_guards.log | cut -c1-300
for v in 0 1; do
  PDPLQR_PDL=$v timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r15_c2_pdl$v.json 2> gpurun_out/r15_c2_pdl$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r15_c2_pdl$v.json")); x=d["detail"]
    print("PDL=$v c2: graph us", round(x["ms_per_step"]*1e3,2), "protocol us", round(x["ms_per_step_protocol_calls"]*1e3,2), "parity", x["parity_rel_err"], "graph==protocol", x["graph_matches_protocol_calls"], "lat", {n: round(e["us"],1) for n,e in x["latency_vs_N_us"].items()})
except Exception as e: print("   c2 failed", e)
PY
done
timeout 600 python bench.py > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/r15_bench.json 2>&1 | cut -c1-600
