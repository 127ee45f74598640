#!/bin/bash
# round-2 late: GPU suite (plain and in guard-band mode = the memcheck substitute), then the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r13_pytest.log
PDPLQR_DEBUG_GUARDS=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r13_pytest_guards.log 2>&1; echo "pytest guards rc=$?"; tail -25 gpurun_out/r13_pytest_guards.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/r13_bench.json 2>&1 | cut -c1-600
python - <<'PY'
import json
d = json.load(open("gpurun_out/r13_bench.json"))
print("to_tolerance:", d["configs"]["c4"].get("to_tolerance"))
for n, e in d["configs"]["c2"]["latency_vs_N_us"].items():
    print(n, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in e.items() if k != "cpu_kind"})
PY
