#!/bin/bash
# parity suite + bench (all configs) + A/B of the warp kernel
TAG=${PDPLQR_RUN_TAG:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python scripts/bench_summary.py gpurun_out/${TAG}_bench.json
for wk in 0 2; do
  PDPLQR_WARP_KERNEL=$wk timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${TAG}_c2_wk$wk.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_c2_wk$wk.json"))
x=d["detail"]
print("c2 WARP_KERNEL=$wk:", round(x["latency_us"],1), "us", {n: round(v["us"],1) for n,v in x["latency_vs_N_us"].items()}, "parity", x.get("parity_rel_err"))
PY
done
