#!/bin/bash
TAG=${PDPLQR_RUN_TAG:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "long_horizon or c1_ or c2_ or c3_batch_segmented or random_dense or equal_split or edge_horizons or horizon_shards or arbitrary or costates" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
PDPLQR_VARIANT=prof PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS PDPLQR_WARP_KERNEL=2 timeout 300 python scripts/prof_phases.py > gpurun_out/${TAG}_phases.txt 2>&1
cat gpurun_out/${TAG}_phases.txt
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_c5.json")); x=d["detail"]
print("c5:", x["ms_per_step"], "kernel", x["roofline"]["kernel_ms"], "parity", x.get("parity_rel_err"))
PY
