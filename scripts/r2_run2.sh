#!/bin/bash
# parity suite + phase clocks + bench (all configs); PDPLQR_RUN_TAG names the outputs
TAG=${PDPLQR_RUN_TAG:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
if [ -f pdp-lqr_b200/libpdplqr_prof.so ]; then
  PDPLQR_VARIANT=prof PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS timeout 300 python scripts/prof_phases.py > gpurun_out/${TAG}_phases.txt 2>&1
  cat gpurun_out/${TAG}_phases.txt
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python scripts/bench_summary.py gpurun_out/${TAG}_bench.json
