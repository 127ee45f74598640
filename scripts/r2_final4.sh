#!/bin/bash
# final 1-GPU evidence of round 2 (late session): parity suite (plain, and in guard-band mode = the memcheck substitute),
# bench line, reference arm, ncu launch list of the bench command, smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final4_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final4_pytest_gpu.log; tail -3 gpurun_out/final4_pytest_gpu.log
PDPLQR_DEBUG_GUARDS=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final4_pytest_gpu_guards.log 2>&1
echo "pytest (PDPLQR_DEBUG_GUARDS=1) rc=$?" >> gpurun_out/final4_pytest_gpu_guards.log; tail -3 gpurun_out/final4_pytest_gpu_guards.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py gpurun_out/final4_bench.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final4_bench_reference.json 2> gpurun_out/final4_bench_reference.err
echo "reference rc=$?"; head -c 300 gpurun_out/final4_bench_reference.json; echo
C4_TOL_MAX_ITER=0 PDPLQR_ADMM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final4_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final4_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python __graft_entry__.py > gpurun_out/final4_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final4_smoke.log
