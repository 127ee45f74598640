#!/bin/bash
# final 1-GPU evidence of round 2 (after the nx30 / ADMM update work): parity suite, bench line, reference arm, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final3_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final3_pytest_gpu.log; tail -3 gpurun_out/final3_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py gpurun_out/final3_bench.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final3_bench_reference.json 2> gpurun_out/final3_bench_reference.err
echo "reference rc=$?"; head -c 400 gpurun_out/final3_bench_reference.json; echo
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final3_plain.log 2>&1 && \
PDPLQR_ADMM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final3_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final3_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 60 compute-sanitizer --tool memcheck python scripts/prof_c2.py 256 > gpurun_out/final3_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/final3_sanitizer.log
python __graft_entry__.py > gpurun_out/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final3_smoke.log
