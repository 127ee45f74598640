#!/bin/bash
# final 1-GPU evidence: parity suite, bench line, ncu launch list of the same bench command, ncu --set full of the C5 stage kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log; tail -4 gpurun_out/final_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py gpurun_out/final_bench.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "reference rc=$?"; head -c 600 gpurun_out/final_bench_reference.json; echo
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/prof_c5.py > gpurun_out/final_prof_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:seg_backward_warp -s 1 -c 1 -o gpurun_out/final_warp_c5 python scripts/prof_c5.py > gpurun_out/final_prof_c5_ncu.log 2>&1
echo "ncu full rc=$?"
