#!/bin/bash
# round-2 GPU check: parity suite, then the default bench line (all configs)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
nproc >> gpurun_out/r2_gpus.txt; free -g >> gpurun_out/r2_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json | head -c 6000
