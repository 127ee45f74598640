#!/bin/bash
# full parity suite + bench line + ncu --set full of the nx30 factorising sweep and the ADMM update kernel (after: symmetric S3,
# fold one stage ahead, single-pass box rows)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r11_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r11_pytest_gpu.log; tail -4 gpurun_out/r11_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py gpurun_out/r11_bench.json
C4_ITERS=6 python scripts/prof_c4.py > gpurun_out/r11_prof_c4_plain.log 2>&1 && \
PDPLQR_ADMM_GRAPH=0 C4_ITERS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_backward_kernel -c 1 -o gpurun_out/r11_seg30 python scripts/prof_c4.py > gpurun_out/r11_ncu_seg30.log 2>&1
echo "ncu seg30 rc=$?"
PDPLQR_ADMM_GRAPH=0 C4_ITERS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_update_kernel -s 2 -c 1 -o gpurun_out/r11_admm_upd python scripts/prof_c4.py > gpurun_out/r11_ncu_admm.log 2>&1
echo "ncu admm rc=$?"
