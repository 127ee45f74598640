#!/bin/bash
# Round-end 1-GPU measurements: one JSON line per workload into gpurun_out/, then ncu evidence.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c3 c1 c2 c4 c5; do
  python bench.py --workload $w > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
  tail -c 400 gpurun_out/final_bench_$w.json
done
python scripts/sweep_c2.py > gpurun_out/final_sweep_c2.txt 2>&1
# ncu: launch list of the default bench command, then full captures of the dominant segment-path kernels
python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c3.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final_ncu_c3.log 2>&1
python bench.py --workload c5small --steps 3 --warmup 3 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:seg_backward_kernel --launch-skip 3 -c 1 -o gpurun_out/final_segbwd -f python bench.py --workload c5small --steps 3 --warmup 3 > gpurun_out/final_ncu_segbwd.log 2>&1
SWEEP_S=128 python scripts/sweep_c2.py > /dev/null 2>&1 && SWEEP_S=128 ncu --set full --import-source on --clock-control none --cache-control none -k regex:tree_sub_up --launch-skip 20 -c 2 -o gpurun_out/final_treeup -f python scripts/sweep_c2.py > gpurun_out/final_ncu_treeup.log 2>&1
SWEEP_S=128 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/final_c2_warm.csv python scripts/sweep_c2.py > /dev/null 2>&1
python bench.py --workload c5 --steps 3 --warmup 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/final_c5_warm.csv python bench.py --workload c5 --steps 3 --warmup 3 > /dev/null 2>&1
echo done
