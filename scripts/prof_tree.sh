export SWEEP_S=32
C="python scripts/sweep_c2.py"
$C > gpurun_out/plain_tree.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 8 --csv --log-file gpurun_out/launches_tree.csv $C > gpurun_out/ncu_tree.log 2>&1
$C > gpurun_out/plain_tree2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tree_top_up|seg_backward" -s 4 -c 2 -o gpurun_out/prof_tree $C > gpurun_out/ncu_tree_full.log 2>&1
