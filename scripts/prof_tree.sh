export SWEEP_S=32
C="python scripts/sweep_c2.py"
$C > gpurun_out/plain_tree2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tree_top_up" -s 4 -c 1 -o gpurun_out/prof_tree $C > gpurun_out/ncu_tree_full.log 2>&1
