# Round-1 profiling recipe (B200_PROFILING.md): launch list of the bench command, then one full capture of the
# two batch kernels.  Each ncu run only after the identical plain command exited 0.
B3="python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline"
$B3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv $B3 > gpurun_out/ncu_c3.log 2>&1
$B3 > gpurun_out/plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_ -s 4 -c 2 -o gpurun_out/prof_c3 $B3 > gpurun_out/ncu_c3_full.log 2>&1
