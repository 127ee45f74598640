python scripts/sweep_c2.py 2>&1 | tail -8
B2="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline"
$B2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv $B2 > gpurun_out/ncu_c2.log 2>&1
