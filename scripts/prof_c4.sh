export C4_BATCH=1024 C4_ITERS=5 C4_STEPS=1
C="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline"
$C > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv $C > gpurun_out/ncu_c4.log 2>&1
