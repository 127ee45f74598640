C="python bench.py --workload c5small --steps 3 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/plain_seg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"seg_backward" -s 2 -c 1 -o gpurun_out/prof_seg_km $C > gpurun_out/ncu_seg_km.log 2>&1
export PDPLQR_USE_KM=0
$C > gpurun_out/plain_seg2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"seg_backward" -s 2 -c 1 -o gpurun_out/prof_seg_gen $C > gpurun_out/ncu_seg_gen.log 2>&1
