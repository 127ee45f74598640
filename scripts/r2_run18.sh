#!/bin/bash
# ncu --set full of the latency-mode interface tree at C2 (the 61 of 94 us of the latency path)
mkdir -p gpurun_out
python scripts/prof_c2.py 1024 > gpurun_out/r18_plain.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tree_sub_up_lat_kernel -s 6 -c 6 -f -o gpurun_out/r2_c2_tree_up python scripts/prof_c2.py 1024 > gpurun_out/r18_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r18_ncu.log; ls -la gpurun_out/r2_c2_tree_up.ncu-rep
