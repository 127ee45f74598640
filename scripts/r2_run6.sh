#!/bin/bash
# ADMM update prefetch variants at C4 + the patched reference example test + ncu --set full of the nx30 factorising sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "reference_example or admm" > gpurun_out/r6_pytest.log 2>&1; tail -3 gpurun_out/r6_pytest.log
for v in "" pf0 pf1mb6 pf2 pf2mb6; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r6_c4_${v:-default}.json 2> gpurun_out/r6_c4_${v:-default}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r6_c4_${v:-default}.json")); x=d["detail"]
    print("variant ${v:-default}: step", round(x["ms_per_step"],2), "ms; affine", round(x["ms_affine_backward"],3), "fact", round(x["ms_factorizing_backward"],2), "parity", x["parity_rel_err"])
except Exception as e: print("variant ${v:-default} failed", e)
PY
done
unset PDPLQR_VARIANT
C4_ITERS=6 python scripts/prof_c4.py > gpurun_out/r6_prof_c4_plain.log 2>&1 && \
PDPLQR_ADMM_GRAPH=0 C4_ITERS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_backward_kernel -c 1 -o gpurun_out/r6_seg30 python scripts/prof_c4.py > gpurun_out/r6_ncu_seg30.log 2>&1
echo "ncu seg30 rc=$?"
PDPLQR_ADMM_GRAPH=0 C4_ITERS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_update_kernel -s 2 -c 1 -o gpurun_out/r6_admm_upd python scripts/prof_c4.py > gpurun_out/r6_ncu_admm.log 2>&1
echo "ncu admm rc=$?"
ls -la gpurun_out/*.ncu-rep
