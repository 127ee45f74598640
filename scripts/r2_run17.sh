#!/bin/bash
# PDL experiment: does an explicit release (__threadfence) of the stage sweep's summary stores cure the failing edge?
# mask bits: 3 = tree-up launches carry the attribute, 5 = fence at the end of seg_backward_kernel
mkdir -p gpurun_out
K="backward_without_factorization or arbitrary_dimensions_with_constraints or horizon_shards_with_constraints or constraint"
for m in 8 40 8 40 8 40 63 63; do
  PDPLQR_PDL=1 PDPLQR_PDL_MASK=$m timeout 300 python -m pytest tests -m gpu -q -k "$K" -p no:cacheprovider > gpurun_out/r17_mask${m}.log 2>&1
  echo "mask $m: $(tail -1 gpurun_out/r17_mask${m}.log) | failed: $(grep FAILED gpurun_out/r17_mask${m}.log | sed 's/.*:://' | tr '\n' ' ')"
done
