#!/bin/bash
# which programmatic edge breaks the constrained latency-path tests?  (mask bits: 0 stage sweep, 2 rollout, 3 tree up, 4 tree down)
mkdir -p gpurun_out
K="backward_without_factorization or arbitrary_dimensions_with_constraints or horizon_shards_with_constraints or constraint"
for m in 0 4 8 16 12 28 31; do
  for rep in 1 2; do
    PDPLQR_PDL=1 PDPLQR_PDL_MASK=$m timeout 300 python -m pytest tests -m gpu -q -k "$K" -p no:cacheprovider > gpurun_out/r16_mask${m}_$rep.log 2>&1
    echo "mask $m rep $rep: $(tail -1 gpurun_out/r16_mask${m}_$rep.log) | $(grep -c FAILED gpurun_out/r16_mask${m}_$rep.log) failed: $(grep FAILED gpurun_out/r16_mask${m}_$rep.log | sed 's/.*:://' | tr '\n' ' ')"
  done
done
