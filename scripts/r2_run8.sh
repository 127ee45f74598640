#!/bin/bash
# isolate the constrained-path regression: debug-switch variants on the constrained / ADMM subset of the parity suite
mkdir -p gpurun_out
for v in "" vb vc vd ve; do
  if [ -n "$v" ]; then export PDPLQR_VARIANT=$v; else unset PDPLQR_VARIANT; fi
  timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -k "constraint or fold or admm or factorization or padded" > gpurun_out/r8_pytest_${v:-default}.log 2>&1
  echo "== variant ${v:-default}: $(tail -1 gpurun_out/r8_pytest_${v:-default}.log)"; grep "^FAILED" gpurun_out/r8_pytest_${v:-default}.log | cut -c1-150 | head -12
done
