#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sharded or nccl" > gpurun_out/multi_tests.log 2>&1
tail -8 gpurun_out/multi_tests.log
