#!/bin/bash
# multi-GPU checks: the real NCCL horizon-sharded test and the bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "nccl" > gpurun_out/multi${N}_pytest.log 2>&1
tail -5 gpurun_out/multi${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi${N}_bench.json 2> gpurun_out/multi${N}_bench.err
echo "bench rc=$?"; grep -E "\[bench\]|Error|error|Traceback" gpurun_out/multi${N}_bench.err | tail -8
python scripts/bench_summary.py gpurun_out/multi${N}_bench.json
