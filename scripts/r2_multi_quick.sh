#!/bin/bash
N=${1:-2}; LEGS=${2:-c5}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --legs $LEGS > gpurun_out/multi${N}_bench.json 2> gpurun_out/multi${N}_bench.err
echo "bench rc=$?"; grep -E "\[bench\]|Error|error|Traceback" gpurun_out/multi${N}_bench.err | tail -8
python scripts/bench_summary.py gpurun_out/multi${N}_bench.json
