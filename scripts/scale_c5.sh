N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5_g$N.json 2> gpurun_out/bench_c5_g$N.err
tail -2 gpurun_out/bench_c5_g$N.err; cat gpurun_out/bench_c5_g$N.json
