#!/bin/bash
# N=8: the liftover bench (weak scaling, shard + all-gather of records) and the sharded depth sweep (strong scaling)
mkdir -p gpurun_out
( time timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 700 gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
