#!/bin/bash
# N=2 check of the multi-GPU paths: the liftover bench (shard + all-gather of records) and the sharded depth sweep.
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
