#!/bin/bash
# Second gpurun call of the session: the GPU tests of what changed since the first call (coalescence limit, --unique,
# wiggle, text layer), the N=1 bench with the in-process clock sampler, and the CLI timing breakdown.
mkdir -p gpurun_out
nproc > gpurun_out/nproc2.txt
( time timeout 420 python -m pytest tests -m gpu -x -q -k "golden_text or cli_cuda or wiggle or fast_path_cuda or equals_oracle" ) > gpurun_out/pytest_gpu2.txt 2>&1
tail -5 gpurun_out/pytest_gpu2.txt
( time timeout 400 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err
tail -c 400 gpurun_out/bench_n1_b.json
( time timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-depth --no-maf --no-cli --no-wiggle ) > gpurun_out/bench_n1_steps20.json 2> gpurun_out/bench_n1_steps20.err
ls -la gpurun_out
