#!/bin/bash
# One gpurun call: GPU tests, the N=1 bench, the launch list of a short bench run, one ncu --set full capture of the
# wiggle-mode mapping kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cli --no-maf > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:liftoverKernel -s 1 -c 1 -f -o gpurun_out/wiggle_full \
    python tools/profile_wiggle.py 50000000 > gpurun_out/wiggle_ncu.log 2>&1
ncu -i gpurun_out/wiggle_full.ncu-rep --page raw --csv > gpurun_out/wiggle_full_raw.csv 2>/dev/null
ls -la gpurun_out
