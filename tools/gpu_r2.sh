#!/bin/bash
# Round-2 GPU session script (one gpurun call): gpurun -- 'bash tools/gpu_r2.sh <stage>'
# Outputs land in gpurun_out/ (merged back); numbers quoted in profiles/ come from these files.
mkdir -p gpurun_out
stage=${1:-a}
case $stage in
a)  # parity of the liftover path + first bench with the one-lane-per-interval kernel + launch list + ncu of both kernels
    timeout 900 python -m pytest tests/test_liftover_gpu.py -x -q -m gpu > gpurun_out/pytest_lift.log 2>&1; tail -3 gpurun_out/pytest_lift.log
    ( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cli --no-maf --no-wiggle ) > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
    tail -c 1500 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_a.csv \
        python bench.py --steps 2 --warmup 1 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth > gpurun_out/bench_ncu_a.json 2> gpurun_out/bench_ncu_a.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastLiftKernel -s 1 -c 1 -o gpurun_out/prof_fast -f \
        python bench.py --steps 1 --warmup 1 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth --no-divergent > /dev/null 2> gpurun_out/prof_fast.err
    ;;
b)  # whole GPU test tier + bench + ncu of the fast kernel (2-level hops), the piece-walk depth kernel and the divergent walk
    timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_b.log 2>&1; tail -5 gpurun_out/pytest_b.log
    ( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
    tail -c 600 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b.csv \
        python bench.py --steps 2 --warmup 1 --no-cli --no-wiggle --no-cpu-baseline > gpurun_out/bench_ncu_b.json 2> gpurun_out/bench_ncu_b.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fastLiftKernel|depthKernel' -s 2 -c 3 -o gpurun_out/prof_b -f \
        python bench.py --steps 1 --warmup 1 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-divergent > /dev/null 2> gpurun_out/prof_b.err
    ;;
c)  # reworked bench (checks, in-run ncu traffic) at N=1; with 2 GPUs: NCCL tests of the C++ multi path + bench at N=2
    timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_liftover_gpu.py -x -q -m gpu > gpurun_out/pytest_c.log 2>&1; tail -5 gpurun_out/pytest_c.log
    ( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
    tail -c 400 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
    if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
        ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_c_n2.json 2> gpurun_out/bench_c_n2.err
        tail -c 1500 gpurun_out/bench_c_n2.json; tail -5 gpurun_out/bench_c_n2.err
    fi
    ;;
d)  # TMA seed-tile A/B (tools/tile_ab.py) with and without ncu
    timeout 900 python tools/tile_ab.py > gpurun_out/tile_ab.json 2> gpurun_out/tile_ab.err; cat gpurun_out/tile_ab.json | head -60; tail -3 gpurun_out/tile_ab.err
    TILE_AB_REPS=1 TILE_AB_INTERVALS=2000000 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:liftoverKernel --csv --log-file gpurun_out/tile_ab_ncu.csv python tools/tile_ab.py > gpurun_out/tile_ab_under_ncu.json 2>> gpurun_out/tile_ab.err
    ;;
e)  # CLI wall-clock breakdowns (halLiftover pipeline, hal2maf with device text vs host text) on the bench files
    python - <<'PY' > gpurun_out/cli_e.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import bench
W = bench.WORKLOADS["C2"]
hal = bench.ensure_hal("C2", W["segs"])
gs, ge = bench.make_intervals(10_000_000, W["segs"] * 32, 2)
bench.write_bed3("/tmp/in.bed", "L0_seq", gs, ge)
env = dict(os.environ, HALGPU_TIMING="1")
for i in range(3):
    t = time.time(); r = subprocess.run(["hal_b200/bin/halLiftover", hal, "L0", "/tmp/in.bed", "L7", "/tmp/out.bed"], env=env, capture_output=True, text=True)
    print("halLiftover wall %.3f s" % (time.time() - t)); print(r.stderr[-1500:])
for name, extra in (("device text", {}), ("host text", {"HALGPU_MAF_HOST_TEXT": "1"})):
    for i in range(2):
        t = time.time(); r = subprocess.run(["hal_b200/bin/hal2maf", hal, "/tmp/out.maf", "--refGenome", "R", "--refSequence", "R_seq"], env=dict(env, **extra), capture_output=True, text=True)
        print("hal2maf (%s) wall %.3f s, %d bytes" % (name, time.time() - t, os.path.getsize("/tmp/out.maf"))); print(r.stderr[-800:])
    os.rename("/tmp/out.maf", "/tmp/out_%s.maf" % name.split()[0])
print("maf outputs identical:", open("/tmp/out_device.maf", "rb").read() == open("/tmp/out_host.maf", "rb").read())
PY
    cat gpurun_out/cli_e.log
    ;;
f)  # whole GPU tier + bench at N=1 (FastRec hops, packed sort) + launch list + ncu of the fast kernel on the full batch
    timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_f.log 2>&1; tail -5 gpurun_out/pytest_f.log
    ( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
    tail -c 300 gpurun_out/bench_f.json; tail -5 gpurun_out/bench_f.err
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f.csv \
        python bench.py --steps 2 --warmup 1 --no-cli --no-wiggle --no-cpu-baseline --no-traffic > gpurun_out/bench_ncu_f.json 2> gpurun_out/bench_ncu_f.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fastLiftKernel' -s 1 -c 1 -o gpurun_out/prof_f_fast -f \
        python bench.py --probe --no-divergent > /dev/null 2> gpurun_out/prof_f.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'depthKernel' -c 1 -o gpurun_out/prof_f_depth -f \
        python bench.py --probe --no-divergent > /dev/null 2>> gpurun_out/prof_f.err
    ;;
g)  # multi-GPU: NCCL tests of the C++ path + bench at N = number of GPUs on the box
    N=$(nvidia-smi -L | wc -l)
    timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_g.log 2>&1; tail -5 gpurun_out/pytest_g.log
    ( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_g_n$N.json 2> gpurun_out/bench_g_n$N.err
    tail -c 2500 gpurun_out/bench_g_n$N.json; tail -8 gpurun_out/bench_g_n$N.err
    ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gather --no-depth --no-c4 ) > gpurun_out/bench_g_n${N}_nogather.json 2> gpurun_out/bench_g_n${N}_nogather.err
    tail -c 600 gpurun_out/bench_g_n${N}_nogather.json
    ;;
h)  # source-level profile of the warp-per-interval walk on the divergent file (2 M intervals)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'liftoverKernel' -s 3 -c 1 -o gpurun_out/prof_h_walk -f \
        python tools/walk_profile.py > gpurun_out/prof_h.log 2> gpurun_out/prof_h.err
    ;;
i)  # 1-GPU experiments: sort granularity, CLI / hal2maf breakdowns after the text pipeline changes
    Q="--steps 20 --warmup 5 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth --no-divergent --no-traffic"
    python bench.py $Q > gpurun_out/bench_i_default.json 2> gpurun_out/bench_i.err
    HALGPU_SORT_BITS=8 python bench.py $Q > gpurun_out/bench_i_sort8.json 2>> gpurun_out/bench_i.err
    HALGPU_SORT_BITS=24 python bench.py $Q > gpurun_out/bench_i_sort24.json 2>> gpurun_out/bench_i.err
    HALGPU_PACKED_SORT=1 python bench.py $Q > gpurun_out/bench_i_packed.json 2>> gpurun_out/bench_i.err
    for f in default sort8 sort24 packed; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_i_$f.json').read().strip().splitlines()[-1])
print('$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'fast_ms %.4f' % d['detail']['fast_kernel_ms'], 'e2e %.4g' % d['e2e']['value'], d['check'])
"; done
    bash tools/gpu_r2.sh e
    bash tools/gpu_r2.sh h
    ;;
j)  # quick N=1 bench variants (sliced sort on/off) + liftover GPU tests + divergent walk timing
    timeout 900 python -m pytest tests/test_liftover_gpu.py -x -q -m gpu > gpurun_out/pytest_j.log 2>&1; tail -3 gpurun_out/pytest_j.log
    Q="--steps 20 --warmup 5 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth --no-traffic"
    python bench.py $Q > gpurun_out/bench_j_default.json 2> gpurun_out/bench_j.err
    HALGPU_SLICES=1 python bench.py $Q --no-divergent > gpurun_out/bench_j_slices1.json 2>> gpurun_out/bench_j.err
    HALGPU_SLICES=2 python bench.py $Q --no-divergent > gpurun_out/bench_j_slices2.json 2>> gpurun_out/bench_j.err
    for f in default slices1 slices2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_j_$f.json').read().strip().splitlines()[-1])
print('$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'fast_ms %.4f' % d['detail']['fast_kernel_ms'], 'e2e %.4g' % d['e2e']['value'], d['check'], (d.get('secondary_divergent') or {}).get('value'), (d.get('secondary_divergent') or {}).get('kernel_ms'), (d.get('secondary_divergent') or {}).get('check'))
"; done
    ;;
k)  # bucket sort / tile scheduling A/B at N=1 + liftover GPU tests + launch list of the default step
    timeout 900 python -m pytest tests/test_liftover_gpu.py -x -q -m gpu > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
    Q="--steps 20 --warmup 5 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth --no-traffic"
    python bench.py $Q > gpurun_out/bench_k_default.json 2> gpurun_out/bench_k.err
    HALGPU_RADIX_SORT=1 python bench.py $Q --no-divergent > gpurun_out/bench_k_radix.json 2>> gpurun_out/bench_k.err
    HALGPU_TILE_GRAB=1 python bench.py $Q --no-divergent > gpurun_out/bench_k_grab1.json 2>> gpurun_out/bench_k.err
    HALGPU_TILE_GRAB=4 python bench.py $Q --no-divergent > gpurun_out/bench_k_grab4.json 2>> gpurun_out/bench_k.err
    HALGPU_SORT_BITS=14 python bench.py $Q --no-divergent > gpurun_out/bench_k_bits14.json 2>> gpurun_out/bench_k.err
    for f in default radix grab1 grab4 bits14; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_k_$f.json').read().strip().splitlines()[-1])
print('$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'fast_ms %.4f' % d['detail']['fast_kernel_ms'], 'e2e %.4g' % d['e2e']['value'], d['check'], (d.get('secondary_divergent') or {}).get('value'), (d.get('secondary_divergent') or {}).get('kernel_ms'), (d.get('secondary_divergent') or {}).get('check'))
"; done
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k.csv \
        python bench.py --probe --no-divergent > /dev/null 2> gpurun_out/bench_ncu_k.err
    ;;
l)  # lane-kernel tile hand-out A/B at N=1 (HALGPU_TILE_GRAB) with the radix sort
    Q="--steps 20 --warmup 5 --no-cli --no-maf --no-wiggle --no-cpu-baseline --no-depth --no-traffic --no-divergent"
    for g in 4 2 8 16 32; do
        HALGPU_TILE_GRAB=$g python bench.py $Q > gpurun_out/bench_l_grab$g.json 2>> gpurun_out/bench_l.err
        python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_l_grab$g.json').read().strip().splitlines()[-1])
print('grab$g', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'fast_ms %.4f' % d['detail']['fast_kernel_ms'], 'e2e %.4g' % d['e2e']['value'], d['check'])
"
    done
    ;;
m)  # multi-GPU (gpurun --gpus N): NCCL tests of the C++ path, the gather timeline, bench at N with the default / forced wire forms
    N=$(nvidia-smi -L | wc -l)
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    Q="--gpus $N --no-cpu-baseline --no-depth --no-c4"
    timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_m.log 2>&1; tail -3 gpurun_out/pytest_m.log
    HALGPU_GATHER_TIMELINE=1 timeout 600 $T --master-port 29611 bench.py $Q --steps 6 --warmup 3 > gpurun_out/bench_m_n${N}_timeline.json 2> gpurun_out/bench_m_n${N}_timeline.err
    grep "timeline rank 0" gpurun_out/bench_m_n${N}_timeline.err | tail -8
    timeout 600 $T --master-port 29612 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_m_n${N}.json 2> gpurun_out/bench_m_n${N}.err
    HALGPU_GATHER_WIRE16=1 timeout 600 $T --master-port 29613 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_m_n${N}_wire16.json 2> gpurun_out/bench_m_n${N}_wire16.err
    HALGPU_GATHER_WIRE32=1 timeout 600 $T --master-port 29614 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_m_n${N}_wire32.json 2> gpurun_out/bench_m_n${N}_wire32.err
    HALGPU_GATHER_NCCL=1 timeout 600 $T --master-port 29615 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_m_n${N}_nccl.json 2> gpurun_out/bench_m_n${N}_nccl.err
    timeout 600 $T --master-port 29616 bench.py $Q --steps 20 --warmup 5 --no-gather > gpurun_out/bench_m_n${N}_nogather.json 2> gpurun_out/bench_m_n${N}_nogather.err
    for f in "" _wire16 _wire32 _nccl _nogather; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_m_n${N}$f.json').read().strip().splitlines()[-1])
print('n$N$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'kernels %.4f' % d['detail']['mapping_kernels_ms'], d['check'])
"; done
    ;;
n)  # multi-GPU at N >= 4 (gpurun --gpus N): default gather (peer copies, compact wire) against NCCL and the sharded lift alone
    N=$(nvidia-smi -L | wc -l)
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    Q="--gpus $N --no-cpu-baseline --no-depth --no-c4"
    HALGPU_GATHER_TIMELINE=1 timeout 600 $T --master-port 29611 bench.py $Q --steps 6 --warmup 3 > gpurun_out/bench_n_n${N}_timeline.json 2> gpurun_out/bench_n_n${N}_timeline.err
    grep "timeline rank 0" gpurun_out/bench_n_n${N}_timeline.err | tail -4
    timeout 600 $T --master-port 29612 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_n_n${N}.json 2> gpurun_out/bench_n_n${N}.err
    HALGPU_GATHER_NCCL=1 timeout 600 $T --master-port 29613 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_n_n${N}_nccl.json 2> gpurun_out/bench_n_n${N}_nccl.err
    HALGPU_GATHER_WIRE32=1 timeout 600 $T --master-port 29614 bench.py $Q --steps 20 --warmup 5 > gpurun_out/bench_n_n${N}_wire32.json 2> gpurun_out/bench_n_n${N}_wire32.err
    for f in "" _nccl _wire32; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n_n${N}$f.json').read().strip().splitlines()[-1])
print('n$N$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'kernels %.4f' % d['detail']['mapping_kernels_ms'], d['check'])
"; done
    ;;
o)  # short multi-GPU run (gpurun --gpus N): NCCL tests + the peer-memory gather forced in both wire forms + the default
    N=$(nvidia-smi -L | wc -l)
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    Q="--gpus $N --no-cpu-baseline --no-depth --no-c4 --steps 20 --warmup 5"
    if [ "$N" -le 2 ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_o.log 2>&1; tail -3 gpurun_out/pytest_o.log; fi
    if [ "$N" -le 2 ]; then
        HALGPU_GATHER_PULL=1 HALGPU_GATHER_TIMELINE=1 timeout 600 $T --master-port 29612 bench.py $Q > gpurun_out/bench_o_n${N}_pull.json 2> gpurun_out/bench_o_n${N}_pull.err
        grep "timeline rank 0" gpurun_out/bench_o_n${N}_pull.err | tail -3
        V="_pull _pull16 _nccl"
    else
        V="_pull16 _nccl"
    fi
    HALGPU_GATHER_PULL=1 HALGPU_GATHER_WIRE16=1 HALGPU_GATHER_TIMELINE=1 timeout 600 $T --master-port 29613 bench.py $Q > gpurun_out/bench_o_n${N}_pull16.json 2> gpurun_out/bench_o_n${N}_pull16.err
    grep "timeline rank 0" gpurun_out/bench_o_n${N}_pull16.err | tail -3
    HALGPU_GATHER_NCCL=1 timeout 600 $T --master-port 29614 bench.py $Q > gpurun_out/bench_o_n${N}_nccl.json 2> gpurun_out/bench_o_n${N}_nccl.err
    for f in $V; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_o_n${N}$f.json').read().strip().splitlines()[-1])
print('n$N$f', 'value %.4g' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'kernels %.4f' % d['detail']['mapping_kernels_ms'], d['check'])
"; done
    ;;
p)  # 2 GPUs: the default bench line at N=2 with the C4 secondary forced (communicator closed and re-made on a second file)
    N=$(nvidia-smi -L | wc -l)
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    HALGPU_BENCH_C4=1 timeout 500 $T --master-port 29612 bench.py --gpus $N --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_p_n${N}.json 2> gpurun_out/bench_p_n${N}.err
    tail -c 1800 gpurun_out/bench_p_n${N}.json; tail -3 gpurun_out/bench_p_n${N}.err
    ;;
q)  # final N=1 evidence: whole GPU tier (xdist), the default bench line, launch list, ncu --set full of the lane kernel
    ( time timeout 330 python -m pytest tests -q -m gpu -n 6 --durations=12 ) > gpurun_out/pytest_q.log 2>&1; tail -22 gpurun_out/pytest_q.log
    ( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
    tail -c 300 gpurun_out/bench_q.json; tail -4 gpurun_out/bench_q.err
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_q.csv \
        python bench.py --probe --no-divergent > /dev/null 2> gpurun_out/bench_ncu_q.err
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:'fastLiftKernel' -s 1 -c 1 -o gpurun_out/prof_q_fast -f \
        python bench.py --probe --no-divergent > /dev/null 2> gpurun_out/prof_q.err
    ;;
r)  # last N=1 evidence of the round: liftover GPU tests, the default bench line, ncu --set full of the lane kernel (32 registers again)
    timeout 200 python -m pytest tests/test_liftover_gpu.py -x -q -m gpu > gpurun_out/pytest_r.log 2>&1; tail -2 gpurun_out/pytest_r.log
    ( time timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
    tail -c 200 gpurun_out/bench_r.json; tail -4 gpurun_out/bench_r.err
    timeout 120 ncu --set full --clock-control none --import-source on -k regex:'fastLiftKernel' -s 1 -c 1 -o gpurun_out/prof_r_fast -f \
        python bench.py --probe --no-divergent > /dev/null 2> gpurun_out/prof_r.err
    ;;
*)  echo "unknown stage $stage"; exit 2;;
esac
