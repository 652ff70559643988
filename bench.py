#!/usr/bin/env python
"""bench.py -- liftover throughput of the B200 hot path on BASELINE.json's configs[1].

Workload (config.workload = "C2"): halRandGen-shaped 16-genome 4-level tree
(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R, 1,562,500 x 32 bp segments = 50 Mbp per genome
(written by hal_b200/bin/halSynth, branch length 0 == what halRandGen produces), 10 M BED3 intervals on L0_seq,
length U[50,2000], lifted L0 -> L7 (3 hops up, 2 down).  One "step" = one pass over the whole batch.

  value : input intervals / s, inputs resident in HBM, device-timed (CUDA events on the library's stream)
  e2e   : the same through halgpu_liftover with pinned HOST buffers (H2D of the batch + D2H of the result inside)
  roofline : algorithmic bytes (visit counts from the CPU oracle on a sample) / mapping-kernel time
  cpu_baseline / --impl reference : the reference's own halLiftover (oracle/_ref, built from /root/reference) on
             all host cores over a bounded sample of the same batch (the reference has no threads: one process per core)

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 via torch.distributed.run (one rank per GPU).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NEWICK = "(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R;"
SRC, TGT = "L0", "L7"
SEG_LEN = 32


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class Secondary:
    """A secondary measurement must never take the headline line down: an exception inside the block is recorded under the
    secondary's key as {"error": ...} and the bench goes on."""

    def __init__(self, line, key):
        self.line, self.key = line, key

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None or not issubclass(et, Exception):
            return False
        cur = self.line.get(self.key)
        if not isinstance(cur, dict):
            cur = self.line[self.key] = {}
        cur["error"] = ("%s: %s" % (et.__name__, ev))[:300]
        log("secondary", self.key, "failed:", cur["error"])
        return True


def make_intervals(n, genome_len, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    ln = rng.integers(50, 2001, n)
    gs = rng.integers(0, genome_len - 2 * SEG_LEN - ln)  # stay clear of the unaligned tail segment
    return gs.astype(np.int64), (gs + ln - 1).astype(np.int64)


def hal_path(segs, branch="0"):
    d = os.environ.get("HALB200_BENCH_DIR", os.path.join(tempfile.gettempdir(), "hal_b200_bench"))
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, f"c2_{segs}x{SEG_LEN}" + ("" if branch == "0" else f"_b{branch}") + ".hal")


def ensure_hal(segs, branch="0"):
    from hal_b200 import build
    build.build()
    p = hal_path(segs, branch)
    if not os.path.exists(p):
        t = time.time()
        subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", NEWICK, "--segs", str(segs),
                               "--segLen", str(SEG_LEN), "--branch", branch, "--seed", "7", p + ".tmp"])
        os.replace(p + ".tmp", p)
        log(f"[bench] wrote {p} ({os.path.getsize(p) / 1e9:.2f} GB) in {time.time() - t:.1f}s")
    return p


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region.

    NVML is queried in-process (pynvml, initialised in __init__, i.e. BEFORE the warm-up): starting an `nvidia-smi -lms`
    child right at the timed region cost it ~50 ms of driver stalls (its NVML start-up serialises with this process's
    cudaMallocAsync / cudaFree calls) and turned a 30 ms step into 80 ms.  `nvidia-smi` is only the fallback when pynvml
    is missing, and then the sampler waits for its first row before the timed region starts.  Only rank 0 samples."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu, enabled=True, period_s=0.02):
        self.rows, self.gpu, self.enabled, self.period = [], gpu, enabled, period_s
        self.nvml = self.handle = self.proc = self.thread = None
        self.stop = threading.Event()
        if not enabled:
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(gpu)
            try:  # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES is set: go through the PCI address
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.nvml = pynvml
            self._sample()  # first query (lazy driver paths) outside the timed region
            self.rows.clear()
        except Exception as e:  # noqa: BLE001
            log(f"[bench] pynvml unavailable ({e}); falling back to nvidia-smi")
            self.nvml = None

    def _sample(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits])

    def _loop(self):
        while not self.stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self.stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __enter__(self):
        if not self.enabled:
            return self
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:  # its start-up must not overlap the timed region
                time.sleep(0.02)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def algorithmic_bytes_per_interval(hal, gs, ge, n_src_segs, sample=20000):
    """24 B input + 8*ceil(log2(N+1)) search + sum of visited record bytes (+8 B successor start each)
    + 40 B per output line; visit counts from the instrumented CPU oracle on a sample of the batch."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(hal)
    r = o.liftover(o.genome_id(SRC), o.genome_id(TGT), gs[:sample], ge[:sample])
    s = r["stats"]
    m = min(sample, len(gs))
    per = 24 + 8 * math.ceil(math.log2(n_src_segs + 1)) + s["visitBytes"] / m + 40 * s["outLines"] / m
    o.close()
    return per, s


def reference_throughput(hal, gs, ge, seq_name, sample, cores):
    """Times oracle/_ref/halLiftover (the reference's own CLI, parse + map + print) as `cores` independent processes
    over a `cores`-way split of the first `sample` intervals; falls back to the single-threaded oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref", "halLiftover")
    sample = min(sample, len(gs))
    if os.path.exists(ref):
        d = tempfile.mkdtemp(prefix="halb200_ref_")
        per = (sample + cores - 1) // cores
        files = []
        for c in range(cores):
            lo, hi = c * per, min(sample, (c + 1) * per)
            if lo >= hi:
                break
            p = os.path.join(d, f"in{c}.bed")
            with open(p, "w") as f:
                f.write("".join(f"{seq_name}\t{gs[i]}\t{ge[i] + 1}\n" for i in range(lo, hi)))
            files.append(p)
        subprocess.run(["cat", hal], stdout=subprocess.DEVNULL)  # pre-fault the page cache
        t = time.time()
        procs = [subprocess.Popen([ref, hal, SRC, p, TGT, p + ".out"]) for p in files]
        rc = [p.wait() for p in procs]
        dt = time.time() - t
        assert all(r == 0 for r in rc), "reference halLiftover failed"
        lines = sum(sum(1 for _ in open(p + ".out")) for p in files)
        return dict(value=sample / dt, kind="reference", cores=len(files), seconds=dt, lines=lines,
                    sample=f"first {sample} intervals of the batch, {len(files)} processes of oracle/_ref/halLiftover (BED3 in, BED3 out)")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(hal)
    sample = min(sample, 200000)
    t = time.time()
    o.liftover(o.genome_id(SRC), o.genome_id(TGT), gs[:sample], ge[:sample])
    dt = time.time() - t
    return dict(value=sample / dt, kind="port", cores=1, seconds=dt,
                sample=f"first {sample} intervals of the batch, oracle/liboracle.so restatement, 1 thread")


def write_bed3(path, seq_name, gs, ge):
    try:
        import pyarrow as pa
        import pyarrow.csv as pacsv
        t = pa.table({"c": pa.array([seq_name] * len(gs)), "s": pa.array(gs), "e": pa.array(ge + 1)})
        pacsv.write_csv(t, path, pacsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
    except ImportError:
        with open(path, "w") as f:
            for lo in range(0, len(gs), 1 << 20):
                f.write("".join(f"{seq_name}\t{s}\t{e + 1}\n" for s, e in zip(gs[lo:lo + (1 << 20)].tolist(), ge[lo:lo + (1 << 20)].tolist())))


def cli_throughput(hal, gs, ge, seq_name):
    """hal_b200/bin/halLiftover on the whole batch written as a BED3 file (same arguments the reference CLI takes)."""
    d = tempfile.mkdtemp(prefix="halb200_cli_")
    inp, outp = os.path.join(d, "in.bed"), os.path.join(d, "out.bed")
    write_bed3(inp, seq_name, gs, ge)
    cli = os.path.join(ROOT, "hal_b200", "bin", "halLiftover")
    best = None
    for _ in range(2):  # the second run has the input file and the binary in the page cache
        t0 = time.time()
        r = subprocess.run([cli, hal, SRC, inp, TGT, outp], env=dict(os.environ, HALGPU_TIMING="1"), capture_output=True, text=True)
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr
        if best is None or dt < best[0]:
            best = (dt, [l for l in r.stderr.splitlines() if l.startswith("[halLiftover]")])
    out_bytes = os.path.getsize(outp)
    res = {"metric": "halLiftover_cli_lines_per_sec", "value": len(gs) / best[0], "unit": "BED lines/s", "seconds": best[0],
           "lines": len(gs), "in_bytes": os.path.getsize(inp), "out_bytes": out_bytes, "breakdown": best[1][0] if best[1] else None,
           "includes": "process start, CUDA context, open+stage, tokenise, halgpu_liftover (host buffers), format, file write"}
    os.remove(inp)
    os.remove(outp)
    os.rmdir(d)
    return res


def main():
    # stdout carries exactly one JSON line: anything native libraries print there (NCCL's version banner ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--intervals", type=int, default=10_000_000)
    ap.add_argument("--segs", type=int, default=1_562_500)
    ap.add_argument("--cpu-sample", type=int, default=0, help="intervals in the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-depth", action="store_true")
    ap.add_argument("--no-maf", action="store_true")
    ap.add_argument("--maf-columns", type=int, default=50_000_000)
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-wiggle", action="store_true")
    ap.add_argument("--no-divergent", action="store_true", help="skip the branch-length-0.05 variant of C2 (SURVEY 8(d): report both variants)")
    ap.add_argument("--wiggle-bases", type=int, default=50_000_000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    genome_len = args.segs * SEG_LEN
    cores = os.cpu_count() or 1
    config = {"workload": "C2: halRandGen-shaped 16-genome 4-level tree, %d x %d bp segments (%.0f Mbp/genome), "
                          "%d BED3 intervals U[50,2000] bp on L0_seq, L0->L7 (3 up, 2 down), dupes on"
                          % (args.segs, SEG_LEN, genome_len / 1e6, args.intervals),
              "intervals_per_gpu": args.intervals, "parallelism": f"index replicated, intervals sharded x{world}",
              "l2": "staged index (~1.7 GB) and the 10M-interval batch are far larger than the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        hal = ensure_hal(args.segs)
        gs, ge = make_intervals(args.intervals, genome_len, 2)
        sample = args.cpu_sample or min(1_000_000, max(2000, cores * 3000))
        vals = []
        for i in range(args.warmup + args.steps):
            r = reference_throughput(hal, gs, ge, SRC + "_seq", sample, cores)
            if i >= args.warmup:
                vals.append(r)
        dt = sum(v["seconds"] for v in vals) / len(vals)
        v = sample / dt
        print(file=real_stdout, flush=True, *[json.dumps({"impl": "reference", "metric": "liftover_intervals_per_sec", "value": v, "unit": "intervals/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "intervals/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"],
                                           "sample": vals[-1]["sample"]},
                          "e2e": {"value": v, "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})])
        return

    import numpy as np
    import torch
    import hal_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if rank == 0:
        hal = ensure_hal(args.segs)
    if dist:
        dist.barrier()
    hal = hal_path(args.segs)

    t0 = time.time()
    a = hal_b200.Alignment(hal, device=local)
    stage_s = time.time() - t0
    src, tgt = a.genome_id(SRC), a.genome_id(TGT)
    n = args.intervals
    gs, ge = make_intervals(n, genome_len, 2 + rank)  # every rank lifts its own shard of the (conceptual) N*n batch
    d_gs = torch.from_numpy(gs).cuda()
    d_ge = torch.from_numpy(ge).cuda()
    h_gs = torch.from_numpy(gs).pin_memory()
    h_ge = torch.from_numpy(ge).pin_memory()
    stream = torch.cuda.ExternalStream(a.stream)
    torch.cuda.synchronize()

    class _Arr:  # expose a raw device pointer to torch
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    last_info = {}

    def step_resident(gather):
        res = a.liftover_ptrs(src, tgt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True)
        if gather and dist:
            # one all-gather of the output interval buffer over NCCL (hal_b200/parallel.py)
            from hal_b200 import parallel
            offs = torch.as_tensor(_Arr(res.offsets_ptr, (n + 1) * 8), device="cuda").view(torch.int64)
            recs = torch.as_tensor(_Arr(res.recs_ptr, max(res.n_rec, 1) * 32), device="cuda")[: res.n_rec * 32]
            parallel.all_gather_records(offs[1:] - offs[:-1], recs)
        out = (res.n_rec, res.kernel_ms, res.launches, res.n_retry)
        last_info.update(fast_ms=res.fast_ms, n_complex=res.n_complex)
        res.close()
        return out

    def step_e2e():
        res = a.liftover_ptrs(src, tgt, n, h_gs.data_ptr(), h_ge.data_ptr(), None, 0, device=False)
        out = (res.n_rec, res.kernel_ms)
        res.close()
        return out

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, enabled=(rank == 0))  # NVML initialised here, before the warm-up
    for _ in range(args.warmup):
        nrec, *_ = step_resident(True)
    kms = []
    with sampler as clocks:
        # the sampler is started BEFORE the barrier: spawning nvidia-smi takes rank 0 ~0.1 s, and ranks that entered the
        # timed loop earlier would sit in the first all-gather waiting for it (their event time is what MAX-over-ranks reports)
        barrier()
        l0 = a.L.halgpu_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        # Every step ends with the library synchronising its own stream on the host, and the N>1 all-gather runs on
        # torch's stream: events on torch's current stream bracket all of it (device clock, includes in-step gaps).
        cur = torch.cuda.current_stream()
        e0.record(cur)
        step_wall = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            nrec, k, launches, nretry = step_resident(True)
            kms.append(k)
            step_wall.append((time.perf_counter() - ts) * 1e3)
        e1.record(cur)
        barrier()
        wall = time.perf_counter() - w0
        dev_ms = e0.elapsed_time(e1)
    launches_total = a.L.halgpu_launch_count() - l0
    # e2e through the host-buffer ABI call
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - w0) / args.steps

    # BASELINE.json configs[4] on N > 1 GPUs: the halAlignmentDepth sweep of the whole reference genome, one window per rank,
    # ONE all-gather of the per-column values (hal_b200/parallel.py); strong scaling: the sweep is the same 50 M columns
    depth_multi = None
    if dist and not args.no_depth:
        from hal_b200 import parallel
        lo, hi = parallel.shard_bounds(genome_len, world)[rank]
        d_win = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
        dms = []
        for i in range(2 + 3):
            barrier()
            cur = torch.cuda.current_stream()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record(cur)
            a.depth(src, lo, hi - 1, 1, (), 0, out_ptr=d_win.data_ptr())  # returns with the library's stream idle
            whole = parallel.all_gather_columns(d_win, genome_len)
            d1.record(cur)
            barrier()
            if i >= 2:
                dms.append(d0.elapsed_time(d1))
        t = torch.tensor([float(np.mean(dms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        depth_multi = {"metric": "alignment_depth_columns_per_sec", "value": genome_len / (float(t[0]) / 1e3), "unit": "columns/s",
                       "ms_per_sweep": float(t[0]), "columns": genome_len, "rows_per_column": 16, "scaling": "strong",
                       "parallelism": f"reference windows x{world}, one all-gather of {genome_len * 4} B",
                       "check": {"depth15_fraction": float((whole == 15).float().mean()), "gathered": int(whole.numel())}}
        del whole, d_win

    ms_step = max(dev_ms, 0.0) / args.steps
    if dist:
        t = torch.tensor([ms_step, e2e_s, float(np.mean(kms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s, kmean = (float(x) for x in t)
    else:
        kmean = float(np.mean(kms))
    if rank != 0:
        a.close()
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = world * n / (ms_step / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_interval, ostats = algorithmic_bytes_per_interval(hal, gs, ge, args.segs)
    achieved = per_interval * n / (kmean / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("liftoverKernel_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    line = {
        "metric": "liftover_intervals_per_sec", "value": value, "unit": "intervals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic", "config": config,
        "e2e": {"value": world * n / e2e_s, "unit": "intervals/s", "h2d_bytes_per_step": int(n * 16),
                "d2h_bytes_per_step": int((n + 1) * 8 + nrec * 32)},
        "gpu_launches": int(launches_total),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "liftoverKernel", "kernel_ms": kmean,
                     "algorithmic_bytes_per_interval": per_interval,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"},
        "detail": {"output_lines_per_step": int(nrec), "retry_intervals": int(nretry), "wall_s_per_step": wall / args.steps,
                   "stage_seconds": stage_s, "staged_bytes": a.staged_bytes, "kernel_share_of_step": kmean / ms_step, "fast_kernel_ms": last_info.get("fast_ms"), "complex_intervals": last_info.get("n_complex"), "step_wall_ms": [round(x, 3) for x in step_wall],
                   "oracle_sample_stats": ostats},
    }
    if depth_multi is not None:
        line["secondary"] = depth_multi
    # secondary (BASELINE.json configs[4] shape on one GPU): halAlignmentDepth column sweep, ref = leaf L0, all targets
    if world == 1 and not args.no_depth:
        with Secondary(line, "secondary"):
            d_out = torch.empty(genome_len, dtype=torch.int32, device="cuda")
            dk = []
            for i in range(2 + 3):
                _, ms = a.depth(src, 0, genome_len - 1, 1, (), 0, out_ptr=d_out.data_ptr())
                if i >= 2:
                    dk.append(ms)
            torch.cuda.synchronize()
            dms = float(np.mean(dk))
            line["secondary"] = {"metric": "alignment_depth_columns_per_sec", "value": genome_len / (dms / 1e3), "unit": "columns/s",
                                 "kernel_ms": dms, "columns": genome_len, "rows_per_column": 16,
                                 "check": {"depth15_fraction": float((d_out == 15).float().mean())}}
            if not args.no_cpu_baseline:
                ref = os.path.join(ROOT, "oracle", "_ref", "halAlignmentDepth")
                if os.path.exists(ref):
                    win = 100000
                    t0 = time.time()
                    procs = [subprocess.Popen([ref, hal, SRC, "--start", str(c * win), "--length", str(win)], stdout=subprocess.DEVNULL)
                             for c in range(cores)]
                    assert all(p.wait() == 0 for p in procs)
                    dt = time.time() - t0
                    line["secondary"]["cpu_baseline"] = {"value": cores * win / dt, "unit": "columns/s", "cores": cores, "kind": "reference",
                                                         "sample": f"{cores} processes of oracle/_ref/halAlignmentDepth, {win} columns each"}
    # secondary (BASELINE.json configs[2]): hal2maf block extraction, ref = root, through the product CLI (GPU column
    # runs + host block state machine + text), whole genome, output to a file on the box
    if world == 1 and not args.no_maf:
        with Secondary(line, "secondary_maf"):
            cli = os.path.join(ROOT, "hal_b200", "bin", "hal2maf")
            mcols = min(genome_len, args.maf_columns)
            outp = os.path.join(os.path.dirname(hal), "bench_out.maf")
            if os.path.exists(outp):
                os.remove(outp)
            t0 = time.time()
            subprocess.check_call([cli, hal, outp, "--refGenome", "R", "--refSequence", "R_seq", "--start", "0", "--length", str(mcols)])
            dt = time.time() - t0
            msize = os.path.getsize(outp)
            line["secondary_maf"] = {"metric": "hal2maf_columns_per_sec", "value": mcols / dt, "unit": "columns/s", "seconds": dt,
                                     "columns": mcols, "maf_bytes": msize, "includes": "open+stage (%.2f s), GPU column runs, host blocker, text, file write" % stage_s}
            if not args.no_cpu_baseline:
                ref = os.path.join(ROOT, "oracle", "_ref", "hal2maf")
                if os.path.exists(ref):
                    win = 40000
                    d = tempfile.mkdtemp(prefix="halb200_maf_")
                    t0 = time.time()
                    procs = [subprocess.Popen([ref, hal, os.path.join(d, f"o{c}.maf"), "--refGenome", "R", "--refSequence", "R_seq", "--start",
                                               str(c * win), "--length", str(win)]) for c in range(cores)]
                    assert all(p.wait() == 0 for p in procs)
                    dt = time.time() - t0
                    line["secondary_maf"]["cpu_baseline"] = {"value": cores * win / dt, "unit": "columns/s", "cores": cores, "kind": "reference",
                                                             "sample": f"{cores} processes of oracle/_ref/hal2maf, {win}-column windows (hal2mafMP style)"}
            os.remove(outp)
    # secondary ("report BOTH"): the divergent variant of C2 -- branch length 0.05, i.e. random-parent
    # transpositions (paralogy rings), inversions and insertions on every branch -- same batch, same direction
    if world == 1 and not args.no_divergent:
        try:
            dhal = ensure_hal(args.segs, "0.05")
            with hal_b200.Alignment(dhal, device=local) as b:
                bs, bt = b.genome_id(SRC), b.genome_id(TGT)
                best = None
                for i in range(4):
                    t0 = time.perf_counter()
                    res = b.liftover_ptrs(bs, bt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True)
                    dt = time.perf_counter() - t0
                    cur = (dt, res.kernel_ms, res.n_rec, res.n_retry, res.launches, res.fast_ms, res.n_complex)
                    res.close()
                    if i > 0 and (best is None or dt < best[0]):
                        best = cur
            line["secondary_divergent"] = {"metric": "liftover_intervals_per_sec", "value": n / best[0], "unit": "intervals/s",
                                           "workload": "C2 with --branch 0.05 (transpositions/paralogy rings, inversions, insertions)",
                                           "seconds": best[0], "kernel_ms": best[1], "output_lines": int(best[2]),
                                           "retry_intervals": int(best[3]), "launches": int(best[4]), "fast_kernel_ms": best[5], "complex_intervals": int(best[6])}
        except Exception as e:  # noqa: BLE001 -- a secondary line must not take the headline down
            line["secondary_divergent"] = {"error": str(e)[:300]}
    # secondary: halWiggleLiftover's mapping core: one value per base of L7, lifted to L0 through
    # halgpu_wiggle_liftover with HOST buffers (runs + values in, set target bases out); the pair is L7 -> L0 because the
    # reference's own halWiggleLiftover cannot map L0 -> L7 (its wrong turn at the MRCA, oracle/restate/wiggle.cpp)
    if world == 1 and not args.no_wiggle:
        with Secondary(line, "secondary_wiggle"):
            wsrc, wtgt = a.genome_id("L7"), a.genome_id("L0")
            nb = min(args.wiggle_bases, genome_len - 2 * SEG_LEN)
            run = 2048
            wf = np.arange(0, nb, run, dtype=np.int64)
            wl = np.minimum(wf + run - 1, nb - 1)
            wv = np.random.default_rng(9).random(nb) * 100.0
            best = None
            for i in range(3):
                t0 = time.perf_counter()
                wpos, wval, winfo = a.wiggle_liftover(wsrc, wtgt, wf, wl, wf.copy(), wv)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, winfo["kernel_ms"], len(wpos))
            line["secondary_wiggle"] = {"metric": "wiggle_liftover_bases_per_sec", "value": nb / best[0], "unit": "source bases/s",
                                        "seconds": best[0], "mapping_kernel_ms": best[1], "bases_in": int(nb), "bases_out": int(best[2]),
                                        "runs": int(len(wf)), "h2d_bytes": int(nb * 8 + len(wf) * 24), "d2h_bytes": int(best[2] * 16),
                                        "check": {"max_equals_input_max": bool(len(wval) and wval.max() <= wv.max()), "all_nonnegative": bool((wval >= 0).all())}}
            try:
                # roofline of the wiggle-mode kernel, same accounting as the headline: algorithmic bytes = per source base 8 B of value
                # + 16 B read-modify-write of the target key, + the index records the reference walk visits (oracle visit count on a
                # sample of the runs) + 24 B per run of input; DRAM traffic from the committed ncu capture
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                from pyoracle import Oracle
                o = Oracle(hal)
                k = min(200, len(wf))
                st_ = o.liftover(o.genome_id("L7"), o.genome_id("L0"), wf[:k], wl[:k])["stats"]
                o.close()
                per_base = 24.0 + st_["visitBytes"] / float((wl[:k] - wf[:k] + 1).sum()) + 24.0 / run
                ach = per_base * nb / (best[1] / 1e3) / 1e9
                wtraffic = None
                try:
                    wtraffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("wiggleKernel_dram_bytes_per_launch")
                except (OSError, ValueError):
                    pass
                line["secondary_wiggle"]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": wtraffic,
                                                        "kernel": "liftoverKernel<LIFT_WIG>", "kernel_ms": best[1], "algorithmic_bytes_per_source_base": per_base}
            except Exception as e:  # noqa: BLE001
                line["secondary_wiggle"]["roofline"] = {"error": str(e)[:200]}
            if not args.no_cpu_baseline:
                ref = os.path.join(ROOT, "oracle", "_ref", "halWiggleLiftover")
                if os.path.exists(ref):
                    win = 20000
                    d = tempfile.mkdtemp(prefix="halb200_wig_")
                    for c in range(cores):
                        with open(os.path.join(d, f"i{c}.wig"), "w") as f:
                            f.write(f"fixedStep chrom=L7_seq start={c * win + 1} step=1\n" + "".join(f"{x:.4f}\n" for x in wv[c * win:(c + 1) * win]))
                    t0 = time.time()
                    procs = [subprocess.Popen([ref, hal, "L7", os.path.join(d, f"i{c}.wig"), "L0", os.path.join(d, f"o{c}.wig")]) for c in range(cores)]
                    assert all(p.wait() == 0 for p in procs)
                    dt = time.time() - t0
                    line["secondary_wiggle"]["cpu_baseline"] = {"value": cores * win / dt, "unit": "source bases/s", "cores": cores, "kind": "reference",
                                                                "sample": f"{cores} processes of oracle/_ref/halWiggleLiftover, {win} fixedStep bases each (text in, text out)"}
    # secondary: the whole halLiftover CLI (text I/O at GPU rate) on the same batch as a BED3 file:
    # process start + CUDA context + open/stage + multi-threaded tokeniser + halgpu_liftover + multi-threaded printer + file write
    if world == 1 and not args.no_cli:
        with Secondary(line, "secondary_cli"):
            a.close()
            a = None
            line["secondary_cli"] = cli_throughput(hal, gs, ge, SRC + "_seq")
    if not args.no_cpu_baseline:
        sample = args.cpu_sample or min(2_000_000, max(2000, cores * 6000))
        r = reference_throughput(hal, gs, ge, SRC + "_seq", sample, cores)
        line["cpu_baseline"] = {"value": r["value"], "unit": "intervals/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    print(json.dumps(line), file=real_stdout, flush=True)
    if a is not None:
        a.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
