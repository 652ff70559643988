#!/usr/bin/env python
"""bench.py -- liftover throughput of the B200 hot path on BASELINE.json's configs[1] (and, at 8 GPUs, configs[3]).

Workload C2 (config.workload): halRandGen-shaped 16-genome 4-level tree
(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R, 1,562,500 x 32 bp segments = 50 Mbp per genome
(written by hal_b200/bin/halSynth, branch length 0 == what halRandGen produces), 10 M BED3 intervals on L0_seq,
length U[50,2000], lifted L0 -> L7 (3 hops up, 2 down).  One "step" = one pass over the whole batch.
Workload C4 (--config C4, and as secondary_c4 of the default 8-GPU run): 64 genomes / 6 levels, 3,125,000 x 32 bp =
100 Mbp per genome, 100 M intervals sharded over the ranks, leaf -> far leaf (5 up, 5 down).

  value : input intervals / s, inputs resident in HBM, device-timed (CUDA events on the library's stream).  At N > 1
          every rank lifts its own shard and the step ends with the all-gather of the output records (C++ / NCCL,
          include/halgpu.h: halgpu_liftover_allgather_begin/end; the gather of step k overlaps the lift of step k+1)
  e2e   : the same through halgpu_liftover with pinned HOST buffers (H2D of the batch + D2H of the result inside)
  roofline : SURVEY.md 8(d) algorithmic bytes (visit counts from the CPU oracle on a sample) / kernel time, next to the
          DRAM bytes ncu measures for that kernel in this very run (roofline.traffic)
  check : the first 20 k intervals of the TIMED batch against the CPU oracle record by record; at N > 1 the gathered
          records against single-GPU lifts of every shard
  cpu_baseline / --impl reference : the reference's own halLiftover (oracle/_ref, built from /root/reference) on
          all host cores over a bounded sample of the same batch (the reference has no threads: one process per core)

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 via torch.distributed.run (one rank per GPU).
"""
import argparse
import csv
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

SEG_LEN = 32


def _tree(depth, prefix):
    return prefix if depth == 0 else "(" + _tree(depth - 1, prefix + "a") + "," + _tree(depth - 1, prefix + "b") + ")" + prefix


WORKLOADS = {
    "C2": dict(newick="(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R;", src="L0", tgt="L7", segs=1_562_500, seed=7,
               intervals=10_000_000, genomes=16, levels=4, hops="3 up, 2 down", root="R"),
    # 63-node balanced binary tree of depth 5 plus one more leaf under the root: 64 genomes, 6 levels
    "C4": dict(newick="(" + _tree(4, "Na") + "," + _tree(4, "Nb") + ",Nx)N;", src="Naaaaa", tgt="Nbbbbb", segs=3_125_000, seed=11,
               intervals=100_000_000, genomes=64, levels=6, hops="5 up, 5 down", root="N"),
}


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


def hal_path(wl, segs, branch="0"):
    d = os.environ.get("HALB200_BENCH_DIR", os.path.join(tempfile.gettempdir(), "hal_b200_bench"))
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, f"{wl.lower()}_{segs}x{SEG_LEN}" + ("" if branch == "0" else f"_b{branch}") + ".hal")


def ensure_hal(wl, segs, branch="0"):
    from hal_b200 import build
    build.build()
    p = hal_path(wl, segs, branch)
    if not os.path.exists(p):
        t = time.time()
        w = WORKLOADS[wl]
        subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", w["newick"], "--segs", str(segs),
                               "--segLen", str(SEG_LEN), "--branch", branch, "--seed", str(w["seed"]), p + ".tmp"])
        os.replace(p + ".tmp", p)
        log(f"[bench] wrote {p} ({os.path.getsize(p) / 1e9:.2f} GB) in {time.time() - t:.1f}s")
    return p


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region.

    NVML is queried in-process (pynvml, initialised in __init__, i.e. BEFORE the warm-up): starting an `nvidia-smi -lms`
    child right at the timed region cost it ~50 ms of driver stalls (its NVML start-up serialises with this process's
    CUDA calls) and turned a 30 ms step into 80 ms.  `nvidia-smi` is only the fallback when pynvml is missing, and then
    the sampler waits for its first row before the timed region starts.  Only rank 0 samples."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu, enabled=True, period_s=0.004):
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


def oracle_sample(hal, src, tgt, gs, ge, n_src_segs, sample=20000):
    """The CPU oracle on the first `sample` intervals: its records (the parity check of the timed batch) and its visit counts,
    which define SURVEY.md 8(d)'s algorithmic bytes per interval: 24 B input + 8*ceil(log2(N+1)) search + sum of visited
    record bytes (+8 B successor start each) + 40 B per output line."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(hal)
    m = min(sample, len(gs))
    r = o.liftover(o.genome_id(src), o.genome_id(tgt), gs[:m], ge[:m])
    s = r["stats"]
    per = 24 + 8 * math.ceil(math.log2(n_src_segs + 1)) + s["visitBytes"] / m + 40 * s["outLines"] / m
    o.close()
    return per, s, r


def records_equal_oracle(r, off, recs):
    """True when CSR offsets and every record field the reference prints agree with the oracle's."""
    import numpy as np
    if not np.array_equal(np.asarray(off, dtype=np.uint64), r["offsets"]):
        return False
    return all(np.array_equal(recs[k], r[ok]) for k, ok in (("start", "start"), ("end", "end"), ("src_start", "srcStart"),
                                                             ("tgt_seq", "tgtSeq"), ("strand", "strand"), ("src_strand", "srcStrand")))


def reference_throughput(hal, src, tgt, gs, ge, seq_name, sample, cores):
    """Times oracle/_ref/halLiftover (the reference's own CLI, parse + map + print) as `cores` independent processes
    over a `cores`-way split of the first `sample` intervals; falls back to the single-threaded oracle port.  The time of
    the same processes on an EMPTY BED (exec + mmap open + genome lookup) is measured too and reported beside the value."""
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
        empty = os.path.join(d, "empty.bed")
        open(empty, "w").close()
        subprocess.run(["cat", hal], stdout=subprocess.DEVNULL)  # pre-fault the page cache

        def run(inputs):
            t = time.time()
            procs = [subprocess.Popen([ref, hal, src, p, tgt, p + f".out{i}"]) for i, p in enumerate(inputs)]
            rc = [p.wait() for p in procs]
            assert all(r == 0 for r in rc), "reference halLiftover failed"
            return time.time() - t
        t_empty = run([empty] * len(files))
        dt = run(files)
        lines = sum(sum(1 for _ in open(p + f".out{i}")) for i, p in enumerate(files))
        return dict(value=sample / dt, kind="reference", cores=len(files), seconds=dt, lines=lines, startup_seconds=t_empty,
                    sample=f"first {sample} intervals of the batch, {len(files)} processes of oracle/_ref/halLiftover (BED3 in, BED3 out), "
                           f"{per} intervals each; {t_empty:.2f} s of the {dt:.2f} s is process start + mmap open (empty-BED run)")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(hal)
    sample = min(sample, 200000)
    t = time.time()
    o.liftover(o.genome_id(src), o.genome_id(tgt), gs[:sample], ge[:sample])
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


def cli_throughput(hal, src, tgt, gs, ge, seq_name):
    """hal_b200/bin/halLiftover on the whole batch written as a BED3 file (same arguments the reference CLI takes)."""
    d = tempfile.mkdtemp(prefix="halb200_cli_")
    inp, outp = os.path.join(d, "in.bed"), os.path.join(d, "out.bed")
    write_bed3(inp, seq_name, gs, ge)
    cli = os.path.join(ROOT, "hal_b200", "bin", "halLiftover")
    best = None
    for _ in range(2):  # the second run has the input file and the binary in the page cache
        t0 = time.time()
        r = subprocess.run([cli, hal, src, inp, tgt, outp], env=dict(os.environ, HALGPU_TIMING="1"), capture_output=True, text=True)
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr
        if best is None or dt < best[0]:
            best = (dt, [l for l in r.stderr.splitlines() if l.startswith("[halLiftover]")])
    out_bytes = os.path.getsize(outp)
    res = {"metric": "halLiftover_cli_lines_per_sec", "value": len(gs) / best[0], "unit": "BED lines/s", "seconds": best[0],
           "lines": len(gs), "in_bytes": os.path.getsize(inp), "out_bytes": out_bytes, "breakdown": " | ".join(best[1]) if best[1] else None,
           "includes": "process start, CUDA context, open+stage, tokenise, halgpu_liftover (host buffers), format, file write"}
    os.remove(inp)
    os.remove(outp)
    os.rmdir(d)
    return res


def ncu_traffic(args):
    """DRAM bytes per launch of the mapping kernels, measured by ncu on a short probe run of THIS script (same HAL, same
    batch): {"fastLiftKernel": {...}, "liftoverKernel": {...} (largest launch: the divergent walk), "depthKernel": {...}} or {}
    when ncu is unavailable.  One replay pass (three metrics), so the probe takes seconds."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {}
    out = os.path.join(tempfile.mkdtemp(prefix="halb200_ncu_"), "t.csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:fastLiftKernel|liftoverKernel|depthKernel", "--csv", "--log-file", out, sys.executable, os.path.abspath(__file__), "--probe",
           "--config", args.config, "--intervals", str(args.intervals), "--segs", str(args.segs)] + (["--no-divergent"] if args.no_divergent else [])
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300, check=True)
        rows = list(csv.reader(open(out)))
    except Exception as e:  # noqa: BLE001
        log("[bench] ncu traffic probe failed:", e)
        return {}
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hi:
        return {}
    hdr = rows[hi[0]]
    ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = {}
    unit_scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
                  "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}
    for r in rows[hi[0] + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * unit_scale.get(r[ui].lower(), 1)
        d = per.setdefault(r[ii], {"name": r[ki]})
        if r[mi].startswith("dram__bytes"):
            d["bytes"] = d.get("bytes", 0) + v
        else:
            d["ms"] = v
    res = {}
    for d in per.values():  # launches come in program order: keep the LAST fast / depth launch and the LARGEST walk launch
        name = "fastLiftKernel" if "fastLiftKernel" in d["name"] else ("depthKernel" if "depthKernel" in d["name"] else "liftoverKernel")
        if name == "liftoverKernel" and name in res and res[name]["bytes"] >= d.get("bytes", 0):
            continue
        res[name] = {"bytes": d.get("bytes", 0), "ms_under_ncu": d.get("ms")}
    return res


def dev_view(ptr, nbytes):
    """a raw device pointer as a torch uint8 tensor (no copy)"""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Arr(), device="cuda")


def result_arrays(res, lo=0, hi=None):
    """(offsets, records) of intervals lo..hi of a device result, as numpy"""
    import numpy as np
    import torch
    import hal_b200
    hi = res.n if hi is None else hi
    off = dev_view(res.offsets_ptr, (res.n + 1) * 8).view(torch.int64)[lo:hi + 1].cpu().numpy()
    r0, r1 = int(off[0]), int(off[-1])
    recs = dev_view(res.recs_ptr, max(res.n_rec, 1) * 32)[r0 * 32:r1 * 32].cpu().numpy().view(hal_b200.REC_DTYPE)
    return (off - off[0]).astype(np.uint64), recs


def new_comm(a, dist, rank, world):
    """the C++ communicator of this rank: rank 0's 128-byte id travels out of band (torch.distributed), the rest is NCCL from C++"""
    import torch
    import hal_b200
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(hal_b200.Comm.unique_id(a.L)), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    return hal_b200.Comm(a, world, rank, uid.cpu().numpy().tobytes())


def run_steps(a, comm, src, tgt, n, d_gs, d_ge, k, keep_last=False, info=None):
    """k passes over the batch.  N > 1: begin(step i+1) is issued before end(step i), so the all-gather of one step
    overlaps the lift of the next; every step's gathered result is complete when this returns."""
    wall, kms, pending, last = [], [], None, None

    def note(res, is_last):
        nonlocal last
        if info is not None:
            info.update(n_rec=res.n_rec, kernel_ms=res.kernel_ms, launches=res.launches, n_retry=res.n_retry, fast_ms=res.fast_ms,
                        n_complex=res.n_complex)
        kms.append(res.kernel_ms)
        if keep_last and is_last:
            last = res
        else:
            res.close()
    for i in range(k):
        ts = time.perf_counter()
        if comm is None:
            note(a.liftover_ptrs(src, tgt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True), i == k - 1)
        else:
            h = comm.begin(src, tgt, n, d_gs.data_ptr(), d_ge.data_ptr())
            if pending is not None:
                note(comm.end(pending)[0], False)
            pending = h
        wall.append((time.perf_counter() - ts) * 1e3)
    if pending is not None:
        ts = time.perf_counter()
        note(comm.end(pending)[0], True)
        wall[-1] += (time.perf_counter() - ts) * 1e3
    return wall, kms, last


def main():
    # stdout carries exactly one JSON line: anything native libraries print there (NCCL's version banner ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--intervals", type=int, default=0, help="intervals per GPU (0: the workload's own; C4: 100 M / ranks)")
    ap.add_argument("--segs", type=int, default=0, help="segments per genome (0: the workload's own)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="intervals in the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-depth", action="store_true")
    ap.add_argument("--no-maf", action="store_true")
    ap.add_argument("--maf-columns", type=int, default=50_000_000)
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-wiggle", action="store_true")
    ap.add_argument("--no-divergent", action="store_true", help="skip the branch-length-0.05 variant of C2 (SURVEY 8(d): report both variants)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu probe that measures roofline.traffic")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: time the sharded lift alone (no all-gather)")
    ap.add_argument("--no-c4", action="store_true", help="8 GPUs: skip the BASELINE configs[3] secondary")
    ap.add_argument("--wiggle-bases", type=int, default=50_000_000)
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)  # the short run ncu_traffic() profiles
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = WORKLOADS[args.config]
    SRC, TGT = W["src"], W["tgt"]
    args.segs = args.segs or W["segs"]
    if not args.intervals:
        args.intervals = W["intervals"] if args.config == "C2" else W["intervals"] // max(world, 1)
    genome_len = args.segs * SEG_LEN
    cores = os.cpu_count() or 1
    config = {"workload": "%s: halRandGen-shaped %d-genome %d-level tree, %d x %d bp segments (%.0f Mbp/genome), "
                          "%d BED3 intervals per GPU U[50,2000] bp on %s_seq, %s->%s (%s), dupes on"
                          % (args.config, W["genomes"], W["levels"], args.segs, SEG_LEN, genome_len / 1e6, args.intervals, SRC, SRC, TGT, W["hops"]),
              "intervals_per_gpu": args.intervals,
              "parallelism": f"index replicated, intervals sharded x{world}" + ("" if world == 1 else (", no gather" if args.no_gather else ", one all-gather of the output records per step (NCCL, C++)")),
              "l2": "the staged index and the interval batch are far larger than the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        hal = ensure_hal(args.config, args.segs)
        gs, ge = make_intervals(args.intervals, genome_len, 2)
        sample = args.cpu_sample or min(4_000_000, max(50_000, cores * 50_000))  # >= 50 k intervals per process: the start-up share stays small
        vals = []
        for i in range(args.warmup + args.steps):
            r = reference_throughput(hal, SRC, TGT, gs, ge, SRC + "_seq", sample, cores)
            if i >= min(args.warmup, 1):
                vals.append(r)
            if sum(v["seconds"] for v in vals) > 120:  # bounded: the whole arm ends within a few minutes
                break
        dt = sum(v["seconds"] for v in vals) / len(vals)
        v = sample / dt
        print(file=real_stdout, flush=True, *[json.dumps({"impl": "reference", "metric": "liftover_intervals_per_sec", "value": v, "unit": "intervals/s",
                          "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "intervals/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"],
                                           "sample": vals[-1]["sample"], "runs": [round(sample / x["seconds"], 1) for x in vals]},
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
        hal = ensure_hal(args.config, args.segs)
    if dist:
        dist.barrier()
    hal = hal_path(args.config, args.segs)

    t0 = time.time()
    a = hal_b200.Alignment(hal, device=local)
    stage_s = time.time() - t0
    src, tgt = a.genome_id(SRC), a.genome_id(TGT)
    n = args.intervals
    gs, ge = make_intervals(n, genome_len, 2 + rank)  # every rank lifts its own shard of the (conceptual) N*n batch
    d_gs = torch.from_numpy(gs).cuda()
    d_ge = torch.from_numpy(ge).cuda()
    stream = torch.cuda.ExternalStream(a.stream)
    torch.cuda.synchronize()

    if args.probe:  # what ncu_traffic() profiles: two resident steps, one depth sweep, two steps on the divergent file; nothing timed
        for _ in range(2):
            a.liftover_ptrs(src, tgt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True).close()
        if args.config == "C2":
            d_out = torch.empty(genome_len, dtype=torch.int32, device="cuda")
            a.depth(src, 0, genome_len - 1, 1, (), 0, out_ptr=d_out.data_ptr())
        a.close()
        dp = hal_path(args.config, args.segs, "0.05")
        if not args.no_divergent and os.path.exists(dp):
            with hal_b200.Alignment(dp, device=local) as b:
                for _ in range(2):
                    b.liftover_ptrs(b.genome_id(SRC), b.genome_id(TGT), n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True).close()
        return

    comm = new_comm(a, dist, rank, world) if (dist and not args.no_gather) else None
    last_info = {}

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, enabled=(rank == 0))  # NVML initialised here, before the warm-up
    if comm is not None:
        # communicator priming, before the W warm-up steps: the first batches create the send slots and every rank maps the
        # other ranks' slots (CUDA IPC), which costs milliseconds per mapping, once
        run_steps(a, comm, src, tgt, n, d_gs, d_ge, 4)
    run_steps(a, comm, src, tgt, n, d_gs, d_ge, args.warmup)
    with sampler as clocks:
        barrier()
        l0 = a.L.halgpu_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        # every step ends with the library synchronising its own stream on the host (and, at N > 1, that stream waiting for
        # the communicator's): events on the library's stream bracket all of it on the device clock, in-step gaps included
        e0.record(stream)
        step_wall, kms, last = run_steps(a, comm, src, tgt, n, d_gs, d_ge, args.steps, keep_last=True, info=last_info)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        dev_ms = e0.elapsed_time(e1)
    launches_total = a.L.halgpu_launch_count() - l0

    # ---- parity of the timed batch ----
    check = {}
    per_interval = ostats = None
    if rank == 0:
        per_interval, ostats, oref = oracle_sample(hal, SRC, TGT, gs, ge, args.segs)
        m = min(20000, n)
        off, recs = result_arrays(last, 0, m)
        check["first_%d_intervals_equal_oracle" % m] = bool(records_equal_oracle(oref, off, recs))
    if comm is not None:
        # the gathered result of the last timed step against single-GPU lifts of every shard (every rank checks its own copy
        # of the whole batch; the shards are regenerated from their seeds)
        ok = last.n == world * n
        goff = dev_view(last.offsets_ptr, (last.n + 1) * 8).view(torch.int64)
        grec = dev_view(last.recs_ptr, max(last.n_rec, 1) * 32)
        for r in range(world):
            sg, se = make_intervals(n, genome_len, 2 + r)
            dsg, dse = torch.from_numpy(sg).cuda(), torch.from_numpy(se).cuda()
            one = a.liftover_ptrs(src, tgt, n, dsg.data_ptr(), dse.data_ptr(), None, 0, device=True)
            o1 = dev_view(one.offsets_ptr, (n + 1) * 8).view(torch.int64)
            r1 = dev_view(one.recs_ptr, max(one.n_rec, 1) * 32)[: one.n_rec * 32]
            seg = goff[r * n:(r + 1) * n + 1]
            ok = ok and bool(torch.equal(seg - seg[0], o1)) and bool(torch.equal(grec[int(seg[0]) * 32:int(seg[-1]) * 32], r1))
            one.close()
            del dsg, dse
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        check["gathered_equals_single_gpu_lifts_on_every_rank"] = bool(t[0] > 0.5)
    n_rec_last = last.n_rec
    last.close()

    # e2e through the host-buffer ABI call (every rank its own shard; the result lands in pinned host memory)
    h_gs = torch.from_numpy(gs).pin_memory()
    h_ge = torch.from_numpy(ge).pin_memory()
    nrec_local = 0

    def step_e2e():
        res = a.liftover_ptrs(src, tgt, n, h_gs.data_ptr(), h_ge.data_ptr(), None, 0, device=False)
        out = res.n_rec
        res.close()
        return out
    for _ in range(max(1, args.warmup - 1)):
        nrec_local = step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - w0) / args.steps
    # latency of the smallest call (what a one-interval-per-liftInterval binding pays): n = 1, host buffers
    one_s, one_e = np.array([gs[0]], np.int64), np.array([ge[0]], np.int64)
    for _ in range(50):
        a.liftover(src, tgt, one_s, one_e)
    w0 = time.perf_counter()
    for _ in range(500):
        a.liftover_ptrs(src, tgt, 1, one_s.ctypes.data, one_e.ctypes.data, None, 0, device=False).close()
    call_us = (time.perf_counter() - w0) / 500 * 1e6

    # BASELINE.json configs[4] on N > 1 GPUs: the halAlignmentDepth sweep of the whole reference genome, one window per rank,
    # ONE all-gather of the per-column values (hal_b200/parallel.py); strong scaling: the sweep is the same 50 M columns
    depth_multi = None
    if dist and not args.no_depth and args.config == "C2":
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

    # BASELINE.json configs[3] as a secondary of the 8-GPU run: its own file, staged after C2's context is gone
    c4 = None
    staged_bytes = a.staged_bytes
    if dist and (world == 8 or os.environ.get("HALGPU_BENCH_C4")) and args.config == "C2" and not args.no_c4 and comm is not None:
        comm.close()
        comm = None
        a.close()
        a = None
        del d_gs, d_ge
        c4 = run_c4(dist, rank, local, world)
    if rank != 0:
        if comm is not None:
            comm.close()
        if a is not None:
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
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    fast_ms = last_info.get("fast_ms") or 0.0
    n_complex = int(last_info.get("n_complex") or 0)
    dom_fast = fast_ms > 0 and n_complex * 50 < n  # which kernel dominates the mapping
    dom_ms = fast_ms if dom_fast else kmean
    # Algorithmic bytes per interval of the dominant kernel.  The warp-per-interval walk does the
    # reference walk piece by piece: SURVEY 8(d)'s figure with the oracle's visit counts.  The lane kernel maps a whole
    # collinear run per hop: what it MUST move through the HBM is the sorted work item (start + id|length, 16 B) and the
    # 32-byte output record of every interval, plus every FastRec of the path's transitions once (one 32-byte record per
    # segment-sized position bucket; the sorted batch shares them through L1/L2) -- that is the roofline's numerator.  The
    # bytes its loads and stores touch (16 + 32 per hop + 32 per interval, L1/L2-served) and the SURVEY figure of the
    # reference walk over the same time are reported next to it.
    n_hops = sum(int(x.split()[0]) for x in W["hops"].split(","))
    touched_bytes = 16 + 32 * n_hops + 32
    index_bytes = min(n * n_hops * 32, n_hops * args.segs * 32)
    own_bytes = (48 * n + index_bytes) / n if dom_fast else per_interval
    achieved = own_bytes * n / (dom_ms / 1e3) / 1e9
    ref_walk_gbs = per_interval * n / (dom_ms / 1e3) / 1e9
    touched_gbs = touched_bytes * n / (dom_ms / 1e3) / 1e9
    traffic = {}
    if world == 1 and not args.no_traffic:
        if not args.no_divergent and args.config == "C2":
            ensure_hal(args.config, args.segs, "0.05")  # so that the probe sees the divergent walk too
        a.close()  # the probe stages its own copy
        traffic = ncu_traffic(args)
        a = hal_b200.Alignment(hal, device=local)
    dom_name = "fastLiftKernel" if dom_fast else "liftoverKernel"
    tr = traffic.get(dom_name, {}).get("bytes")
    sw = sorted(step_wall)
    line = {
        "metric": "liftover_intervals_per_sec", "value": value, "unit": "intervals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic", "config": config,
        "e2e": {"value": world * n / e2e_s, "unit": "intervals/s", "h2d_bytes_per_step": int(n * 16),
                "d2h_bytes_per_step": int((n + 1) * 8 + nrec_local * 32)},
        "gpu_launches": int(launches_total),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": tr, "kernel": dom_name, "kernel_ms": dom_ms,
                     "algorithmic_bytes_per_interval": own_bytes,
                     "algorithmic_bytes_formula": ("compulsory HBM bytes: 16 B sorted work item + 32 B output record per interval + the FastRec tables of the "
                                                   "%d transitions once (%d B: one 32-byte record per segment-sized bucket)" % (n_hops, index_bytes)) if dom_fast
                                                  else "SURVEY 8(d): 24 B in + search + visited records (oracle visit counts) + 40 B per output line",
                     "touched": ({"bytes_per_interval": touched_bytes, "gbs": touched_gbs, "frac_of_hbm_peak": touched_gbs / peak,
                                  "note": "bytes the kernel's loads and stores touch (16 + 32 per hop + 32), served by L1/L2 where neighbouring intervals share a FastRec"}
                                 if dom_fast else None),
                     "reference_walk": {"bytes_per_interval": per_interval, "gbs": ref_walk_gbs, "frac": ref_walk_gbs / peak,
                                        "note": "SURVEY 8(d) bytes of the REFERENCE walk (one record per piece per hop, oracle visit counts) over this kernel's time"},
                     "peak_source": peak_src,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel, measured by a probe of this run" if tr else None,
                     "dram_gbs": (tr / (dom_ms / 1e3) / 1e9) if tr else None, "dram_frac": (tr / (dom_ms / 1e3) / 1e9 / peak) if tr else None,
                     "note": "achieved/frac: the bytes this kernel must move through the HBM (algorithmic_bytes_formula) over its CUDA-event time; "
                             "traffic / dram_gbs / dram_frac: what the HBM delivered, measured by ncu in this run; touched: the bytes its loads and "
                             "stores address (L1/L2-served); reference_walk: the same time against the bytes the reference's piece-by-piece walk touches"},
        "check": check,
        "detail": {"output_lines_per_step": int(n_rec_last), "retry_intervals": int(last_info.get("n_retry") or 0), "wall_s_per_step": wall / args.steps,
                   "stage_seconds": stage_s, "staged_bytes": staged_bytes, "mapping_kernels_ms": kmean, "kernel_share_of_step": kmean / ms_step,
                   "fast_kernel_ms": fast_ms, "complex_intervals": n_complex, "single_interval_call_us": call_us,
                   "step_wall_ms": {"median": sw[len(sw) // 2], "p95": sw[min(len(sw) - 1, int(0.95 * len(sw)))], "max": sw[-1], "all": [round(x, 3) for x in step_wall]},
                   "oracle_sample_stats": ostats},
    }
    if c4 is not None:
        line["secondary_c4"] = c4
    if depth_multi is not None:
        line["secondary"] = depth_multi
    # secondary (BASELINE.json configs[4] shape on one GPU): halAlignmentDepth column sweep, ref = leaf L0, all targets
    if world == 1 and not args.no_depth and args.config == "C2":
        with Secondary(line, "secondary"):
            d_out = torch.empty(genome_len, dtype=torch.int32, device="cuda")
            dk = []
            for i in range(3 + 5):
                _, ms = a.depth(src, 0, genome_len - 1, 1, (), 0, out_ptr=d_out.data_ptr())
                if i >= 3:
                    dk.append(ms)
            torch.cuda.synchronize()
            dms = float(np.mean(dk))
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from pyoracle import Oracle
            o = Oracle(hal)
            win = 200_000
            exp, _ = o.depth(o.genome_id(SRC), genome_len // 2, genome_len // 2 + win - 1)
            o.close()
            got = d_out[genome_len // 2: genome_len // 2 + win].cpu().numpy()
            # per column 4 B out + (records visited per reference segment) / segment length; one walk of a leaf column
            # of this tree visits 16 top records (40 B + 8 B successor start) and 15 bottom records (their stride + 8 B, ~48 B)
            per_col = 4.0 + (16 * 48.0 + 15 * 48.0) / SEG_LEN
            dtr = traffic.get("depthKernel", {}).get("bytes")
            line["secondary"] = {"metric": "alignment_depth_columns_per_sec", "value": genome_len / (dms / 1e3), "unit": "columns/s",
                                 "kernel_ms": dms, "columns": genome_len, "rows_per_column": 16,
                                 "check": {"depth15_fraction": float((d_out == 15).float().mean()),
                                           "window_of_%d_columns_equals_oracle" % win: bool(np.array_equal(got, exp))},
                                 "roofline": {"bound": "hbm", "achieved": per_col * genome_len / (dms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                              "frac": per_col * genome_len / (dms / 1e3) / 1e9 / peak, "traffic": dtr, "kernel": "depthKernel",
                                              "kernel_ms": dms, "algorithmic_bytes_per_column": per_col,
                                              "dram_frac": (dtr / (dms / 1e3) / 1e9 / peak) if dtr else None}}
            del d_out
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
    if world == 1 and not args.no_maf and args.config == "C2":
        with Secondary(line, "secondary_maf"):
            cli = os.path.join(ROOT, "hal_b200", "bin", "hal2maf")
            mcols = min(genome_len, args.maf_columns)
            outp = os.path.join(os.path.dirname(hal), "bench_out.maf")
            if os.path.exists(outp):
                os.remove(outp)
            root_args = ["--refGenome", W["root"], "--refSequence", W["root"] + "_seq"]
            t0 = time.time()
            r = subprocess.run([cli, hal, outp] + root_args + ["--start", "0", "--length", str(mcols)],
                               env=dict(os.environ, HALGPU_TIMING="1"), capture_output=True, text=True)
            dt = time.time() - t0
            assert r.returncode == 0, r.stderr[-300:]
            msize = os.path.getsize(outp)
            line["secondary_maf"] = {"metric": "hal2maf_columns_per_sec", "value": mcols / dt, "unit": "columns/s", "seconds": dt,
                                     "columns": mcols, "maf_bytes": msize, "breakdown": [l for l in r.stderr.splitlines() if l.startswith("[hal2maf]")][-1:],
                                     "includes": "process start, open+stage (%.2f s), GPU column runs, host blocker, text, file write" % stage_s}
            os.remove(outp)
            if not args.no_cpu_baseline:
                ref = os.path.join(ROOT, "oracle", "_ref", "hal2maf")
                if os.path.exists(ref):
                    win = 40000
                    d = tempfile.mkdtemp(prefix="halb200_maf_")
                    t0 = time.time()
                    procs = [subprocess.Popen([ref, hal, os.path.join(d, f"o{c}.maf")] + root_args + ["--start", str(c * win), "--length", str(win)])
                             for c in range(cores)]
                    assert all(p.wait() == 0 for p in procs)
                    dt = time.time() - t0
                    # parity at bench scale: the first reference window against the same columns from the product CLI
                    sub = os.path.join(d, "ours0.maf")
                    subprocess.check_call([cli, hal, sub] + root_args + ["--start", "0", "--length", str(win)])
                    same = open(sub, "rb").read() == open(os.path.join(d, "o0.maf"), "rb").read()
                    line["secondary_maf"]["check"] = {"first_%d_columns_equal_reference_hal2maf" % win: bool(same)}
                    line["secondary_maf"]["cpu_baseline"] = {"value": cores * win / dt, "unit": "columns/s", "cores": cores, "kind": "reference",
                                                             "sample": f"{cores} processes of oracle/_ref/hal2maf, {win}-column windows (hal2mafMP style)"}
    # secondary ("report BOTH"): the divergent variant of C2 -- branch length 0.05, i.e. random-parent
    # transpositions (paralogy rings), inversions and insertions on every branch -- same batch, same direction
    if world == 1 and not args.no_divergent and args.config == "C2":
        with Secondary(line, "secondary_divergent"):
            dhal = ensure_hal(args.config, args.segs, "0.05")
            with hal_b200.Alignment(dhal, device=local) as b:
                bs, bt = b.genome_id(SRC), b.genome_id(TGT)
                best = keep = None
                for i in range(5):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    res = b.liftover_ptrs(bs, bt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True)
                    dt = time.perf_counter() - t0
                    cur = (dt, res.kernel_ms, res.n_rec, res.n_retry, res.launches, res.fast_ms, res.n_complex)
                    if i >= 2 and (best is None or dt < best[0]):  # the first pass also learns the record pool size
                        best = cur
                    if i == 4:
                        keep = res
                    else:
                        res.close()
                dper, dstats, dref = oracle_sample(dhal, SRC, TGT, gs, ge, args.segs)
                m = min(20000, n)
                off, recs = result_arrays(keep, 0, m)
                dcheck = {"first_%d_intervals_equal_oracle" % m: bool(records_equal_oracle(dref, off, recs))}
                keep.close()
            dach = dper * n / (best[1] / 1e3) / 1e9
            dtr = traffic.get("liftoverKernel", {}).get("bytes")
            line["secondary_divergent"] = {"metric": "liftover_intervals_per_sec", "value": n / best[0], "unit": "intervals/s",
                                           "workload": "C2 with --branch 0.05 (transpositions/paralogy rings, inversions, insertions)",
                                           "seconds": best[0], "kernel_ms": best[1], "output_lines": int(best[2]),
                                           "retry_intervals": int(best[3]), "launches": int(best[4]), "fast_kernel_ms": best[5], "complex_intervals": int(best[6]),
                                           "check": dcheck,
                                           "roofline": {"bound": "hbm", "achieved": dach, "peak": peak, "unit": "GB/s", "frac": dach / peak, "traffic": dtr,
                                                        "kernel": "liftoverKernel<LIFT_BED> (+ scratch rung)", "kernel_ms": best[1], "algorithmic_bytes_per_interval": dper,
                                                        "dram_frac": (dtr / (best[1] / 1e3) / 1e9 / peak) if dtr else None,
                                                        "traffic_note": "traffic = the largest single launch (rung 1) under ncu"}}
            if not args.no_cpu_baseline:
                r = reference_throughput(dhal, SRC, TGT, gs, ge, SRC + "_seq", max(16000, cores * 1500), cores)
                line["secondary_divergent"]["cpu_baseline"] = {"value": r["value"], "unit": "intervals/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    # secondary: halWiggleLiftover's mapping core: one value per base of L7, lifted to L0 through
    # halgpu_wiggle_liftover with HOST buffers (runs + values in, set target bases out); the pair is L7 -> L0 because the
    # reference's own halWiggleLiftover cannot map L0 -> L7 (its wrong turn at the MRCA, oracle/restate/wiggle.cpp)
    if world == 1 and not args.no_wiggle and args.config == "C2":
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
                                        "check": {"identity_alignment_values_equal_input": bool(len(wval) == nb and np.array_equal(wval, wv[wpos]))}}
    # secondary: the whole halLiftover CLI (text I/O at GPU rate) on the same batch as a BED3 file:
    # process start + CUDA context + open/stage + multi-threaded tokeniser + halgpu_liftover + multi-threaded printer + file write
    if world == 1 and not args.no_cli:
        with Secondary(line, "secondary_cli"):
            a.close()
            a = None
            line["secondary_cli"] = cli_throughput(hal, SRC, TGT, gs, ge, SRC + "_seq")
    if not args.no_cpu_baseline:
        sample = args.cpu_sample or min(4_000_000, max(50_000, cores * 50_000))
        r = reference_throughput(hal, SRC, TGT, gs, ge, SRC + "_seq", sample, cores)
        line["cpu_baseline"] = {"value": r["value"], "unit": "intervals/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    print(json.dumps(line), file=real_stdout, flush=True)
    if comm is not None:
        comm.close()
    if a is not None:
        a.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_c4(dist, rank, local, world):
    """BASELINE.json configs[3]: 64 genomes / 6 levels / 100 Mbp per genome, 100 M intervals sharded over the 8 ranks, one
    all-gather per step.  Returns the secondary's dict on rank 0 (None elsewhere); failures are reported, not raised."""
    import numpy as np
    import torch
    import hal_b200
    W = WORKLOADS["C4"]
    out = {"metric": "liftover_intervals_per_sec", "workload": "C4: 64-genome 6-level tree, %d x %d bp segments (100 Mbp/genome), %d intervals over %d GPUs, %s->%s (%s)"
                     % (W["segs"], SEG_LEN, W["intervals"], world, W["src"], W["tgt"], W["hops"])}
    try:
        t0 = time.time()
        if rank == 0:
            ensure_hal("C4", W["segs"])
        dist.barrier()
        gen_s = time.time() - t0
        hal = hal_path("C4", W["segs"])
        t0 = time.time()
        a = hal_b200.Alignment(hal, device=local)
        stage_s = time.time() - t0
        src, tgt = a.genome_id(W["src"]), a.genome_id(W["tgt"])
        n = W["intervals"] // world
        glen = W["segs"] * SEG_LEN
        gs, ge = make_intervals(n, glen, 3 + rank)
        d_gs, d_ge = torch.from_numpy(gs).cuda(), torch.from_numpy(ge).cuda()
        comm = new_comm(a, dist, rank, world)
        stream = torch.cuda.ExternalStream(a.stream)
        run_steps(a, comm, src, tgt, n, d_gs, d_ge, 4 + 3)  # (communicator priming + warm-up)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record(stream)
        _, kms, last = run_steps(a, comm, src, tgt, n, d_gs, d_ge, K, keep_last=True)
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K, float(np.mean(kms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kmean = float(t[0]), float(t[1])
        # parity: this rank's shard of the gathered result against its own single-GPU lift, and a sample against the oracle
        one = a.liftover_ptrs(src, tgt, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True)
        goff = dev_view(last.offsets_ptr, (last.n + 1) * 8).view(torch.int64)
        seg = goff[rank * n:(rank + 1) * n + 1]
        o1 = dev_view(one.offsets_ptr, (n + 1) * 8).view(torch.int64)
        ok = last.n == world * n and bool(torch.equal(seg - seg[0], o1)) and bool(torch.equal(
            dev_view(last.recs_ptr, max(last.n_rec, 1) * 32)[int(seg[0]) * 32:int(seg[-1]) * 32], dev_view(one.recs_ptr, max(one.n_rec, 1) * 32)[: one.n_rec * 32]))
        tt = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        out.update(value=world * n / (ms / 1e3), unit="intervals/s", ms_per_step=ms, mapping_kernels_ms=kmean, n_gpus=world, intervals_per_gpu=n,
                   staged_bytes=a.staged_bytes, stage_seconds=stage_s, generate_seconds=gen_s, output_lines_per_step=int(last.n_rec),
                   check={"every_rank_shard_of_gather_equals_its_single_gpu_lift": bool(tt[0] > 0.5)})
        if rank == 0:
            m = 5000
            _, _, oref = oracle_sample(hal, W["src"], W["tgt"], gs, ge, W["segs"], sample=m)
            off, recs = result_arrays(one, 0, m)
            out["check"]["first_%d_intervals_equal_oracle" % m] = bool(records_equal_oracle(oref, off, recs))
        one.close()
        last.close()
        comm.close()
        a.close()
    except Exception as e:  # noqa: BLE001
        out["error"] = ("%s: %s" % (type(e).__name__, e))[:300]
        log("secondary_c4 failed:", out["error"])
    return out if rank == 0 else None


if __name__ == "__main__":
    main()
