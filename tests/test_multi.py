"""The C++ multi-GPU entry points (include/halgpu.h: halgpu_comm_*, halgpu_liftover_allgather_begin/end) on the host
harness: one thread per rank, each with its own context, joined by the harness communicator (hal_b200/csrc/comm.hpp).
The gathered result must equal one lift of the concatenated shards -- uniform shards take the plain all-gather, ragged
shards the per-rank broadcast group.  (The NCCL build of the same code runs in tests/test_multi_gpu.py, -m gpu.)"""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import random_intervals


def _run_ranks(emul_lib, hal, src, tgt, shards, flags=0, batches=1):
    import hal_b200
    world = len(shards)
    lib = hal_b200.load_library(emul_lib)
    uid = hal_b200.Comm.unique_id(lib)
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            a = hal_b200.Alignment(hal, lib_path=emul_lib)
            cm = hal_b200.Comm(a, world, r, uid)
            s, t = a.genome_id(src), a.genome_id(tgt)
            gs, ge, st = shards[r]
            handles = [cm.begin(s, t, len(gs), gs.ctypes.data, ge.ctypes.data, st.ctypes.data, flags) for _ in range(batches)]
            got = []
            for h in handles:  # begin k+1 before end k: the gather of one batch overlaps the lift of the next
                res, npr, nrr = cm.end(h)
                off = np.ctypeslib.as_array(C.cast(res.offsets_ptr, C.POINTER(C.c_uint64)), shape=(res.n + 1,)).copy()
                recs = np.frombuffer((C.c_char * (res.n_rec * 32)).from_address(res.recs_ptr), dtype=hal_b200.REC_DTYPE).copy() if res.n_rec else np.zeros(0, hal_b200.REC_DTYPE)
                got.append((off, recs, npr, nrr))
                res.close()
            out[r] = got
            cm.close()
            a.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    return out


@pytest.mark.parametrize("sizes,maxlen,wire", [
    ((120, 120), 12, "default"), ((150, 40), 200, "default"), ((0, 60), 100, "default"), ((70, 70, 70), 8, "default"),
    ((40, 50, 0, 60), 30, "default"),
    ((150, 40), 200, "HALGPU_GATHER_WIRE16"), ((70, 70, 70), 8, "HALGPU_GATHER_WIRE16"),
    ((150, 40), 200, "HALGPU_GATHER_NCCL"), ((40, 50, 0, 60), 30, "HALGPU_GATHER_NCCL"), ((0, 60), 100, "HALGPU_GATHER_NCCL"),
])
def test_emulated_allgather_equals_single_lift(emul_lib, monkeypatch, sizes, maxlen, wire):
    """(the records travel as they are below 4 ranks, in compact 16-byte form from 4 ranks up -- the first switch forces the latter; by default
    every rank copies the other ranks' shards out of their buffers, the last switch sends them through the all-gather collective)"""
    import hal_b200
    if wire != "default":
        monkeypatch.setenv(wire, "1")
    hal = os.path.join(GOLDEN, "varlen8.hal")
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L3")
    shards = [random_intervals(a.genome_length(s), n, maxlen, seed=7 + r) for r, n in enumerate(sizes)]
    gs, ge, st = (np.concatenate([sh[k] for sh in shards]) for k in range(3))
    off, recs, _ = a.liftover(s, t, gs, ge, st)
    a.close()
    out = _run_ranks(emul_lib, hal, "L0", "L3", shards, batches=2)
    for r in range(len(sizes)):
        for goff, grecs, npr, nrr in out[r]:
            assert npr == list(sizes)
            assert np.array_equal(goff, off), f"rank {r}: gathered offsets differ"
            assert np.array_equal(grecs, recs), f"rank {r}: gathered records differ"
            assert sum(nrr) == len(recs)


@pytest.mark.parametrize("wire32", [False, True])
def test_emulated_allgather_collinear_uniform_shards(emul_lib, tmp_path, monkeypatch, wire32):
    """equal shards of a collinear alignment: every interval ends in the one-lane-per-interval kernel, the offsets are the
    identity on every rank (not sent) and the records travel in compact 16-byte form -- or, with HALGPU_GATHER_WIRE32, as
    they are"""
    monkeypatch.setenv("HALGPU_GATHER_WIRE32" if wire32 else "HALGPU_GATHER_WIRE16", "1")
    import subprocess
    import hal_b200
    from conftest import ROOT
    from hal_b200 import build
    build.build()
    hal = str(tmp_path / "flat.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", "((L0,L1)A0,(L2)A1)R;", "--segs", "3000", "--segLen", "16", hal])
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L2")
    shards = [random_intervals(a.genome_length(s) - 32, 200, 300, seed=3 + r, strands=b"+-") for r in range(2)]
    gs, ge, st = (np.concatenate([sh[k] for sh in shards]) for k in range(3))
    off, recs, info = a.liftover(s, t, gs, ge, st)
    a.close()
    assert info["n_complex"] == 0 and np.array_equal(off, np.arange(401, dtype=np.uint64))
    out = _run_ranks(emul_lib, hal, "L0", "L2", shards)
    for r in range(2):
        goff, grecs, npr, nrr = out[r][0]
        assert np.array_equal(goff, off) and np.array_equal(grecs, recs)


def test_emulated_allgather_compact_wanted_but_records_do_not_fit(emul_lib, tmp_path, monkeypatch):
    """the compact wire form holds n_frag < 16: lines the warp-per-interval walk merges from more pieces do not fit, the ranks
    learn that from the headers and the batch travels as 32-byte records (through the collective) -- same result"""
    import subprocess
    import hal_b200
    from conftest import ROOT
    from hal_b200 import build
    build.build()
    monkeypatch.setenv("HALGPU_GATHER_WIRE16", "1")
    hal = str(tmp_path / "flat.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", "((L0,L1)A0,(L2)A1)R;", "--segs", "3000", "--segLen", "16", hal])
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L2")
    shards = [random_intervals(a.genome_length(s) - 32, 60 + 20 * r, 600, seed=5 + r) for r in range(2)]
    gs, ge, st = (np.concatenate([sh[k] for sh in shards]) for k in range(3))
    off, recs, info = a.liftover(s, t, gs, ge, st, hal_b200.HALGPU_NO_FAST)
    a.close()
    assert recs["n_frag"].max() >= 16
    out = _run_ranks(emul_lib, hal, "L0", "L2", shards, flags=hal_b200.HALGPU_NO_FAST, batches=2)
    for r in range(2):
        for goff, grecs, npr, nrr in out[r]:
            assert np.array_equal(goff, off) and np.array_equal(grecs, recs)
