"""CPU tier: the kernel SOURCE (hal_b200/csrc/liftover_kernel.cuh + engine) compiled for the host warp
emulator (tests/simt) must reproduce the oracle bit for bit -- covers the walk, the work pool, the retry
ladder and CSR assembly without a GPU.  (The product library is the sm_100a build; this is a harness.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_same_as_oracle, random_intervals, bed_to_batch, batch_to_bed


@pytest.mark.parametrize("hal,src,tgt,n,maxlen,flags", [
    ("varlen8.hal", "L0", "L3", 250, 300, 0),
    ("varlen8.hal", "L3", "L1", 200, 200, 1),
    ("varlen8.hal", "R", "L3", 120, 600, 0),
    ("varlen8.hal", "A0", "R", 150, 300, 0),
    ("varlen8.hal", "A1", "A1", 60, 300, 0),
    ("randgenSmallSeed0.hal", "Genome_0", "Genome_2", 150, 900, 0),
    ("refBedLiftoverTest.hal", "leaf3", "leaf1", 80, 40, 0),
])
def test_emulated_kernel_equals_oracle(emul_lib, oracle_lib, hal, src, tgt, n, maxlen, flags):
    import hal_b200
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id(src), a.genome_id(tgt)
    assert a.genomes == o.genomes
    assert a.sequences(s) == o.sequences(s)
    gs, ge, st = random_intervals(a.genome_length(s), n, maxlen, seed=n)
    off, recs, info = a.liftover(s, t, gs, ge, st, flags)
    assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge, st, no_dupes=bool(flags & 1)))
    # unsorted visiting order gives the same result
    off2, recs2, _ = a.liftover(s, t, gs, ge, st, flags | hal_b200.HALGPU_NO_SORT)
    assert np.array_equal(off, off2) and np.array_equal(recs, recs2)
    a.close()


def test_emulated_kernel_reference_golden_text(emul_lib, golden_cases):
    import hal_b200
    path = os.path.join(GOLDEN, "refBedLiftoverTest.hal")
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    for c in [c for c in golden_cases if c["name"].startswith("ref_") and "_all_" not in c["name"] and "bed12" not in c["name"] and not c["name"].startswith("psl_")]:
        bed = open(os.path.join(GOLDEN, "cases", c["name"] + ".in.bed")).read()
        exp = open(os.path.join(GOLDEN, "cases", c["name"] + ".out.bed")).read()
        s, t = a.genome_id(c["src"]), a.genome_id(c["tgt"])
        rows, gs, ge, st = bed_to_batch(a.sequences(s), bed)
        off, recs, _ = a.liftover(s, t, gs, ge, st, 1 if "--noDupes" in c["args"] else 0)
        assert batch_to_bed(rows, a.sequences(t), off, recs) == exp, c["name"]
    a.close()


@pytest.mark.parametrize("hal,ref,flags,targets", [
    ("varlen8.hal", "L0", 0, ()),
    ("varlen8.hal", "L0", 1, ()),
    ("varlen8.hal", "L3", 2, ()),
    ("varlen8.hal", "A1", 0, ()),
    ("varlen8.hal", "R", 1, ()),
    ("varlen8.hal", "L1", 0, ("L3", "A0")),
    ("varlen8.hal", "L2", 4, ()),
    ("refBedLiftoverTest.hal", "leaf3", 1, ()),
    ("randgenSmallSeed0.hal", "Genome_3", 0, ()),
])
def test_emulated_depth_equals_oracle(emul_lib, oracle_lib, hal, ref, flags, targets):
    import hal_b200
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    g = a.genome_id(ref)
    last = min(a.genome_length(g), 6000) - 1
    t = [a.genome_id(x) for x in targets]
    got, _ = a.depth(g, 0, last, 1, t, flags)
    exp, _ = o.depth(g, 0, last, 1, t, count_dupes=bool(flags & 1), no_ancestors=bool(flags & 2), no_dupes=bool(flags & 4))
    assert np.array_equal(got, exp)
    got3, _ = a.depth(g, 7, last, 3, t, flags)
    assert np.array_equal(got3, exp[7::3])
    a.close()


def test_emulated_host_pipeline_chunks(emul_lib, oracle_lib, monkeypatch):
    """halgpu_liftover cuts big host batches into chunks (copy/compute overlap); forced here on a small batch."""
    import hal_b200
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    monkeypatch.setenv("HALGPU_HOST_CHUNKS", "3")
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L3")
    gs, ge, st = random_intervals(a.genome_length(s), 100, 300, seed=3)
    off, recs, info = a.liftover(s, t, gs, ge, st, hal_b200.HALGPU_PSL)
    assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge, st))
    monkeypatch.delenv("HALGPU_HOST_CHUNKS")
    off1, recs1, info1 = a.liftover(s, t, gs, ge, st, hal_b200.HALGPU_PSL)
    assert np.array_equal(off, off1) and np.array_equal(recs, recs1) and np.array_equal(info["psl"], info1["psl"])
    a.close()


@pytest.mark.parametrize("src,tgt,flags", [("L0", "L3", 8), ("L3", "L1", 9), ("R", "L2", 8), ("A0", "R", 8), ("L1", "L1", 8), ("A1", "L0", 9)])
def test_emulated_column_liftover_equals_oracle(emul_lib, oracle_lib, src, tgt, flags):
    """HALGPU_COLUMN_LIFTOVER (hal::ColumnLiftover semantics) against the oracle's column walk"""
    import hal_b200
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id(src), a.genome_id(tgt)
    gs, ge, st = random_intervals(a.genome_length(s), 60, 250, seed=len(src) + flags)
    off, recs, _ = a.liftover(s, t, gs, ge, st, flags)
    for i in range(len(gs)):
        got = [(int(r["tgt_seq"]), int(r["start"]), int(r["end"]), chr(r["strand"])) for r in recs[off[i]:off[i + 1]]]
        exp = o.column_liftover(s, t, int(gs[i]), int(ge[i]), chr(st[i]), no_dupes=bool(flags & 1))
        assert got == exp, (i, gs[i], ge[i], chr(st[i]))
        assert (recs[off[i]:off[i + 1]]["src_start"] == -1).all()
    a.close()


COAL_CASES = [("L3", "L2", "R"), ("L3", "A2", "A1"), ("L3", "A2", "R"), ("L2", "L3", "R"), ("A2", "L3", "A1"), ("L1", "L1", "R"),
              ("L3", "L3", "A2"), ("A1", "L2", "R"), ("L0", "L1", "R"), ("L1", "A0", "R")]


@pytest.mark.parametrize("src,tgt,limit", COAL_CASES)
def test_emulated_coalescence_limit_equals_oracle(emul_lib, oracle_lib, src, tgt, limit):
    """halLiftover --coalescenceLimit (mapRecursiveParalogies): the extended path of the kernel
    against the oracle, whose restatement of it is pinned on the reference CLI (tests/test_oracle.py)."""
    import hal_b200
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t, lim = a.genome_id(src), a.genome_id(tgt), a.genome_id(limit)
    gs, ge, st = random_intervals(a.genome_length(s), 90, 250, seed=len(src + tgt + limit))
    off, recs, _ = a.liftover(s, t, gs, ge, st, 0, coalescence_limit=lim)
    exp = o.liftover(s, t, gs, ge, st, coalescence_limit=lim)
    assert_same_as_oracle(off, recs, exp)
    base = o.liftover(s, t, gs, ge, st)
    assert len(exp["start"]) > len(base["start"]), "the limit should add paralogous lines on this fixture"
    # with --noDupes the limit is ignored (halSegmentMapper.cpp:619)
    off1, recs1, _ = a.liftover(s, t, gs, ge, st, 1, coalescence_limit=lim)
    assert_same_as_oracle(off1, recs1, o.liftover(s, t, gs, ge, st, no_dupes=True))
    a.close()


def test_emulated_coalescence_limit_must_be_an_ancestor(emul_lib):
    import hal_b200
    a = hal_b200.Alignment(os.path.join(GOLDEN, "varlen8.hal"), lib_path=emul_lib)
    gs, ge, st = random_intervals(a.genome_length(a.genome_id("L3")), 5, 100, seed=1)
    with pytest.raises(hal_b200.HalGpuError, match="Hit root genome when attempting to map paralogies"):
        a.liftover(a.genome_id("L3"), a.genome_id("L2"), gs, ge, st, 0, coalescence_limit=a.genome_id("A0"))
    a.close()


def test_emulated_reference_coalescence_limit_kat(emul_lib):
    """the same known-answer test (halMappedSegmentTest.cpp:478-613) through the kernel"""
    import hal_b200
    a = hal_b200.Alignment(os.path.join(GOLDEN, "refMapExtraParalogsTest.hal"), lib_path=emul_lib)
    s, t, root = a.genome_id("grandChild2"), a.genome_id("grandChild1"), a.genome_id("root")
    gs, ge = np.array([0], np.int64), np.array([2], np.int64)
    off, recs, _ = a.liftover(s, t, gs, ge)
    assert [(int(r["start"]), int(r["end"]), chr(r["strand"])) for r in recs] == [(0, 3, "-")]
    off, recs, _ = a.liftover(s, t, gs, ge, coalescence_limit=root)
    # the three paralogous target segments are adjacent and collinear on the reverse strand but share ONE source piece:
    # extractSegment cannot merge them (source deltas differ), so three lines
    assert sorted((int(r["start"]), int(r["end"]), chr(r["strand"])) for r in recs) == [(0, 3, "-"), (3, 6, "-"), (6, 9, "-")]
    a.close()


@pytest.mark.parametrize("branch", ["0", "0.02"])
def test_emulated_config1_shape_and_fast_kernel(emul_lib, oracle_lib, tmp_path, branch):
    """BASELINE.json configs[0] shape scaled down (3-genome linear tree, root -> leaf) through both mapping kernels: the
    one-lane-per-interval kernel (collinear intervals) and the walk (HALGPU_NO_FAST) must agree with the oracle"""
    import subprocess
    import hal_b200
    from conftest import ROOT
    from hal_b200 import build
    build.build()
    hal = str(tmp_path / "c1.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", "((G2)G1)G0;", "--segs", "4000", "--segLen", "32",
                           "--branch", branch, "--seed", "1", hal])
    o = oracle_lib.Oracle(hal)
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("G0"), a.genome_id("G2")
    gs, ge, st = random_intervals(a.genome_length(s) - 64, 500, 900, seed=1)
    exp = o.liftover(s, t, gs, ge, st)
    off, recs, info = a.liftover(s, t, gs, ge, st)
    assert_same_as_oracle(off, recs, exp)
    assert info["n_complex"] < (5 if branch == "0" else 500)
    off2, recs2, info2 = a.liftover(s, t, gs, ge, st, hal_b200.HALGPU_NO_FAST)
    assert_same_as_oracle(off2, recs2, exp)
    assert info2["n_complex"] == 0
    a.close()


def test_emulated_sliced_sort_equals_unsliced(emul_lib, oracle_lib, monkeypatch):
    """large batches are sorted in slices on a second stream while the lane kernel runs the previous slice (engine.cu);
    forced here on a small batch: same result as the single-slice run and the oracle"""
    import hal_b200
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id("L2"), a.genome_id("A1")
    gs, ge, st = random_intervals(a.genome_length(s), 333, 40, seed=12)
    exp = o.liftover(s, t, gs, ge, st)
    off1, recs1, info1 = a.liftover(s, t, gs, ge, st)
    for k in ("2", "3", "4"):
        monkeypatch.setenv("HALGPU_SLICES", k)
        off, recs, info = a.liftover(s, t, gs, ge, st)
        assert_same_as_oracle(off, recs, exp)
        assert np.array_equal(recs, recs1) and info["n_complex"] == info1["n_complex"] and info["n_complex"] < 333
    a.close()


@pytest.mark.parametrize("hal,src,tgt,n,maxlen", [
    ("varlen8.hal", "L0", "L3", 250, 300),
    ("varlen8.hal", "A0", "L2", 150, 500),
    ("randgenSmallSeed0.hal", "Genome_3", "Genome_2", 100, 900),
])
def test_emulated_fused_walk_equals_oracle(emul_lib, oracle_lib, monkeypatch, hal, src, tgt, n, maxlen):
    """HALGPU_FUSE=1 (measurement switch, engine.cu): fragments follow whole collinear runs; intervals whose fused fragments
    clash in the target are walked again piece by piece -- the lines are the oracle's either way (n_frag, a diagnostic, is not)"""
    import hal_b200
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id(src), a.genome_id(tgt)
    gs, ge, st = random_intervals(a.genome_length(s), n, maxlen, seed=n)
    exp = o.liftover(s, t, gs, ge, st)
    monkeypatch.setenv("HALGPU_FUSE", "1")
    for _ in range(2):  # (the second call sizes its record pool from the first one's history)
        off, recs, _ = a.liftover(s, t, gs, ge, st)
        assert_same_as_oracle(off, recs, exp)
    a.close()


@pytest.mark.parametrize("env", [{"HALGPU_TILE_GRAB": "1"}, {"HALGPU_TILE_GRAB": "16"}, {"HALGPU_SORT_BITS": "8"}])
def test_emulated_order_switches_equal_default(emul_lib, oracle_lib, monkeypatch, env):
    """the sort granularity and the way the lane kernel's warps take their tiles (4 per atomicAdd by default) never change a result"""
    import hal_b200
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    a = hal_b200.Alignment(path, lib_path=emul_lib)
    s, t = a.genome_id("L2"), a.genome_id("A1")
    gs, ge, st = random_intervals(a.genome_length(s), 700, 40, seed=5)
    exp = o.liftover(s, t, gs, ge, st)
    off1, recs1, info1 = a.liftover(s, t, gs, ge, st)
    assert_same_as_oracle(off1, recs1, exp)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    off, recs, info = a.liftover(s, t, gs, ge, st)
    assert np.array_equal(off, off1) and np.array_equal(recs, recs1) and info["n_complex"] == info1["n_complex"]
    a.close()
