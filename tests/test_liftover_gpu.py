"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from helpers import assert_same_as_oracle, random_intervals, bed_to_batch, batch_to_bed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hal_b200
    return hal_b200


@pytest.mark.parametrize("hal,src,tgt,n,maxlen,flags", [
    ("varlen8.hal", "L0", "L3", 20000, 400, 0),
    ("varlen8.hal", "L3", "L1", 20000, 300, 1),
    ("varlen8.hal", "R", "L3", 5000, 2000, 0),
    ("varlen8.hal", "R", "A2", 5000, 2000, 0),
    ("varlen8.hal", "A0", "R", 10000, 500, 0),
    ("varlen8.hal", "A0", "L2", 10000, 500, 0),
    ("varlen8.hal", "A1", "L3", 10000, 500, 0),
    ("varlen8.hal", "A1", "A1", 3000, 500, 0),
    ("varlen8.hal", "L2", "A1", 10000, 500, 0),
    ("randgenSmallSeed0.hal", "Genome_0", "Genome_2", 5000, 900, 0),
    ("randgenSmallSeed0.hal", "Genome_3", "Genome_2", 5000, 900, 0),
    ("refBedLiftoverTest.hal", "leaf3", "leaf1", 2000, 60, 0),
    ("refBedLiftoverTest.hal", "root", "leaf2", 2000, 100, 0),
])
def test_cuda_equals_oracle(hb, oracle_lib, hal, src, tgt, n, maxlen, flags):
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        s, t = a.genome_id(src), a.genome_id(tgt)
        gs, ge, st = random_intervals(a.genome_length(s), n, maxlen, seed=n + len(src))
        off, recs, info = a.liftover(s, t, gs, ge, st, flags)
        assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge, st, no_dupes=bool(flags & 1)))
        assert info["launches"] > 0
        off2, recs2, _ = a.liftover(s, t, gs, ge, st, flags | hb.HALGPU_NO_SORT)
        assert np.array_equal(off, off2) and np.array_equal(recs, recs2)


def test_cuda_golden_text(hb, golden_cases):
    opened = {}
    try:
        for c in golden_cases:
            if "bed12" in c["name"] or c["name"].startswith("psl_"):
                continue  # BED12 regrouping is host C++ (tests/test_cli.py)
            a = opened.get(c["hal"]) or opened.setdefault(c["hal"], hb.Alignment(os.path.join(GOLDEN, c["hal"])))
            bed = open(os.path.join(GOLDEN, "cases", c["name"] + ".in.bed")).read()
            exp = open(os.path.join(GOLDEN, "cases", c["name"] + ".out.bed")).read()
            s, t = a.genome_id(c["src"]), a.genome_id(c["tgt"])
            rows, gs, ge, st = bed_to_batch(a.sequences(s), bed)
            lim = a.genome_id(c["args"][c["args"].index("--coalescenceLimit") + 1]) if "--coalescenceLimit" in c["args"] else -1
            off, recs, _ = a.liftover(s, t, gs, ge, st, 1 if "--noDupes" in c["args"] else 0, coalescence_limit=lim)
            assert batch_to_bed(rows, a.sequences(t), off, recs) == exp, c["name"]
    finally:
        for a in opened.values():
            a.close()


def test_cuda_reference_repo_golden_bed3(hb):
    with hb.Alignment(os.path.join(GOLDEN, "randgenSmallSeed0.hal")) as a:
        s, t = a.genome_id("Genome_0"), a.genome_id("Genome_2")
        bed = open(os.path.join(GOLDEN, "ref_liftover", "test1.bed3")).read()
        exp = open(os.path.join(GOLDEN, "ref_liftover", "halLiftoverBed3Test.bed")).read()
        rows, gs, ge, st = bed_to_batch(a.sequences(s), bed)
        off, recs, _ = a.liftover(s, t, gs, ge, st)
        assert batch_to_bed(rows, a.sequences(t), off, recs) == exp


def test_cuda_edge_cases(hb, oracle_lib):
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        s, t = a.genome_id("L0"), a.genome_id("L3")
        L = a.genome_length(s)
        # empty batch
        off, recs, _ = a.liftover(s, t, [], [])
        assert list(off) == [0] and len(recs) == 0
        # single base at both ends, whole genome in one interval (deep fan-out -> retry ladder)
        gs = np.array([0, L - 1, 0, 0], np.int64)
        ge = np.array([0, L - 1, L - 1, min(L - 1, 4999)], np.int64)
        st = np.frombuffer(b"+-+.", np.uint8)
        off, recs, info = a.liftover(s, t, gs, ge, st)
        assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge, st))
        assert info["n_retry"] >= 1
        # out-of-range is an error, not a crash
        with pytest.raises(hb.HalGpuError):
            a.liftover(s, t, [0], [L])
        with pytest.raises(hb.HalGpuError):
            a.liftover(99, t, [0], [1])


def test_cuda_synthetic_faithful_properties(hb, tmp_path):
    """halRandGen-faithful shape at a size the oracle cannot cover quickly: every leaf->leaf interval that avoids
    the unaligned last segment maps to exactly itself (identity alignment), so the result is checkable in closed form."""
    from hal_b200 import build
    build.build()
    hal = str(tmp_path / "synth.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick",
                           "(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R;", "--segs", "200000", "--segLen", "32", hal])
    with hb.Alignment(hal) as a:
        s, t = a.genome_id("L0"), a.genome_id("L7")
        n = 1_000_000
        L = a.genome_length(s) - 64  # the last two segments are unaligned by construction
        gs, ge, st = random_intervals(L, n, 2000, seed=5, strands=b"+")
        off, recs, info = a.liftover(s, t, gs, ge, st)
        assert np.array_equal(off, np.arange(n + 1, dtype=np.uint64))
        assert np.array_equal(recs["start"], gs) and np.array_equal(recs["end"], ge + 1)
        assert np.array_equal(recs["src_start"], gs) and (recs["strand"] == ord("+")).all()
        assert info["n_retry"] == 0


@pytest.mark.parametrize("hal,ref,flags,targets", [
    ("varlen8.hal", "L0", 0, ()), ("varlen8.hal", "L0", 1, ()), ("varlen8.hal", "L3", 2, ()), ("varlen8.hal", "A1", 0, ()),
    ("varlen8.hal", "R", 1, ()), ("varlen8.hal", "L1", 0, ("L3", "A0")), ("varlen8.hal", "L2", 4, ()), ("varlen8.hal", "A0", 5, ()),
    ("refBedLiftoverTest.hal", "leaf3", 1, ()), ("refBedLiftoverTest.hal", "root", 0, ()),
    ("randgenSmallSeed0.hal", "Genome_3", 0, ()), ("randgenSmallSeed0.hal", "Genome_0", 1, ()),
])
def test_cuda_depth_equals_oracle(hb, oracle_lib, hal, ref, flags, targets):
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        g = a.genome_id(ref)
        last = a.genome_length(g) - 1
        t = [a.genome_id(x) for x in targets]
        got, ms = a.depth(g, 0, last, 1, t, flags)
        exp, _ = o.depth(g, 0, last, 1, t, count_dupes=bool(flags & 1), no_ancestors=bool(flags & 2), no_dupes=bool(flags & 4))
        assert np.array_equal(got, exp)
        got5, _ = a.depth(g, 3, last, 5, t, flags)
        assert np.array_equal(got5, exp[3::5])
        with pytest.raises(hb.HalGpuError):
            a.depth(g, 0, last + 1)


def test_cuda_depth_synthetic_properties(hb, tmp_path):
    """Identity-aligned synthetic tree: every base of a leaf aligns to all 15 other genomes except in the last
    (unaligned) segment of each genome, where the column is the reference base alone."""
    hal = str(tmp_path / "synth.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick",
                           "(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R;", "--segs", "100000", "--segLen", "32", hal])
    with hb.Alignment(hal) as a:
        g = a.genome_id("L0")
        L = a.genome_length(g)
        d, _ = a.depth(g, 0, L - 1)
        assert (d[: L - 64] == 15).all() and (d[L - 32:] == 0).all()
        d2, _ = a.depth(g, 0, L - 1, flags=hb.HALGPU_NO_ANCESTORS)
        assert (d2[: L - 64] == 7).all()


def _tree(depth, prefix):
    return prefix if depth == 0 else "(" + _tree(depth - 1, prefix + "a") + "," + _tree(depth - 1, prefix + "b") + ")" + prefix


def test_cuda_deep_tree_equals_oracle(hb, oracle_lib, tmp_path):
    """63 genomes, 6 levels (BASELINE configs[3] shape, scaled down), halRandGen-style events (transpositions ->
    paralogy rings, inversions, insertions): leaf -> far leaf is 5 hops up + 5 down; compare with the oracle."""
    hal = str(tmp_path / "t63.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", _tree(5, "N") + ";", "--segs", "20000",
                           "--segLen", "24", "--branch", "0.08", "--seed", "5", hal])
    o = oracle_lib.Oracle(hal)
    with hb.Alignment(hal) as a:
        for src, tgt, n in (("Naaaaa", "Nbbbbb", 20000), ("Nababa", "Nabbbb", 20000), ("N", "Nbabab", 5000), ("Nbbbba", "N", 5000)):
            s, t = a.genome_id(src), a.genome_id(tgt)
            gs, ge, st = random_intervals(a.genome_length(s), n, 600, seed=n + len(src))
            off, recs, info = a.liftover(s, t, gs, ge, st)
            assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge, st))
        g = a.genome_id("Nabab")
        d, _ = a.depth(g, 0, 99999)
        e, _ = o.depth(g, 0, 99999)
        assert np.array_equal(d, e)


@pytest.mark.parametrize("src,tgt,flags", [("L0", "L3", 8), ("L3", "L1", 9), ("R", "L2", 8), ("A0", "R", 8), ("L1", "L1", 8), ("A1", "L0", 9)])
def test_cuda_column_liftover_equals_oracle(hb, oracle_lib, src, tgt, flags):
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        s, t = a.genome_id(src), a.genome_id(tgt)
        gs, ge, st = random_intervals(a.genome_length(s), 400, 250, seed=len(src) + flags)
        off, recs, _ = a.liftover(s, t, gs, ge, st, flags)
        for i in range(len(gs)):
            got = [(int(r["tgt_seq"]), int(r["start"]), int(r["end"]), chr(r["strand"])) for r in recs[off[i]:off[i + 1]]]
            assert got == o.column_liftover(s, t, int(gs[i]), int(ge[i]), chr(st[i]), no_dupes=bool(flags & 1)), i


def test_cuda_column_liftover_cli_vs_reference_class(tmp_path):
    """the reference's ColumnLiftover (driven by oracle/gen/halColumnLiftoverCli.cpp) on pairs where it terminates;
    it orders target sequences by pointer value, so lines are compared after sorting (north_star: bit-exact after sort)"""
    import random
    from conftest import ref_bin
    drv = ref_bin("halColumnLiftoverCli")
    if drv is None:
        pytest.skip("oracle/_ref not built")
    cli = os.path.join(ROOT, "hal_b200", "bin", "halLiftover")
    hal = os.path.join(GOLDEN, "varlen8.hal")
    import hal_b200
    with hal_b200.Alignment(hal) as a:
        for src, tgt, extra in (("L0", "L3", []), ("L3", "L1", ["--noDupes"]), ("R", "L2", [])):
            seqs = a.sequences(a.genome_id(src))
            rng = random.Random(len(src) + len(extra))
            lines = []
            for i in range(300):
                nm, st, ln = rng.choice(seqs)
                l = rng.randint(1, min(200, ln))
                x = rng.randint(0, ln - l)
                lines.append(f"{nm}\t{x}\t{x + l}\tn{i}\t0\t{rng.choice('+-.')}")
            bed = tmp_path / "in.bed"
            bed.write_text("\n".join(lines) + "\n")
            subprocess.check_call(["timeout", "120", drv, hal, src, str(bed), tgt, str(tmp_path / "ref.bed")] + extra)
            subprocess.check_call([cli, "--columnLiftover"] + extra + [hal, src, str(bed), tgt, str(tmp_path / "got.bed")])
            assert sorted(open(tmp_path / "ref.bed").read().splitlines()) == sorted(open(tmp_path / "got.bed").read().splitlines()), (src, tgt)


def _config1(tmp_path, branch):
    """BASELINE.json configs[0]: 3-genome linear tree, 1 Mbp per genome (31 250 x 32 bp), 10 k BED3 intervals of 50..2000 bp
    (python random.seed(1)), root -> leaf; SURVEY.md 8(d) "C1" """
    import random
    hal = str(tmp_path / f"c1_{branch}.hal")
    subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", "((G2)G1)G0;", "--segs", "31250", "--segLen", "32",
                           "--branch", branch, "--seed", "1", hal])
    rng = random.Random(1)
    glen = 31250 * 32
    gs, ge = [], []
    for _ in range(10000):
        ln = rng.randint(50, 2000)
        s = rng.randint(0, glen - ln)
        gs.append(s)
        ge.append(s + ln - 1)
    return hal, np.array(gs, np.int64), np.array(ge, np.int64)


@pytest.mark.parametrize("branch", ["0", "0.05"])
def test_cuda_config1_root_to_leaf(hb, oracle_lib, tmp_path, branch):
    from conftest import ref_bin
    hal, gs, ge = _config1(tmp_path, branch)
    o = oracle_lib.Oracle(hal)
    with hb.Alignment(hal) as a:
        s, t = a.genome_id("G0"), a.genome_id("G2")
        off, recs, info = a.liftover(s, t, gs, ge)
        assert_same_as_oracle(off, recs, o.liftover(s, t, gs, ge))
        if branch == "0":
            assert info["n_complex"] < 100  # collinear data: the one-lane-per-interval kernel does (nearly) all of it
        off2, recs2, info2 = a.liftover(s, t, gs, ge, None, hb.HALGPU_NO_FAST)
        assert info2["n_complex"] == 0 and np.array_equal(off, off2)
        for k in ("start", "end", "src_start", "tgt_seq", "strand", "src_strand"):
            assert np.array_equal(recs[k], recs2[k])
    ref = ref_bin("halLiftover")
    if ref is not None:  # the reference's own CLI on the same BED, byte for byte against the product CLI
        bed = tmp_path / "c1.bed"
        bed.write_text("".join(f"G0_seq\t{a}\t{b + 1}\n" for a, b in zip(gs.tolist(), ge.tolist())))
        subprocess.check_call([ref, hal, "G0", str(bed), "G2", str(tmp_path / "ref.bed")])
        subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halLiftover"), hal, "G0", str(bed), "G2", str(tmp_path / "got.bed")])
        assert open(tmp_path / "ref.bed").read() == open(tmp_path / "got.bed").read()


@pytest.mark.parametrize("hal,src,tgt,n,maxlen", [
    ("varlen8.hal", "L0", "L3", 20000, 400),
    ("varlen8.hal", "A0", "L2", 10000, 500),
    ("randgenSmallSeed0.hal", "Genome_3", "Genome_2", 5000, 900),
])
def test_cuda_fused_walk_equals_oracle(hb, oracle_lib, monkeypatch, hal, src, tgt, n, maxlen):
    """HALGPU_FUSE=1 (measurement switch, engine.cu): whole collinear runs per fragment, clashing intervals re-walked piece by
    piece; the lines equal the oracle's (n_frag, a diagnostic, counts fused pieces and is not compared)"""
    path = os.path.join(GOLDEN, hal)
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        s, t = a.genome_id(src), a.genome_id(tgt)
        gs, ge, st = random_intervals(a.genome_length(s), n, maxlen, seed=n + len(src))
        exp = o.liftover(s, t, gs, ge, st)
        monkeypatch.setenv("HALGPU_FUSE", "1")
        for _ in range(2):
            off, recs, _ = a.liftover(s, t, gs, ge, st)
            assert_same_as_oracle(off, recs, exp)


@pytest.mark.parametrize("env", [{"HALGPU_TILE_GRAB": "1"}, {"HALGPU_TILE_GRAB": "16"}, {"HALGPU_SORT_BITS": "8"}])
def test_cuda_order_switches_equal_default(hb, oracle_lib, monkeypatch, env):
    """sort granularity and tile hand-out of the lane kernel (HALGPU_SORT_BITS, HALGPU_TILE_GRAB): identical results"""
    path = os.path.join(GOLDEN, "varlen8.hal")
    o = oracle_lib.Oracle(path)
    with hb.Alignment(path) as a:
        s, t = a.genome_id("L2"), a.genome_id("A1")
        gs, ge, st = random_intervals(a.genome_length(s), 50000, 40, seed=5)
        off1, recs1, info1 = a.liftover(s, t, gs, ge, st)
        assert_same_as_oracle(off1, recs1, o.liftover(s, t, gs, ge, st))
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        off, recs, info = a.liftover(s, t, gs, ge, st)
        assert np.array_equal(off, off1) and np.array_equal(recs, recs1) and info["n_complex"] == info1["n_complex"]
