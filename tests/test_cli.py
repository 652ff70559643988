"""halLiftover CLI (host C++ mirror of hal::Liftover::convert + BedLine) -- byte-for-byte against the
reference CLI's outputs.  CPU tier: linked against the emulated library; GPU tier: the real binary."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, ref_bin

REF_GOLDENS = [("test1.bed3", "halLiftoverPsl3Test.psl", ["--outPSL"]), ("test1.bed12", "halLiftoverPsl12Test.psl", ["--outPSL"]),
               ("test1.bed3", "halLiftoverBed3Test.bed", []), ("test1.bed12", "halLiftoverBed12Test.bed", []),
               ("test1.bed12+2", "halLiftoverBed12ExtraTest.bed", []), ("test1.bed4+2", "halLiftoverBed4ExtraTest.bed", ["--bedType", "4"])]


def run_cli(cli, args, hal, src, inp, tgt, out):
    return subprocess.run([cli] + args + [hal, src, inp, tgt, out], capture_output=True, text=True)


def check_all(cli, golden_cases, tmp_path, pick):
    out = str(tmp_path / "o.bed")
    # liftover/Makefile:32-63 -- the reference's CLI tests
    for inp, exp, args in REF_GOLDENS:
        r = run_cli(cli, args, os.path.join(GOLDEN, "randgenSmallSeed0.hal"), "Genome_0", os.path.join(GOLDEN, "ref_liftover", inp), "Genome_2", out)
        assert r.returncode == 0, r.stderr
        assert open(out).read() == open(os.path.join(GOLDEN, "ref_liftover", exp)).read(), exp
    for c in golden_cases:
        if not pick(c):
            continue
        r = run_cli(cli, c["args"], os.path.join(GOLDEN, c["hal"]), c["src"], os.path.join(GOLDEN, "cases", c["name"] + ".in.bed"), c["tgt"], out)
        assert r.returncode == 0, r.stderr
        assert open(out).read() == open(os.path.join(GOLDEN, "cases", c["name"] + ".out.bed")).read(), c["name"]


def test_cli_emulated_matches_reference_outputs(emul_cli, golden_cases, tmp_path):
    check_all(emul_cli, golden_cases, tmp_path,
              lambda c: "bed12" in c["name"] or c["name"].startswith(("psl_", "coal_")) or c["name"].startswith("ref_") and "_all_" not in c["name"])


def test_cli_errors(emul_cli, tmp_path):
    hal = os.path.join(GOLDEN, "randgenSmallSeed0.hal")
    bed = tmp_path / "i.bed"
    out = str(tmp_path / "o.bed")
    bed.write_text("Genome_0_seq\t10\t5\n")
    r = run_cli(emul_cli, [], hal, "Genome_0", str(bed), "Genome_2", out)
    assert r.returncode == 1 and "Error zero or negative length BED range" in r.stderr and "in input bed line 1" in r.stderr
    bed.write_text("Genome_0_seq\t1\t5\nnope\t1\t5\nnope\t2\t9\nGenome_0_seq\t1\t999999999\n")
    r = run_cli(emul_cli, [], hal, "Genome_0", str(bed), "Genome_2", out)
    assert r.returncode == 0
    assert r.stderr.count("Unable to find sequence nope in genome Genome_0") == 1
    assert "Skipping interval with endpoint 999999999" in r.stderr
    r = run_cli(emul_cli, [], hal, "Genome_X", str(bed), "Genome_2", out)
    assert r.returncode == 1 and "srcGenome, Genome_X, not found in alignment" in r.stderr
    r = run_cli(emul_cli, [], str(tmp_path / "missing.hal"), "Genome_0", str(bed), "Genome_2", out)
    assert r.returncode == 1 and "can't open HAL file" in r.stderr
    r = run_cli(emul_cli, ["--bedType", "2"], hal, "Genome_0", str(bed), "Genome_2", out)
    assert r.returncode == 1 and "--bedType must be between 3 and 12" in r.stderr


@pytest.mark.gpu
def test_cli_cuda_matches_reference_outputs(golden_cases, tmp_path):
    from hal_b200 import build
    build.build()
    # every case except four of five of the 50 "all pairs" ones: each is a process with its own CUDA context (~1 s), and
    # test_liftover_gpu.py::test_cuda_golden_text lifts ALL of them in-process through the same library
    all_pairs = [c["name"] for c in golden_cases if "_all_" in c["name"]]
    keep = set(all_pairs[::5])
    check_all(os.path.join(ROOT, "hal_b200", "bin", "halLiftover"), golden_cases, tmp_path, lambda c: "_all_" not in c["name"] or c["name"] in keep)


def check_depth(cli, tmp_path):
    import json
    for c in json.load(open(os.path.join(GOLDEN, "cases", "depth_index.json"))):
        out = str(tmp_path / "d.wig")
        r = subprocess.run([cli, os.path.join(GOLDEN, c["hal"]), c["ref"], "--outWiggle", out] + c["args"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out).read() == open(os.path.join(GOLDEN, "cases", c["name"] + ".wig")).read(), c["name"]


def test_depth_cli_emulated_matches_reference_outputs(emul_depth_cli, tmp_path):
    check_depth(emul_depth_cli, tmp_path)
    r = subprocess.run([emul_depth_cli, os.path.join(GOLDEN, "varlen8.hal"), "A0", "--noAncestors"], capture_output=True, text=True)
    assert r.returncode == 1 and "--noAncestors cannot be used when reference genome (A0) is ancetral" in r.stderr


@pytest.mark.gpu
def test_depth_cli_cuda_matches_reference_outputs(tmp_path):
    from hal_b200 import build
    build.build()
    check_depth(os.path.join(ROOT, "hal_b200", "bin", "halAlignmentDepth"), tmp_path)


def check_maf(cli, tmp_path, whole_genome_hal=None):
    import json
    for k, c in enumerate(json.load(open(os.path.join(GOLDEN, "cases", "maf_index.json")))):
        out = str(tmp_path / "o.maf")
        if os.path.exists(out):
            os.remove(out)
        # the row text is decoded by a pool of threads from queued block descriptors: vary the pool and the queue size
        env = dict(os.environ, HALGPU_TEXT_THREADS=str(1 + k % 5), HALGPU_MAF_QUEUE_BYTES=str([1 << 28, 0, 5000, 200000][k % 4]))
        r = subprocess.run([cli, os.path.join(GOLDEN, c["hal"]), out] + c["args"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == open(os.path.join(GOLDEN, "cases", c["name"] + ".maf"), "rb").read(), c["name"]


def check_maf_targets(cli, tmp_path):
    """hal2maf --refTargets (MafBed, maf/impl/halMafBed.cpp): one convertSequence per BED interval / BED12 block; output, the
    per-line messages and the scanner's error text equal the reference's (tests/golden/make_golden_maf_targets.py)"""
    import json
    d = os.path.join(GOLDEN, "maf_targets")
    for c in json.load(open(os.path.join(d, "index.json"))):
        out = str(tmp_path / "t.maf")
        if os.path.exists(out):
            os.remove(out)
        r = subprocess.run([cli, os.path.join(GOLDEN, c["hal"]), out, "--refTargets", os.path.join(d, c["name"] + ".bed")] + c["args"],
                           capture_output=True, text=True)
        assert r.returncode == c.get("returncode", 0), r.stderr
        err = "".join(x + "\n" for x in r.stderr.splitlines() if not x.startswith("[halgpu"))
        assert err == c["stderr"], c["name"]
        want = os.path.join(d, c["name"] + ".maf")
        assert os.path.exists(out) == os.path.exists(want), c["name"]
        if os.path.exists(want):
            assert open(out, "rb").read() == open(want, "rb").read(), c["name"]


def test_maf_cli_emulated_matches_reference_outputs(emul_maf_cli, tmp_path):
    check_maf(emul_maf_cli, tmp_path)
    check_maf_targets(emul_maf_cli, tmp_path)
    r = subprocess.run([emul_maf_cli, os.path.join(GOLDEN, "varlen8.hal"), str(tmp_path / "x.maf"), "--refTargets", "x.bed", "--refSequence", "R_s0"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "unsupported when using BED input" in r.stderr
    r = subprocess.run([emul_maf_cli, os.path.join(GOLDEN, "varlen8.hal"), str(tmp_path / "x.maf"), "--noAncestors"], capture_output=True, text=True)
    assert r.returncode == 1 and "the --noAncestors option is invalid" in r.stderr
    r = subprocess.run([emul_maf_cli, os.path.join(GOLDEN, "varlen8.hal"), str(tmp_path / "x.maf"), "--printTree"], capture_output=True, text=True)
    assert r.returncode == 1 and "not implemented in the GPU build" in r.stderr


@pytest.mark.skipif(ref_bin("hal2maf") is None, reason="oracle/_ref not built")
def test_maf_cli_emulated_whole_genome_live(emul_maf_cli, tmp_path):
    """multi-sequence whole-genome export: 4 convertSequence calls sharing one MafBlock"""
    hal = os.path.join(GOLDEN, "varlen8.hal")
    for args in (["--refGenome", "L1"], ["--refGenome", "A2", "--maxBlockLen", "50"], ["--refGenome", "L0", "--unique"]):
        a, b = str(tmp_path / "a.maf"), str(tmp_path / "b.maf")
        subprocess.check_call([ref_bin("hal2maf"), hal, a] + args)
        subprocess.check_call([emul_maf_cli, hal, b] + args)
        assert open(a, "rb").read() == open(b, "rb").read(), args


def test_maf_cli_emulated_device_text_equals_host_text(emul_maf_cli, tmp_path):
    """the row text written by mafTextKernel (halgpu_maf_text) against the host formatter, with a queue small enough for
    many device calls and hand-offs to the writer thread"""
    hal = os.path.join(GOLDEN, "varlen8.hal")
    for args in (["--refGenome", "L1"], ["--refGenome", "R", "--maxBlockLen", "37"], ["--refGenome", "L2", "--noAncestors"]):
        outs = []
        for env in ({}, {"HALGPU_MAF_HOST_TEXT": "1"}, {"HALGPU_MAF_QUEUE_BYTES": "20000", "HALGPU_TEXT_THREADS": "3"}):
            o = str(tmp_path / "o.maf")
            subprocess.check_call([emul_maf_cli, hal, o] + args, env=dict(os.environ, **env))
            outs.append(open(o, "rb").read())
        assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 1000, args


@pytest.mark.skipif(ref_bin("hal2maf") is None, reason="oracle/_ref not built")
def test_maf_cli_emulated_unique_across_chunks(emul_maf_cli, tmp_path):
    """--unique keeps one visit cache per convertSequence sweep: the column range is processed in chunks, every chunk must
    classify its columns against the START OF THE SWEEP (halgpu_column_runs_in_sweep)."""
    hal = os.path.join(GOLDEN, "varlen8.hal")
    args = ["--refGenome", "L0", "--refSequence", "L0_s1", "--start", "9000", "--length", "7000", "--unique"]
    a, b = str(tmp_path / "a.maf"), str(tmp_path / "b.maf")
    subprocess.check_call([ref_bin("hal2maf"), hal, a] + args)
    subprocess.check_call([emul_maf_cli, hal, b] + args, env=dict(os.environ, HALGPU_MAF_CHUNK_COLUMNS="613"))
    assert open(a, "rb").read() == open(b, "rb").read()
    plain = str(tmp_path / "p.maf")
    subprocess.check_call([ref_bin("hal2maf"), hal, plain] + args[:-1])
    assert open(plain, "rb").read() != open(a, "rb").read(), "the window should contain reference paralogs"


@pytest.mark.gpu
def test_maf_cli_cuda_matches_reference_outputs(tmp_path):
    from hal_b200 import build
    build.build()
    cli = os.path.join(ROOT, "hal_b200", "bin", "hal2maf")
    check_maf(cli, tmp_path)
    check_maf_targets(cli, tmp_path)
    if ref_bin("hal2maf"):
        hal = os.path.join(GOLDEN, "varlen8.hal")
        for args in (["--refGenome", "L1"], ["--refGenome", "R"], ["--refGenome", "A2", "--maxBlockLen", "50"], ["--refGenome", "L3", "--noDupes"],
                     ["--refGenome", "L0", "--unique"], ["--refGenome", "A0", "--unique", "--onlyOrthologs"]):
            a, b = str(tmp_path / "a.maf"), str(tmp_path / "b.maf")
            subprocess.check_call([ref_bin("hal2maf"), hal, a] + args)
            subprocess.check_call([cli, hal, b] + args)
            assert open(a, "rb").read() == open(b, "rb").read(), args
