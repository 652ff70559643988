"""The multi-threaded BED text layer of the halLiftover CLI (hal_b200/csrc/host/bed_fast.cpp).

 * volume: millions of lines through the CLI's text code linked against tests/cpp/halgpu_stub.cpp (a synthetic mapping,
   no GPU): the multi-threaded path must print exactly what the serial BedLine path prints, for every BED width, every
   thread count and block size, with blank lines / missing sequences / out-of-range lines / dialect corners mixed in
 * parity: the same messy corpora through the emulated CLI (real kernel sources on the host warp emulator) against the
   reference's own halLiftover binary (oracle/_ref), stdout and stderr
"""
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT, ref_bin

HOST = os.path.join(ROOT, "hal_b200", "csrc", "host")


@pytest.fixture(scope="session")
def stub_cli():
    d = os.path.join(ROOT, "tests", "cpp")
    lib, cli = os.path.join(d, "libhalgpu_stub.so"), os.path.join(d, "halLiftover_stub")
    srcs = [os.path.join(HOST, f) for f in ("halLiftoverMain.cpp", "gpu_liftover.cpp", "bed.cpp", "bed_fast.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in ("gpu_liftover.hpp", "bed.hpp", "bed_fast.hpp")] + [os.path.join(d, "halgpu_stub.cpp")]
    if not os.path.exists(cli) or any(os.path.getmtime(x) > os.path.getmtime(cli) for x in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", lib, os.path.join(d, "halgpu_stub.cpp")])
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", cli] + srcs + ["-L" + d, "-lhalgpu_stub", "-Wl,-rpath,$ORIGIN"])
    return cli


def run(cli, hal, src, bed_path, tgt, out_path, args=(), threads=None, block=None):
    env = dict(os.environ, HALGPU_TIMING="1")
    env.pop("HALGPU_TEXT_THREADS", None)
    env.pop("HALGPU_BLOCK_BYTES", None)
    if threads is not None:
        env["HALGPU_TEXT_THREADS"] = str(threads)
    if block is not None:
        env["HALGPU_BLOCK_BYTES"] = str(block)
    r = subprocess.run([cli] + list(args) + [hal, src, bed_path, tgt, out_path], capture_output=True, text=True, env=env)
    timing = [l for l in r.stderr.splitlines() if l.startswith("[halLiftover] lines in")]
    stderr = "\n".join(l for l in r.stderr.splitlines() if not l.startswith(("[halLiftover]", "[halgpu timing]")))
    fast = int(timing[0].split("(")[1].split()[0]) if timing else -1
    return r.returncode, stderr, fast


def bed_line(rng, seqs, width, i, extras=0, strict=True):
    nm, ln = rng.choice(seqs)
    L = rng.randint(1, min(400, ln))
    a = rng.randint(0, ln - L)
    cols = [nm, str(a), str(a + L), f"n{i}" if rng.random() < 0.9 else "", str(rng.randint(0, 1000)), rng.choice("+-."),
            str(rng.choice([0, 0, a, a + 1])), str(rng.choice([0, 0, a + L])), rng.choice(["0", "255,0,0", "1,2", "7,8,9"])][:width]
    if not strict:  # corners BedLine::parse tolerates and the strict tokeniser hands to the serial path
        k = rng.randrange(6)
        if k == 0:
            cols[1] = " " + cols[1]
        elif k == 1:
            cols[2] = "+" + cols[2]
        elif k == 2 and width > 4:
            cols[4] = cols[4] + "abc"
        elif k == 3 and width > 5:
            cols[5] = cols[5] + "x"
        elif k == 4 and width > 8:
            cols[8] = cols[8] + ","
        else:
            cols[2] = "00" + cols[2]  # (still strict: leading zeros are plain digits)
    cols += [f"e{j}_{i}" if rng.random() < 0.8 else "" for j in range(extras)]
    if cols[-1] == "":
        cols[-1] = "z"  # a trailing empty column would be dropped by chopString; covered separately
    return "\t".join(cols)


def corpus(rng, seqs, width, n, extras=0, noise=True, missing=("nope", "chrZ"), strict=True):
    lines = []
    for i in range(n):
        lines.append(bed_line(rng, seqs, width, i, extras, strict or rng.random() < 0.97))
        if noise:
            r = rng.random()
            if r < 0.02:
                lines.append("")
            elif r < 0.03:
                lines.append("   ")
            elif r < 0.04:
                lines[-1] = "  \t" + lines[-1]
            elif r < 0.06:
                lines.append("\t".join([rng.choice(missing), "5", "10"] + ["x", "1", "+", "0", "0", "0"][: width - 3]))
            elif r < 0.07:
                nm, ln = seqs[0]
                lines.append("\t".join([nm, str(ln - 3), str(ln + 9)] + ["y", "2", "-", "0", "0", "0"][: width - 3]))
    return "\n".join(lines) + ("\n" if rng.random() < 0.7 else "")


STUB_SEQS = [("chrA", 400000000), ("chrB", 300000000), ("scaffold_17 x", 1000)]


@pytest.mark.parametrize("width,extras", [(3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (9, 0), (3, 2), (6, 3), (9, 1)])
def test_fast_equals_serial_on_synthetic_mapping(stub_cli, tmp_path, width, extras):
    rng = random.Random(width * 10 + extras)
    bed = tmp_path / "in.bed"
    bed.write_text(corpus(rng, STUB_SEQS, width, 20000, extras))
    ref_out = str(tmp_path / "serial.bed")
    args = ["--bedType", str(width)] if extras else []  # pass-through columns need the standard width spelled out
    rc0, err0, fast0 = run(stub_cli, "x", "S", str(bed), "T", ref_out, args=args, threads=0)
    assert rc0 == 0 and fast0 == 0, err0
    assert os.path.getsize(ref_out) > 100000
    for threads, block in ((1, None), (3, None), (8, None), (5, 100), (4, 4096), (7, 300000)):
        out = str(tmp_path / "fast.bed")
        rc, err, fast = run(stub_cli, "x", "S", str(bed), "T", out, args=args, threads=threads, block=block)
        assert rc == 0
        assert fast > 19000, "the multi-threaded path was not taken"
        assert open(out, "rb").read() == open(ref_out, "rb").read(), (threads, block)
        assert err == err0, (threads, block)


def test_fast_falls_back_block_by_block(stub_cli, tmp_path):
    """Mixed widths, tolerated-but-not-strict fields, CRLF, trailing TABs, sticky strand / thickEnd across widths: blocks
    the strict tokeniser rejects go through BedLine::parse, the others through the fast path, and the result does not
    depend on where the block boundaries fall."""
    rng = random.Random(5)
    parts = []
    for k in range(40):
        width = rng.choice([3, 4, 6, 7, 8, 9, 12])
        if width == 12:
            for i in range(rng.randint(1, 30)):
                a = rng.randint(0, 1000000)
                parts.append(f"chrA\t{a}\t{a + 100}\tb{i}\t0\t{rng.choice('+-')}\t{a}\t{a + 100}\t0\t2\t10,20\t0,80")
        else:
            parts.append(corpus(rng, STUB_SEQS, width, rng.randint(1, 400), strict=rng.random() < 0.5).rstrip("\n"))
        if rng.random() < 0.2:
            parts.append("chrB\t10\t20\r")
        if rng.random() < 0.2:
            parts.append("chrB\t10\t20\tnm\t")
    bed = tmp_path / "in.bed"
    bed.write_text("\n".join(parts) + "\n")
    ref_out = str(tmp_path / "serial.bed")
    rc0, err0, _ = run(stub_cli, "x", "S", str(bed), "T", ref_out, threads=0)
    assert rc0 == 0, err0
    took_fast = []
    for threads, block in ((4, None), (4, 64), (3, 700), (8, 5000), (2, 40000)):
        out = str(tmp_path / "fast.bed")
        rc, err, fast = run(stub_cli, "x", "S", str(bed), "T", out, threads=threads, block=block)
        assert rc == 0, err
        assert open(out, "rb").read() == open(ref_out, "rb").read(), (threads, block)
        assert err == err0
        took_fast.append(fast)
    assert took_fast[0] == 0 and max(took_fast) > 1000  # one big mixed block: serial; small blocks: mostly fast


def test_fast_reports_errors_like_serial(stub_cli, tmp_path):
    rng = random.Random(9)
    good = corpus(rng, STUB_SEQS, 6, 3000, noise=False)
    for bad, msg in (("chrA\t10\t5\tq\t0\t+", "Error zero or negative length BED range"), ("chrA\t10", "Expected at least three columns"),
                     ("chrA\tx\t5\tq\t0\t+", "Error converting string to int"), ("chrA\t1\t5\tq\t0\t*", "Strand character must be")):
        lines = good.splitlines()
        lines.insert(1234, bad)
        bed = tmp_path / "in.bed"
        bed.write_text("\n".join(lines) + "\n")
        res = [run(stub_cli, "x", "S", str(bed), "T", str(tmp_path / "o.bed"), threads=t, block=b) for t, b in ((0, None), (4, None), (4, 2000))]
        for rc, err, _ in res:
            assert rc == 1 and msg in err and "in input bed line 1235" in err
        assert res[0][1] == res[1][1] == res[2][1]


@pytest.mark.skipif(ref_bin("halLiftover") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("width,extras,args", [(3, 0, []), (6, 2, ["--bedType", "6"]), (9, 0, ["--noDupes"]), (8, 1, ["--bedType", "8"]), (6, 0, ["--bedType", "4"])])
def test_fast_path_matches_reference_binary(emul_cli, tmp_path, width, extras, args):
    """Real mapping (kernel sources on the warp emulator) + fast text layer == the reference CLI, byte for byte.
    (BED7 is left out: the reference's BedLine constructor does not initialise _thickEnd (halBedLine.cpp:19), so for
    exactly seven columns its cleanResults test reads an indeterminate value; both text paths here define it as 0.)"""
    import pyoracle
    hal = os.path.join(GOLDEN, "varlen8.hal")
    o = pyoracle.Oracle(hal)
    seqs = [(n, l) for (n, s, l) in o.sequences(o.genome_id("L1"))]
    o.close()
    rng = random.Random(width + 100 * extras)
    bed = tmp_path / "in.bed"
    bed.write_text(corpus(rng, seqs, width, 250, extras, missing=("nope",)))
    exp = str(tmp_path / "ref.bed")
    r = subprocess.run([ref_bin("halLiftover")] + args + [hal, "L1", str(bed), "L2", exp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for threads, block in ((4, None), (3, 500)):
        out = str(tmp_path / "o.bed")
        rc, err, fast = run(emul_cli, hal, "L1", str(bed), "L2", out, args=args, threads=threads, block=block)
        assert rc == 0, err
        assert fast > 200
        assert open(out, "rb").read() == open(exp, "rb").read(), (threads, block)
        assert err.strip() == r.stderr.strip()


@pytest.mark.gpu
@pytest.mark.skipif(ref_bin("halLiftover") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("width,extras,args", [(3, 0, []), (6, 1, ["--bedType", "6"]), (9, 0, ["--noDupes"])])
def test_fast_path_cuda_matches_reference_binary(tmp_path, width, extras, args):
    """The shipped CLI (CUDA library, multi-threaded text layer) against the reference CLI on 20k messy lines."""
    import pyoracle
    from hal_b200 import build
    build.build()
    cli = os.path.join(ROOT, "hal_b200", "bin", "halLiftover")
    hal = os.path.join(GOLDEN, "varlen8.hal")
    o = pyoracle.Oracle(hal)
    seqs = [(n, l) for (n, s, l) in o.sequences(o.genome_id("L3"))]
    o.close()
    rng = random.Random(width)
    bed = tmp_path / "in.bed"
    bed.write_text(corpus(rng, seqs, width, 20000, extras, missing=("nope",)))
    exp = str(tmp_path / "ref.bed")
    r = subprocess.run([ref_bin("halLiftover")] + args + [hal, "L3", str(bed), "L0", exp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for threads, block in ((8, None), (5, 100000), (0, None)):
        out = str(tmp_path / "o.bed")
        rc, err, fast = run(cli, hal, "L3", str(bed), "L0", out, args=args, threads=threads, block=block)
        assert rc == 0, err
        assert (fast > 19000) == (threads > 0)
        assert open(out, "rb").read() == open(exp, "rb").read(), (threads, block)
        assert err.strip() == r.stderr.strip()
