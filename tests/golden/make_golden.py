"""Regenerates tests/golden/ (run in the build container, where /root/reference and oracle/_ref exist).

  *.hal                     fixture alignments:
      randgenSmallSeed0.hal   `halRandGen --preset small --seed 0 --testRand --format mmap` (the file behind the
                              reference's CLI goldens, liftover/Makefile:32-63, maf/Makefile:38-54)
      refBedLiftoverTest.hal  the hand-built 5-genome alignment of liftover/tests/halLiftoverTests.cpp:15-252,
                              written by the reference's own test executable (ORACLE_KEEP_FIXTURE)
      varlen8.hal             oracle/gen/halTreeGen --mode varlen: 8 genomes, 4 sequences each, variable segment
                              lengths, inversions, duplications, insertions
  ref_liftover/             verbatim copies of the reference's golden INPUT/EXPECTED data files
                              (liftover/tests/input/*, liftover/tests/expected/*) for randgenSmallSeed0.hal
  cases/<name>.in.bed, .out.bed   random or hand-written BED inputs and the output of oracle/_ref/halLiftover
  cases/index.json          [{name, hal, src, tgt, args}]
"""
import json
import os
import random
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyoracle import Oracle  # noqa: E402


def run(*a, env=None):
    subprocess.check_call(list(a), env=env)


def main():
    os.makedirs(os.path.join(HERE, "cases"), exist_ok=True)
    small = os.path.join(HERE, "randgenSmallSeed0.hal")
    run(REF + "/halRandGen", "--preset", "small", "--seed", "0", "--testRand", "--format", "mmap", "--mmapFileSize", "1", small)
    tmp = "/tmp/golden_fix"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/halLiftoverTests"])
    run(REF + "/halLiftoverTests", env=dict(os.environ, ORACLE_KEEP_FIXTURE=tmp))
    shutil.copy(os.path.join(tmp, "15BedLiftoverTest.hal"), os.path.join(HERE, "refBedLiftoverTest.hal"))
    varlen = os.path.join(HERE, "varlen8.hal")
    run(REF + "/halTreeGen", "--mode", "varlen", "--newick", "((L0,L1)A0,(L2,(L3)A2)A1)R;", "--segs", "1500", "--minLen", "3",
        "--maxLen", "40", "--seqs", "4", "--seed", "11", "--pDup", "0.15", "--pInv", "0.25", varlen)
    for f in (small, os.path.join(HERE, "refBedLiftoverTest.hal"), varlen):
        run(REF + "/halValidate", f)

    cases = []

    def add(name, hal, src, tgt, bed, args=()):
        inp = os.path.join(HERE, "cases", name + ".in.bed")
        out = os.path.join(HERE, "cases", name + ".out.bed")
        open(inp, "w").write(bed)
        run(REF + "/halLiftover", *args, os.path.join(HERE, hal), src, inp, tgt, out)
        cases.append(dict(name=name, hal=hal, src=src, tgt=tgt, args=list(args)))

    # the BED inputs of BedLiftoverTest::testOneBranchLifts / testMultiBranchLifts (BED6 parts)
    t = "refBedLiftoverTest.hal"
    add("ref_child1_root", t, "child1", "root", "Sequence\t0\t20\tPARALOGY1REV\t0\t+\nSequence\t60\t80\tREV\t0\t+\n"
        "Sequence\t20\t40\tINSERTION\t0\t+\nSequence\t80\t100\tPARALOGY2\t0\t+\n")
    add("ref_leaf1_root", t, "leaf1", "root", "Sequence\t0\t5\tNORMALREV\t0\t+\nSequence\t10\t30\tOVERLAP\t0\t+\n"
        "Sequence\t50\t70\tOVERLAPINSERTION\t0\t+\nSequence\t70\t100\tOVERLAPINSERTION2\t0\t+\n")
    add("ref_root_child1", t, "root", "child1", "Sequence\t0\t10\tPARALOGY\t0\t+\nSequence\t30\t50\tOVERLAPINSERTION\t0\t+\n")
    add("ref_leaf2_leaf3", t, "leaf2", "leaf3", "Sequence\t30\t35\tREV\t0\t+\nSequence\t40\t60\tOVERLAP\t0\t+\n")
    add("ref_root_leaf2", t, "root", "leaf2", "Sequence\t0\t20\tBLOCK_A\t0\t+\nSequence\t30\t50\tBLOCK_B\t0\t+\n")
    # every genome pair of the hand-built fixture, every 7-base window, both strands
    o = Oracle(os.path.join(HERE, t))
    for s in o.genomes:
        for d in o.genomes:
            ln = o.genome_length(o.genome_id(s))
            bed = "".join(f"Sequence\t{a}\t{min(a + 7, ln)}\tw{a}\t0\t{'+-'[a % 2]}\n" for a in range(0, ln - 1, 3))
            add(f"ref_all_{s}_{d}", t, s, d, bed)
            add(f"ref_all_nodupes_{s}_{d}", t, s, d, bed, ("--noDupes",))

    def rand_bed(hal, src, n, maxlen, seed, strands="+-."):
        o = Oracle(os.path.join(HERE, hal))
        seqs = o.sequences(o.genome_id(src))
        rng = random.Random(seed)
        lines = []
        for i in range(n):
            nm, st, ln = rng.choice(seqs)
            L = rng.randint(1, min(maxlen, ln))
            a = rng.randint(0, ln - L)
            lines.append(f"{nm}\t{a}\t{a + L}\tn{i}\t{rng.randint(0, 999)}\t{rng.choice(strands)}")
        return "\n".join(lines) + "\n"

    v = "varlen8.hal"
    for i, (s, d, args) in enumerate([("L0", "L3", ()), ("L3", "L1", ()), ("L0", "L3", ("--noDupes",)), ("R", "L3", ()),
                                      ("R", "A2", ()), ("A0", "R", ()), ("A1", "L3", ()), ("A0", "L2", ()), ("A1", "A1", ()),
                                      ("L2", "A1", ()), ("L1", "L0", ())]):
        add(f"varlen_{s}_{d}{'_nodupes' if args else ''}", v, s, d, rand_bed(v, s, 400, 300, 100 + i), args)
    s = "randgenSmallSeed0.hal"
    for i, (a, b) in enumerate([("Genome_0", "Genome_2"), ("Genome_3", "Genome_2"), ("Genome_2", "Genome_3"), ("Genome_1", "Genome_0")]):
        add(f"small_{a}_{b}", s, a, b, rand_bed(s, a, 300, 900, 200 + i))
    # PSL (liftover/Makefile halLiftoverPsl*Test + BedLiftoverTest case3 PSL expectations, halLiftoverTests.cpp:358-373)
    def addp(name, hal, src, tgt, bed, args):
        inp = os.path.join(HERE, "cases", name + ".in.bed")
        out = os.path.join(HERE, "cases", name + ".out.bed")
        open(inp, "w").write(bed)
        run("timeout", "120", REF + "/halLiftover", *args, os.path.join(HERE, hal), src, inp, tgt, out)
        cases.append(dict(name=name, hal=hal, src=src, tgt=tgt, args=list(args)))
    case3 = ("Sequence\t0\t10\tSEGMENT_0\t0\t+\t0\t10\t128,0,0\t1\t10\t0,\n"
             "Sequence\t10\t30\tSEGMENT_1\t0\t+\t10\t30\t128,0,0\t1\t20\t0,\n"
             "Sequence\t30\t45\tSEGMENT_2\t0\t+\t30\t45\t128,0,0\t1\t15\t0,\n"
             "Sequence\t45\t65\tSEGMENT_3\t0\t+\t45\t65\t128,0,0\t1\t20\t0,\n"
             "Sequence\t65\t75\tSEGMENT_4\t0\t+\t65\t75\t128,0,0\t1\t10\t0,\n"
             "Sequence\t75\t100\tSEGMENT_5\t0\t+\t75\t100\t128,0,0\t1\t25\t0,\n")
    addp("psl_ref_leaf3_leaf1", t, "leaf3", "leaf1", case3, ("--outPSL",))
    addp("psl_ref_leaf3_leaf1_name", t, "leaf3", "leaf1", case3, ("--outPSLWithName",))
    # BED12: the unit test's case3 (halLiftoverTests.cpp:339-345) and random multi-block lines
    add("ref_bed12_leaf3_leaf1", t, "leaf3", "leaf1",
        "Sequence\t0\t10\tSEGMENT_0\t0\t+\t0\t10\t128,0,0\t1\t10\t0,\n"
        "Sequence\t10\t30\tSEGMENT_1\t0\t+\t10\t30\t128,0,0\t1\t20\t0,\n"
        "Sequence\t30\t45\tSEGMENT_2\t0\t+\t30\t45\t128,0,0\t1\t15\t0,\n"
        "Sequence\t45\t65\tSEGMENT_3\t0\t+\t45\t65\t128,0,0\t1\t20\t0,\n"
        "Sequence\t65\t75\tSEGMENT_4\t0\t+\t65\t75\t128,0,0\t1\t10\t0,\n"
        "Sequence\t75\t100\tSEGMENT_5\t0\t+\t75\t100\t128,0,0\t1\t25\t0,\n")

    def rand_bed12(hal, src, n, seed):
        o = Oracle(os.path.join(HERE, hal))
        seqs = o.sequences(o.genome_id(src))
        rng = random.Random(seed)
        lines = []
        for i in range(n):
            nm, st, ln = rng.choice(seqs)
            span = rng.randint(30, min(1500, ln))
            a = rng.randint(0, ln - span)
            nb = rng.randint(1, 6)
            cuts = sorted(rng.sample(range(1, span), min(2 * nb - 1, span - 1)))
            pts = [0] + cuts + [span]
            blocks = [(pts[k], pts[k + 1] - pts[k]) for k in range(0, len(pts) - 1, 2)]
            rng.shuffle(blocks) if rng.random() < 0.2 else None
            thick = (a, a + span) if rng.random() < 0.5 else (0, 0)
            lines.append("\t".join([nm, str(a), str(a + span), f"g{i}", str(rng.randint(0, 1000)), rng.choice("+-"), str(thick[0]),
                                    str(thick[1]), rng.choice(["0", "255,0,0", "1,2,3"]), str(len(blocks)),
                                    ",".join(str(b[1]) for b in blocks) + ",", ",".join(str(b[0]) for b in blocks) + ","]
                                   + (["extraA", "extraB"] if rng.random() < 0.3 else [])))
        return "\n".join(lines) + "\n"

    add("varlen_bed12_L0_L3", v, "L0", "L3", rand_bed12(v, "L0", 300, 301))
    add("varlen_bed12_R_L1", v, "R", "L1", rand_bed12(v, "R", 200, 302))
    add("varlen_bed12_L3_A0_nodupes", v, "L3", "A0", rand_bed12(v, "L3", 200, 303), ("--noDupes",))
    add("small_bed12", s, "Genome_0", "Genome_2", rand_bed12(s, "Genome_0", 200, 304))
    # halAlignmentDepth wiggles (no in-tree pin in the reference: alignmentDepth/Makefile:19) -> oracle/_ref outputs
    dcases = []
    def addd(name, hal, ref, args):
        out = os.path.join(HERE, "cases", name + ".wig")
        with open(out, "w") as f:
            subprocess.check_call([REF + "/halAlignmentDepth", os.path.join(HERE, hal), ref] + list(args), stdout=f)
        dcases.append(dict(name=name, hal=hal, ref=ref, args=list(args)))
    addd("depth_varlen_L0", v, "L0", [])
    addd("depth_varlen_L3_dupes", v, "L3", ["--countDupes"])
    addd("depth_varlen_L1_noanc", v, "L1", ["--noAncestors"])
    addd("depth_varlen_A1", v, "A1", [])
    addd("depth_varlen_R_dupes", v, "R", ["--countDupes"])
    addd("depth_varlen_L2_targets", v, "L2", ["--targetGenomes", "L0,A1"])
    addd("depth_varlen_L0_rootA0", v, "L0", ["--rootGenome", "A0"])
    addd("depth_varlen_L0_window", v, "L0", ["--start", "9000", "--length", "7000"])
    addd("depth_varlen_L0_seq", v, "L0", ["--refSequence", "L0_s2", "--start", "100", "--length", "999"])
    addd("depth_varlen_L0_step", v, "L0", ["--refSequence", "L0_s1", "--start", "10", "--length", "3003", "--step", "7"])
    addd("depth_small_G2", s, "Genome_2", [])
    addd("depth_ref_leaf3", t, "leaf3", ["--countDupes"])
    # hal2maf (maf/Makefile:38-54 style): windows keep the fixtures small; the tiny hand-built file is exported whole
    mcases = []
    def addm(name, hal, args):
        out = os.path.join(HERE, "cases", name + ".maf")
        if os.path.exists(out):
            os.remove(out)
        subprocess.check_call([REF + "/hal2maf", os.path.join(HERE, hal), out] + list(args))
        mcases.append(dict(name=name, hal=hal, args=list(args)))
    addm("maf_small_default", s, [])
    addm("maf_small_seqpart", s, ["--refGenome", "Genome_2", "--refSequence", "Genome_2_seq", "--start", "1000", "--length", "2000"])
    addm("maf_ref_root", t, [])
    addm("maf_ref_leaf3", t, ["--refGenome", "leaf3"])
    addm("maf_ref_child1_nodupes", t, ["--refGenome", "child1", "--noDupes"])
    for ref, sq in (("R", "R_s1"), ("L0", "L0_s2"), ("A1", "A1_s0"), ("L3", "L3_s3")):
        addm(f"maf_varlen_{ref}_win", v, ["--refGenome", ref, "--refSequence", sq, "--start", "200", "--length", "2500"])
    addm("maf_varlen_L0_nodupes", v, ["--refGenome", "L0", "--refSequence", "L0_s3", "--start", "0", "--length", "2500", "--noDupes"])
    addm("maf_varlen_L0_noanc", v, ["--refGenome", "L0", "--refSequence", "L0_s1", "--start", "100", "--length", "2500", "--noAncestors"])
    addm("maf_varlen_L1_orth", v, ["--refGenome", "L1", "--refSequence", "L1_s1", "--start", "100", "--length", "2500", "--onlyOrthologs"])
    addm("maf_varlen_L2_targets", v, ["--refGenome", "L2", "--refSequence", "L2_s0", "--length", "2500", "--targetGenomes", "L0,A1"])
    addm("maf_varlen_L2_root", v, ["--refGenome", "L2", "--refSequence", "L2_s2", "--length", "2500", "--rootGenome", "A1"])
    addm("maf_varlen_L3_names_len7", v, ["--refGenome", "L3", "--refSequence", "L3_s0", "--length", "1500", "--onlySequenceNames", "--maxBlockLen", "7"])
    addm("maf_varlen_A0_keepempty", v, ["--refGenome", "A0", "--refSequence", "A0_s2", "--length", "2500", "--keepEmptyRefBlocks"])
    # hal2maf --unique (ColumnIterator unique + isCanonicalOnRef): windows whose columns have reference-genome paralogs
    addm("maf_varlen_L0_unique_win", v, ["--refGenome", "L0", "--refSequence", "L0_s1", "--start", "11000", "--length", "4000", "--unique"])
    addm("maf_varlen_A0_unique_win", v, ["--refGenome", "A0", "--refSequence", "A0_s2", "--start", "300", "--length", "3000", "--unique"])
    addm("maf_varlen_L3_unique_len9", v, ["--refGenome", "L3", "--refSequence", "L3_s1", "--start", "100", "--length", "2500", "--unique", "--maxBlockLen", "9"])
    addm("maf_ref_child1_unique", t, ["--refGenome", "child1", "--unique"])
    addm("maf_small_G2_unique", s, ["--refGenome", "Genome_2", "--unique"])
    json.dump(mcases, open(os.path.join(HERE, "cases", "maf_index.json"), "w"), indent=1)
    json.dump(dcases, open(os.path.join(HERE, "cases", "depth_index.json"), "w"), indent=1)
    addp("psl_varlen_L0_L3", v, "L0", "L3", rand_bed(v, "L0", 300, 300, 401, "+-"), ("--outPSL",))
    addp("psl_varlen_L3_L1_name", v, "L3", "L1", rand_bed(v, "L3", 300, 300, 402, "+-"), ("--outPSLWithName",))
    addp("psl_varlen_R_L2_bed12", v, "R", "L2", rand_bed12(v, "R", 200, 403), ("--outPSL",))
    addp("psl_varlen_L1_A1_bed12_nodupes", v, "L1", "A1", rand_bed12(v, "L1", 200, 404), ("--outPSL", "--noDupes"))
    addp("psl_small_G3_G2", s, "Genome_3", "Genome_2", rand_bed(s, "Genome_3", 200, 900, 405, "+-"), ("--outPSL",))
    json.dump(cases, open(os.path.join(HERE, "cases", "index.json"), "w"), indent=1)
    print(len(cases), "cases written")


if __name__ == "__main__":
    main()
