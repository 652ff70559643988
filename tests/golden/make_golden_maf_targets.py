"""Regenerates tests/golden/maf_targets/: hal2maf --refTargets goldens from the reference binary (oracle/_ref/hal2maf).
Each case: <name>.bed (random BED3..BED12 targets incl. unknown sequences, out-of-range ends, blank lines), <name>.maf and the
reference's stderr in index.json.  usage: python tests/golden/make_golden_maf_targets.py"""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pyoracle  # noqa: E402
from maf_targets_vs_ref import random_targets  # noqa: E402


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "hal2maf")
    out_dir = os.path.join(HERE, "maf_targets")
    os.makedirs(out_dir, exist_ok=True)
    rng = random.Random(2024)
    cases = []
    plan = [("varlen8.hal", "L0", []), ("varlen8.hal", "A1", ["--noDupes"]), ("varlen8.hal", "L3", ["--unique"]),
            ("varlen8.hal", "R", ["--maxBlockLen", "5", "--onlySequenceNames"]), ("refBedLiftoverTest.hal", "leaf3", []),
            ("refBedLiftoverTest.hal", "child1", ["--unique"]), ("randgenSmallSeed0.hal", "Genome_2", ["--noAncestors"])]
    for k, (hal, genome, extra) in enumerate(plan):
        o = pyoracle.Oracle(os.path.join(HERE, hal))
        seqs = o.sequences(o.genome_id(genome))
        o.close()
        name = f"targets_{k}_{genome}"
        bed, maf = os.path.join(out_dir, name + ".bed"), os.path.join(out_dir, name + ".maf")
        open(bed, "w").write(random_targets(rng, seqs, 14))
        if os.path.exists(maf):
            os.remove(maf)
        r = subprocess.run([ref, os.path.join(HERE, hal), maf, "--refGenome", genome, "--refTargets", bed] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        cases.append(dict(name=name, hal=hal, args=["--refGenome", genome] + extra, stderr=r.stderr))
    # a malformed line stops the scan with the scanner's message and line number; what was written before it stays
    name = "targets_bad_line"
    bed, maf = os.path.join(out_dir, name + ".bed"), os.path.join(out_dir, name + ".maf")
    open(bed, "w").write("L0_s0\t5\t40\n\nL0_s1\t7\t7\n")
    if os.path.exists(maf):
        os.remove(maf)
    r = subprocess.run([ref, os.path.join(HERE, "varlen8.hal"), maf, "--refGenome", "L0", "--refTargets", bed], capture_output=True, text=True)
    assert r.returncode == 1
    cases.append(dict(name=name, hal="varlen8.hal", args=["--refGenome", "L0"], stderr=r.stderr, returncode=1))
    json.dump(cases, open(os.path.join(out_dir, "index.json"), "w"), indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
