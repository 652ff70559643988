"""Regenerates tests/golden/synteny/ (run where oracle/_ref/halSynteny exists): the reference's halSynteny output for a set of
(alignment, query, target, options) cases.  tests/golden/ref_synteny/test1.psl is the reference repo's OWN golden
(synteny/tests/expected/test1.psl) for `halRandGen --seed 0 --testRand` (tests/golden/randgenDefaultSeed0.hal, mmap format)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "halSynteny")

CASES = [
    ("varlen8.hal", "L0", "L3", ["--minBlockSize", "50", "--maxAnchorDistance", "200"]),
    ("varlen8.hal", "L3", "L1", ["--minBlockSize", "1", "--maxAnchorDistance", "5"]),
    ("varlen8.hal", "R", "L2", ["--minBlockSize", "100", "--maxAnchorDistance", "1000"]),
    ("varlen8.hal", "A1", "L0", ["--minBlockSize", "20", "--maxAnchorDistance", "60", "--queryChromosome", "A1_s2"]),
    ("varlen8.hal", "L2", "A2", ["--minBlockSize", "300", "--maxAnchorDistance", "100000"]),
    ("randgenSmallSeed0.hal", "Genome_3", "Genome_2", ["--minBlockSize", "500", "--maxAnchorDistance", "3000"]),
    ("randgenSmallSeed0.hal", "Genome_0", "Genome_3", []),
    ("refBedLiftoverTest.hal", "leaf2", "leaf3", ["--minBlockSize", "1", "--maxAnchorDistance", "50"]),
    ("randgenDefaultSeed0.hal", "Genome_3", "Genome_7", []),
]


def main():
    d = os.path.join(HERE, "synteny")
    os.makedirs(d, exist_ok=True)
    index = []
    for k, (hal, q, t, args) in enumerate(CASES):
        name = f"syn{k:02d}_{q}_{t}"
        out = os.path.join(d, name + ".psl")
        subprocess.check_call([REF, "--queryGenome", q, "--targetGenome", t] + args + [os.path.join(HERE, hal), out])
        index.append(dict(name=name, hal=hal, query=q, target=t, args=args))
        print(name, sum(1 for _ in open(out)), "lines")
    json.dump(index, open(os.path.join(d, "index.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
