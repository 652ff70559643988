"""Dev tool: a halSynteny binary of this repo (emulated or CUDA) against oracle/_ref/halSynteny on every genome pair."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402


def main():
    cli, hal = sys.argv[1], sys.argv[2]
    o = pyoracle.Oracle(hal)
    ref = os.path.join(ROOT, "oracle", "_ref", "halSynteny")
    d = tempfile.mkdtemp()
    tot = bad = nonempty = 0
    for src in o.genomes:
        for tgt in o.genomes:
            if src == tgt:
                continue
            for extra in ([], ["--minBlockSize", "50", "--maxAnchorDistance", "200"], ["--minBlockSize", "1", "--maxAnchorDistance", "5"],
                          ["--minBlockSize", "300", "--maxAnchorDistance", "100000", "--queryChromosome", o.sequences(o.genome_id(src))[-1][0]]):
                a, b = os.path.join(d, "a.psl"), os.path.join(d, "b.psl")
                r = subprocess.run([ref, "--queryGenome", src, "--targetGenome", tgt] + extra + [hal, a], capture_output=True, text=True)
                m = subprocess.run([cli, "--queryGenome", src, "--targetGenome", tgt] + extra + [hal, b], capture_output=True, text=True)
                tot += 1
                if r.returncode != 0 or m.returncode != 0:
                    if r.returncode != m.returncode:
                        bad += 1
                        print("RC DIFF", src, tgt, extra, r.returncode, m.returncode, r.stderr[:100], m.stderr[:100])
                    continue
                ea, eb = open(a).read(), open(b).read()
                nonempty += bool(ea)
                if ea != eb:
                    bad += 1
                    print("DIFF", src, tgt, extra, len(ea.splitlines()), len(eb.splitlines()))
    print(tot, "cases,", nonempty, "non-empty,", bad, "mismatches")


if __name__ == "__main__":
    main()
