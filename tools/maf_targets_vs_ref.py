"""Dev tool: hal2maf --refTargets of this repo (emulated or CUDA binary) against oracle/_ref/hal2maf on random BED targets
(BED3..BED12, invalid coordinates, unknown sequences, blank lines).  usage: maf_targets_vs_ref.py <hal2maf binary> <hal> [rounds]"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402


def random_targets(rng, seqs, n, messy=True):
    lines = []
    for i in range(n):
        name, _, ln = rng.choice(seqs)
        if ln < 3:
            continue
        a = rng.randrange(ln - 1)
        b = min(ln, a + 1 + rng.randrange(1, 60))
        kind = rng.random()
        if messy and kind < 0.06:
            name = "nosuchseq"
        elif messy and kind < 0.12:
            b = ln + rng.randrange(1, 5)  # beyond the end
        w = rng.choice([3, 3, 4, 6, 9, 12])
        row = [name, str(a), str(b)]
        if w >= 4:
            row.append(f"t{i}")
        if w >= 6:
            row += [str(rng.randrange(1000)), rng.choice("+-.")]
        if w >= 9:
            row += [str(a), str(b), "0,0,255"]
        if w >= 12:
            nb = rng.randrange(1, 4)
            span = b - a
            cuts = sorted(rng.sample(range(0, span + 1), min(span + 1, 2 * nb)))
            if len(cuts) % 2:
                cuts = cuts[:-1]
            starts = cuts[0::2] or [0]
            sizes = [max(0, e - s) for s, e in zip(cuts[0::2], cuts[1::2])] or [span]
            row += [str(len(starts)), ",".join(map(str, sizes)) + ",", ",".join(map(str, starts)) + ","]
        lines.append("\t".join(row))
        if messy and rng.random() < 0.1:
            lines.append("")
    return "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")


def main():
    cli, hal = sys.argv[1], sys.argv[2]
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ref = os.path.join(ROOT, "oracle", "_ref", "hal2maf")
    o = pyoracle.Oracle(hal)
    rng = random.Random(5)
    d = tempfile.mkdtemp()
    tot = bad = nonempty = errs = 0
    for _ in range(rounds):
        for g in o.genomes:
            seqs = o.sequences(o.genome_id(g))
            for extra in ([], ["--noDupes"], ["--unique"], ["--maxBlockLen", "5", "--onlySequenceNames"], ["--append"]):
                bed = os.path.join(d, "t.bed")
                open(bed, "w").write(random_targets(rng, seqs, rng.randrange(1, 25)))
                a, b = os.path.join(d, "a.maf"), os.path.join(d, "b.maf")
                res = []
                for binary, out in ((ref, a), (cli, b)):
                    if os.path.exists(out):
                        os.remove(out)
                    r = subprocess.run([binary, hal, out, "--refGenome", g, "--refTargets", bed] + extra, capture_output=True, text=True)
                    err = "\n".join(x for x in r.stderr.splitlines() if not x.startswith("[halgpu"))
                    res.append((r.returncode, open(out).read() if os.path.exists(out) else None, err))
                tot += 1
                nonempty += bool(res[0][1])
                errs += bool(res[0][2])
                if res[0][0] < 0:
                    continue  # the reference died on one of its own asserts
                if res[0] != res[1]:
                    bad += 1
                    print("DIFF", g, extra, res[0][0], res[1][0], repr(res[0][2][:300]), repr(res[1][2][:300]))
                    os.system(f"cp {bed} /tmp/bad_targets.bed")
    print(tot, "cases,", nonempty, "with output,", errs, "with messages,", bad, "mismatches")


if __name__ == "__main__":
    main()
