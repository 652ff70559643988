"""Dev tool: this repo's blockViz API (a blockVizCli linked against it) against the reference's (oracle/_ref/blockVizCli)."""
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402


def ancestors(o, g):
    out = []
    i = o.genome_id(g)
    while i >= 0:
        out.append(o.genomes[i])
        i = o.L.oracle_genome_parent(o.h, i)
    return out


def main():
    cli, hal = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    ref = os.path.join(ROOT, "oracle", "_ref", "blockVizCli")
    o = pyoracle.Oracle(hal)
    rng = random.Random(seed)
    bad = nonempty = dupes = 0
    for it in range(n):
        q, t = rng.choice(o.genomes), rng.choice(o.genomes)
        seqs = o.sequences(o.genome_id(t))
        nm, _, ln = rng.choice(seqs)
        L = rng.randint(1, min(ln, rng.choice([30, 400, 3000, 100000])))
        a = rng.randint(0, ln - L)
        b = a + L
        if rng.random() < 0.1:
            a, b = 0, 0  # whole chromosome
        dup = rng.choice([0, 1, 2])
        rev = 1 if (dup < 2 and rng.random() < 0.25) else 0
        seq = rng.choice([0, 0, 1, 2])
        lim = "-"
        if rng.random() < 0.3:
            aq, at = ancestors(o, q), ancestors(o, t)
            common = [x for x in aq if x in at]
            lim = rng.choice(common)
        adj = "1" if (rev == 0 and rng.random() < 0.5) else "0"
        args = ["blocks", q, t, nm, str(a), str(b), str(rev), str(seq), str(dup), adj, lim]
        if rng.random() < 0.15:
            args.append(rng.choice(o.sequences(o.genome_id(q)))[0])
        r = subprocess.run([ref, hal] + args, capture_output=True, text=True)
        m = subprocess.run([cli, hal] + args, capture_output=True, text=True)
        if r.returncode < 0:  # the reference crashed (assert / signal): nothing to compare with
            continue
        nonempty += r.stdout.startswith("B")
        dupes += "\nD\t" in r.stdout
        if r.stdout != m.stdout or r.returncode != m.returncode:
            bad += 1
            print("DIFF", " ".join(args), "| ref", len(r.stdout.splitlines()), "lines rc", r.returncode, "| mine", len(m.stdout.splitlines()), "rc", m.returncode,
                  (r.stdout[:150] + " // " + m.stdout[:150]) if "ERROR" in r.stdout + m.stdout else "")
    print(n, "queries,", nonempty, "with blocks,", dupes, "with target dupes,", bad, "mismatches")


if __name__ == "__main__":
    main()
