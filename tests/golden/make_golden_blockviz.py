"""Regenerates tests/golden/blockviz/cases.json (run where oracle/_ref/blockVizCli exists): queries against the REFERENCE's
blockViz C API (blockViz/impl/halBlockViz.cpp compiled from /root/reference behind tests/cpp/blockviz_cli.cpp) and its answers."""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "blockVizCli")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyoracle import Oracle  # noqa: E402


def ancestors(o, g):
    out, i = [], o.genome_id(g)
    while i >= 0:
        out.append(o.genomes[i])
        i = o.L.oracle_genome_parent(o.h, i)
    return out


def main():
    rng = random.Random(31)
    cases = []
    for hal, n in (("varlen8.hal", 40), ("refBedLiftoverTest.hal", 12), ("randgenSmallSeed0.hal", 10), ("refMapExtraParalogsTest.hal", 6)):
        o = Oracle(os.path.join(HERE, hal))
        fixed = [["species"], ["maxlod"], ["chroms", o.genomes[-1]], ["limits", o.genomes[-1], o.genomes[1]]]
        nm, _, ln = o.sequences(o.genome_id(o.genomes[-1]))[0]
        fixed.append(["dna", o.genomes[-1], nm, "3", str(min(ln, 70))])
        fixed.append(["blocks", "nope", o.genomes[0], nm, "0", "10", "0", "0", "0", "0", "-"])
        fixed.append(["blocks", o.genomes[0], o.genomes[-1], nm, "9", "3", "0", "0", "0", "0", "-"])
        fixed.append(["blocks", o.genomes[0], o.genomes[-1], nm, "0", "10", "1", "0", "2", "0", "-"])
        t0 = o.genomes[-1]
        others = ",".join(g for g in o.genomes[:3] if g != t0)
        fixed.append(["maf", t0, nm, "0", str(min(ln, 600)), "0", "1000", "1", others])
        fixed.append(["maf", t0, nm, "5", str(min(ln, 300)), "0", "9", "0", o.genomes[0]])
        if len(o.sequences(o.genome_id(t0))) > 1:  # a later sequence: the reference passes the ABSOLUTE start to convertSequence
            nm2, _, ln2 = o.sequences(o.genome_id(t0))[1]
            fixed.append(["maf", t0, nm2, "10", str(min(ln2, 200)), "0", "1000", "1", o.genomes[0]])
        if hal == "varlen8.hal":  # whole chromosomes with target duplications (hal_target_dupe_list_t), +- adjacencies
            fixed.append(["blocks", "L1", "L2", "L2_s1", "0", "0", "0", "0", "2", "0", "-"])
            fixed.append(["blocks", "L1", "L0", "L0_s2", "0", "0", "0", "0", "2", "1", "-"])
            fixed.append(["blocks", "L1", "L3", "L3_s3", "200", "3000", "0", "0", "2", "1", "-"])
        queries = fixed
        for _ in range(n):
            q, t = rng.choice(o.genomes), rng.choice(o.genomes)
            nm, _, ln = rng.choice(o.sequences(o.genome_id(t)))
            L = rng.randint(1, min(ln, rng.choice([30, 400, 2500])))
            a = rng.randint(0, ln - L)
            dup = rng.choice([0, 1, 2])
            rev = 1 if (dup < 2 and rng.random() < 0.25) else 0
            lim = "-"
            if rng.random() < 0.3:
                common = [x for x in ancestors(o, q) if x in ancestors(o, t)]
                lim = rng.choice(common)
            adj = "1" if (rev == 0 and rng.random() < 0.4) else "0"
            args = ["blocks", q, t, nm, str(a), str(a + L), str(rev), str(rng.choice([0, 0, 2])), str(dup), adj, lim]
            if rng.random() < 0.15:
                args.append(rng.choice(o.sequences(o.genome_id(q)))[0])
            queries.append(args)
        for args in queries:
            r = subprocess.run([REF, os.path.join(HERE, hal)] + args, capture_output=True, text=True)
            if r.returncode < 0:
                continue
            cases.append(dict(hal=hal, args=args, rc=r.returncode, out=r.stdout))
    # level-of-detail list (tests/golden/varlen8.lod.txt: 0 varlen8.hal / 500 varlen8_lod1.hal / 20000 max; varlen8_lod1.hal =
    # halTreeGen --mode varlen, same tree and names, --segs 400 --minLen 20 --maxLen 120 --seqs 4 --seed 12 --pDup 0.1 --pInv 0.2
    # --meta 'L3:assembly=lod1 test' --meta L3:zebra=1 --meta 'R:note=root genome'):
    # which file answers depends on the query length, DNA always comes from level 0, lengths >= 20000 are refused
    a, b = Oracle(os.path.join(HERE, "varlen8.hal")), Oracle(os.path.join(HERE, "varlen8_lod1.hal"))
    lod = "varlen8.lod.txt"
    for args in (["meta", "L3"], ["meta", "R"], ["meta", "L0"], ["meta", "nope"]):  # --meta of halTreeGen: L3:assembly, L3:zebra, R:note
        r = subprocess.run([REF, os.path.join(HERE, "varlen8_lod1.hal")] + args, capture_output=True, text=True)
        cases.append(dict(hal="varlen8_lod1.hal", args=args, rc=r.returncode, out=r.stdout))
    lodq = [["species"], ["maxlod"], ["dna", "L3", "L3_s0", "0", "50"], ["chroms", "L2"]]
    for _ in range(24):
        q, t = rng.choice(a.genomes), rng.choice(a.genomes)
        si = rng.randrange(4)
        nm, _, la = a.sequences(a.genome_id(t))[si]
        lb = b.sequences(b.genome_id(t))[si][2]
        ln = min(la, lb)
        L = rng.randint(1, min(ln, rng.choice([100, 499, 500, 501, 3000, 19999, 20000, 25000])))
        st = rng.randint(0, ln - L)
        lodq.append(["blocks", q, t, nm, str(st), str(st + L), "0", str(rng.choice([0, 1, 2])), str(rng.choice([0, 1, 2])), rng.choice(["0", "1"]), "-"])
    for args in lodq:
        r = subprocess.run([REF, os.path.join(HERE, lod)] + args, capture_output=True, text=True)
        if r.returncode >= 0:
            cases.append(dict(hal=lod, args=args, rc=r.returncode, out=r.stdout))
    os.makedirs(os.path.join(HERE, "blockviz"), exist_ok=True)
    json.dump(cases, open(os.path.join(HERE, "blockviz", "cases.json"), "w"), indent=0)
    print(len(cases), "cases,", sum(1 for c in cases if "\nD\t" in c["out"]), "with target dupes,", sum(1 for c in cases if c["rc"] != 0), "errors")


if __name__ == "__main__":
    main()
