"""Randomised parity hunt on the CPU: random synthetic alignments (hal_b200/bin/halSynth: random trees, segment lengths,
transposition / inversion / insertion rates) through the kernel sources on the warp emulator, against the oracle.
Covers liftover (dupes / noDupes / strands / coalescence limits), ColumnLiftover mode, alignment depth, wiggle liftover and
hal2maf (+ --unique) text.  usage: python tools/stress_emul.py [rounds] [seed]"""
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hal_b200  # noqa: E402
import pyoracle  # noqa: E402
from wiggen import random_wig  # noqa: E402

EMUL = os.path.join(ROOT, "tests", "simt", "libhalgpu_emul.so")
MAF = os.path.join(ROOT, "tests", "simt", "hal2maf_emul")
WIG = os.path.join(ROOT, "tests", "simt", "halWiggleLiftover_emul")
SYN = os.path.join(ROOT, "tests", "simt", "halSynteny_emul")
REF_SYN = os.path.join(ROOT, "oracle", "_ref", "halSynteny")
VIZ = os.path.join(ROOT, "tests", "simt", "blockVizCli_emul")
REF_VIZ = os.path.join(ROOT, "oracle", "_ref", "blockVizCli")


def rand_tree(rng, names):
    def rec(depth):
        name = f"G{len(names)}"
        names.append(name)
        if depth == 0 or rng.random() < 0.25:
            return name
        kids = [rec(depth - 1) for _ in range(rng.choice([1, 2, 2, 3]))]
        return "(" + ",".join(kids) + ")" + name
    return rec(rng.randint(1, 3)) + ";"


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="stress_")
    bad = 0
    for it in range(rounds):
        names = []
        newick = rand_tree(rng, names)
        if len(names) < 2:
            continue
        hal = os.path.join(d, f"s{it}.hal")
        segs, seglen, branch = rng.choice([300, 800, 2000]), rng.choice([3, 8, 24]), rng.choice([0, 0.05, 0.15, 0.4])
        subprocess.check_call([os.path.join(ROOT, "hal_b200", "bin", "halSynth"), "--newick", newick, "--segs", str(segs), "--segLen", str(seglen),
                               "--branch", str(branch), "--seed", str(rng.randint(1, 10 ** 6)), hal])
        o = pyoracle.Oracle(hal)
        a = hal_b200.Alignment(hal, lib_path=EMUL)
        tag = f"[{it}] {newick} segs={segs}x{seglen} branch={branch}"
        for _ in range(4):
            src, tgt = rng.choice(names), rng.choice(names)
            s, t = a.genome_id(src), a.genome_id(tgt)
            glen = a.genome_length(s)
            n = rng.choice([40, 150])
            ln = np.array([rng.randint(1, min(glen, rng.choice([5, 60, 400, 3000]))) for _ in range(n)])
            gs = np.array([rng.randint(0, glen - int(l)) for l in ln], np.int64)
            ge = gs + ln - 1
            st = np.array([ord(rng.choice("+-.")) for _ in range(n)], np.uint8)
            nd = rng.random() < 0.3
            # coalescence limit: a random ancestor-or-self of the MRCA
            anc = []
            g = a.L.halgpu_mrca(a.h, s, t)
            while g >= 0:
                anc.append(g)
                g = a.L.halgpu_genome_parent(a.h, g)
            lim = rng.choice(anc) if rng.random() < 0.5 else -1
            off, recs, _ = a.liftover(s, t, gs, ge, st, 1 if nd else 0, coalescence_limit=lim)
            e = o.liftover(s, t, gs, ge, st, no_dupes=nd, coalescence_limit=None if lim < 0 else lim)
            ok = np.array_equal(off, e["offsets"]) and all(np.array_equal(recs[k], e[k2]) for k, k2 in (
                ("start", "start"), ("end", "end"), ("src_start", "srcStart"), ("tgt_seq", "tgtSeq"), ("strand", "strand"), ("src_strand", "srcStrand")))
            if not ok:
                bad += 1
                print("LIFTOVER DIFF", tag, src, tgt, "noDupes" if nd else "", "lim", lim)
            # depth
            last = min(glen, 3000) - 1
            fl = rng.choice([0, 1, 4])
            dd, _ = a.depth(s, 0, last, 1, (), fl)
            ee, _ = o.depth(s, 0, last, 1, (), count_dupes=bool(fl & 1), no_dupes=bool(fl & 4))
            if not np.array_equal(dd, ee):
                bad += 1
                print("DEPTH DIFF", tag, src, fl)
        # hal2maf text, +- unique, on one random reference
        ref = rng.choice(names)
        for uniq in (False, True):
            out = os.path.join(d, "o.maf")
            r = subprocess.run([MAF, hal, out, "--refGenome", ref] + (["--unique"] if uniq else []), capture_output=True, text=True)
            try:
                exp = o.hal2maf(ref, unique=uniq)
            except RuntimeError:
                exp = None
            if r.returncode != 0 or exp is None or open(out, "rb").read() != exp:
                bad += 1
                print("MAF DIFF", tag, ref, "unique" if uniq else "", r.stderr[:200])
        # wiggle text (the oracle's correct-path mode is this build's semantics everywhere)
        src, tgt = rng.choice(names), rng.choice(names)
        if src != tgt:
            w = random_wig(rng, o.sequences(o.genome_id(src)), max_lines=150, disorder=rng.choice([0, 0.02]))
            inp, out = os.path.join(d, "i.wig"), os.path.join(d, "o.wig")
            open(inp, "w").write(w)
            nd = rng.random() < 0.3
            r = subprocess.run([WIG] + (["--noDupes"] if nd else []) + [hal, src, inp, tgt, out], capture_output=True, text=True)
            try:
                exp, err = o.wiggle_liftover(src, tgt, w, no_dupes=nd, correct_path=True), None
            except RuntimeError as ex:
                exp, err = None, "hal exception caught: " + str(ex)
            if (exp is not None and (r.returncode != 0 or open(out).read() != exp)) or (exp is None and r.stderr.strip() != err):
                bad += 1
                print("WIGGLE DIFF", tag, src, tgt, nd, r.stderr[:200], err)
        # halSynteny against the reference binary (there is no separate restatement of dag_merge in the oracle)
        if os.path.exists(REF_SYN) and os.path.exists(SYN):
            src, tgt = rng.choice(names), rng.choice(names)
            if src != tgt:
                args = ["--queryGenome", src, "--targetGenome", tgt, "--minBlockSize", str(rng.choice([1, 30, 500])), "--maxAnchorDistance", str(rng.choice([1, 50, 5000]))]
                pa, pb = os.path.join(d, "a.psl"), os.path.join(d, "b.psl")
                ra = subprocess.run([REF_SYN] + args + [hal, pa], capture_output=True, text=True)
                rb = subprocess.run([SYN] + args + [hal, pb], capture_output=True, text=True)
                if ra.returncode != rb.returncode or (ra.returncode == 0 and open(pa).read() != open(pb).read()):
                    bad += 1
                    print("SYNTENY DIFF", tag, args, ra.stderr[:100], rb.stderr[:100])
        # the blockViz C API against the reference's implementation behind the same text driver
        if os.path.exists(REF_VIZ) and os.path.exists(VIZ):
            for _ in range(6):
                q, t = rng.choice(names), rng.choice(names)
                nm, _, ln = rng.choice(o.sequences(o.genome_id(t)))
                L = rng.randint(1, min(ln, rng.choice([40, 600, 20000])))
                st0 = rng.randint(0, ln - L)
                dup = rng.choice([0, 1, 2])
                rev = "1" if (dup < 2 and rng.random() < 0.25) else "0"
                args = ["blocks", q, t, nm, str(st0), str(st0 + L), rev, str(rng.choice([0, 2])), str(dup), "1" if (rev == "0" and rng.random() < 0.5) else "0", "-"]
                if rng.random() < 0.25:
                    args = ["maf", t, nm, str(st0), str(st0 + L), "0", str(rng.choice([1000, 7])), str(rng.choice([0, 1])), ",".join(rng.sample(names, min(2, len(names))))]
                ra = subprocess.run([REF_VIZ, hal] + args, capture_output=True, text=True)
                rb = subprocess.run([VIZ, hal] + args, capture_output=True, text=True)
                if ra.returncode >= 0 and (ra.returncode, ra.stdout) != (rb.returncode, rb.stdout):
                    bad += 1
                    print("BLOCKVIZ DIFF", tag, " ".join(args), ra.stdout[:120].replace("\n", " | "), "//", rb.stdout[:120].replace("\n", " | "))
        a.close()
        o.close()
        print("round", it, "done", tag, flush=True)
    print("mismatches:", bad)


if __name__ == "__main__":
    main()
