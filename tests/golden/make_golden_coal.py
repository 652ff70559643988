"""Adds the halLiftover --coalescenceLimit cases to tests/golden/cases (run where oracle/_ref/halLiftover exists):
random BED6 / BED12 inputs on varlen8.hal and the REFERENCE's output with a coalescence limit above the MRCA
(mapRecursiveParalogies, api/impl/halSegmentMapper.cpp:525-576).  Idempotent: replaces earlier coal_* entries of index.json."""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "halLiftover")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyoracle import Oracle  # noqa: E402

PLAN = [("L3", "L2", "R", 6), ("L3", "A2", "R", 6), ("L2", "L3", "R", 6), ("A2", "L3", "A1", 6), ("L1", "L1", "R", 6), ("L0", "L1", "R", 3),
        ("L3", "A2", "A1", 12), ("A1", "L2", "R", 12)]


def main():
    hal = "varlen8.hal"
    o = Oracle(os.path.join(HERE, hal))
    idx_path = os.path.join(HERE, "cases", "index.json")
    cases = [c for c in json.load(open(idx_path)) if not c["name"].startswith("coal_")]
    rng = random.Random(77)
    for src, tgt, lim, width in PLAN:
        seqs = o.sequences(o.genome_id(src))
        lines = []
        for i in range(160):
            nm, _, ln = rng.choice(seqs)
            L = rng.randint(1, min(300, ln))
            a = rng.randint(0, ln - L)
            st = rng.choice("+-.") if width != 12 else rng.choice("+-")
            if width == 3:
                lines.append(f"{nm}\t{a}\t{a + L}")
            elif width == 6:
                lines.append(f"{nm}\t{a}\t{a + L}\tn{i}\t{i}\t{st}")
            else:
                nb = rng.randint(1, 4)
                cuts = sorted(rng.sample(range(1, L), min(2 * nb - 1, L - 1))) if L > 2 else []
                bounds = [0] + cuts + [L]
                blocks = [(bounds[j], bounds[j + 1] - bounds[j]) for j in range(0, len(bounds) - 1, 2)]
                lines.append(f"{nm}\t{a}\t{a + L}\tn{i}\t0\t{st}\t{a}\t{a + L}\t0\t{len(blocks)}\t" + ",".join(str(b[1]) for b in blocks) + "\t" +
                             ",".join(str(b[0]) for b in blocks))
        name = f"coal_{'bed12_' if width == 12 else ''}{src}_{tgt}_{lim}"
        inp = os.path.join(HERE, "cases", name + ".in.bed")
        out = os.path.join(HERE, "cases", name + ".out.bed")
        open(inp, "w").write("\n".join(lines) + "\n")
        args = ["--coalescenceLimit", lim]
        subprocess.check_call([REF] + args + [os.path.join(HERE, hal), src, inp, tgt, out])
        plain = subprocess.run([REF, os.path.join(HERE, hal), src, inp, tgt, "stdout"], capture_output=True, text=True).stdout
        assert plain != open(out).read(), name + ": the limit changes nothing on this input"
        cases.append(dict(name=name, hal=hal, src=src, tgt=tgt, args=args))
        print(name, len(open(out).read().splitlines()), "lines (default:", len(plain.splitlines()), ")")
    json.dump(cases, open(idx_path, "w"), indent=1)


if __name__ == "__main__":
    main()
