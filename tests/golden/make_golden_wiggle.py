"""Regenerates tests/golden/wig/ (run in the build container, where oracle/_ref/halWiggleLiftover exists):
random wiggle inputs over the fixture alignments and what the REFERENCE's halWiggleLiftover does with them.

  wig/<name>.in.wig     input; wig/<name>.pre.wig (optional) the existing target file for --append
  wig/<name>.out.wig    the reference's output, or
  wig/<name>.err        its stderr when it exits 1
  wig/index.json        [{name, hal, src, tgt, args, status}]   status: "ok" | "error" | "wrong_turn"
"wrong_turn" = the reference throws "Could not find correct child ..." (its spanning-set bug, oracle/restate/wiggle.cpp);
for those the expected output of THIS build is the oracle's correct-path mode, stored as .out.wig by this script.
"""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "halWiggleLiftover")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pyoracle import Oracle  # noqa: E402
from wiggen import random_wig  # noqa: E402

PLAN = [  # (hal, src, tgt, noDupes, append, disorder)
    ("varlen8.hal", "L1", "L0", False, False, 0), ("varlen8.hal", "L1", "A0", False, True, 0), ("varlen8.hal", "L2", "L1", False, False, 0),
    ("varlen8.hal", "L3", "L0", True, False, 0), ("varlen8.hal", "R", "L3", False, False, 0), ("varlen8.hal", "R", "L2", True, True, 0),
    ("varlen8.hal", "A1", "L3", False, False, 0.02), ("varlen8.hal", "L3", "R", False, False, 0.02), ("varlen8.hal", "A2", "A0", False, False, 0),
    ("varlen8.hal", "L0", "L1", False, False, 0), ("varlen8.hal", "L1", "L3", False, False, 0), ("varlen8.hal", "A0", "L2", True, False, 0),
    ("varlen8.hal", "L0", "A1", False, True, 0),
    ("randgenSmallSeed0.hal", "Genome_0", "Genome_2", False, False, 0), ("randgenSmallSeed0.hal", "Genome_3", "Genome_0", False, False, 0),
    ("randgenSmallSeed0.hal", "Genome_2", "Genome_3", False, False, 0), ("randgenSmallSeed0.hal", "Genome_1", "Genome_3", True, True, 0.02),
    ("refBedLiftoverTest.hal", "child1", "root", False, False, 0), ("refBedLiftoverTest.hal", "leaf2", "leaf3", False, False, 0),
    ("refBedLiftoverTest.hal", "root", "leaf2", False, False, 0), ("refBedLiftoverTest.hal", "leaf3", "leaf2", False, False, 0),
    ("refBedLiftoverTest.hal", "leaf1", "child2", True, False, 0.02),
]


def main():
    d = os.path.join(HERE, "wig")
    os.makedirs(d, exist_ok=True)
    rng = random.Random(2024)
    index = []
    oracles = {}
    for k, (hal, src, tgt, nd, app, dis) in enumerate(PLAN):
        o = oracles.setdefault(hal, Oracle(os.path.join(HERE, hal)))
        name = f"w{k:02d}_{src}_{tgt}" + ("_nodupes" if nd else "") + ("_append" if app else "")
        w = random_wig(rng, o.sequences(o.genome_id(src)), sections=(2, 5), max_lines=250, disorder=dis)
        pre = None
        if app:
            while pre is None or "variableStep" in pre:
                pre = random_wig(rng, o.sequences(o.genome_id(tgt)), sections=(1, 3), max_lines=50)
        inp, out = os.path.join(d, name + ".in.wig"), os.path.join(d, name + ".out.wig")
        open(inp, "w").write(w)
        for f in (out, os.path.join(d, name + ".err"), os.path.join(d, name + ".pre.wig")):
            if os.path.exists(f):
                os.remove(f)
        if pre is not None:
            open(os.path.join(d, name + ".pre.wig"), "w").write(pre)
            open(out, "w").write(pre)
        args = (["--noDupes"] if nd else []) + (["--append"] if app else [])
        r = subprocess.run([REF] + args + [os.path.join(HERE, hal), src, inp, tgt, out], capture_output=True, text=True)
        status = "ok"
        if r.returncode != 0:
            if "Could not find correct child" in r.stderr:
                status = "wrong_turn"
                try:
                    open(out, "w").write(o.wiggle_liftover(src, tgt, w, no_dupes=nd, preload_text=pre, correct_path=True))
                except RuntimeError as e:
                    status = "error"
                    os.remove(out)
                    open(os.path.join(d, name + ".err"), "w").write("hal exception caught: " + str(e) + "\n")
            else:
                status = "error"
                if os.path.exists(out):
                    os.remove(out)
                open(os.path.join(d, name + ".err"), "w").write(r.stderr)
        index.append(dict(name=name, hal=hal, src=src, tgt=tgt, args=args, status=status, ref_stderr=r.stderr if status == "wrong_turn" else ""))
        print(name, status, os.path.getsize(inp))
    json.dump(index, open(os.path.join(d, "index.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
