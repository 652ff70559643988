"""Dev tool: oracle/restate/wiggle.cpp against oracle/_ref/halWiggleLiftover on random wiggle files, every genome pair."""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle  # noqa: E402
from wiggen import random_wig  # noqa: E402


def main():
    hal = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "varlen8.hal")
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    o = pyoracle.Oracle(hal)
    ref = os.path.join(ROOT, "oracle", "_ref", "halWiggleLiftover")
    d = tempfile.mkdtemp()
    rng = random.Random(3)
    tot = bad = exc = 0
    for _ in range(rounds):
        for src in o.genomes:
            for tgt in o.genomes:
                if src == tgt:
                    continue
                for nd in (False, True):
                    w = random_wig(rng, o.sequences(o.genome_id(src)), disorder=rng.choice([0, 0, 0.02]))
                    inp, out = os.path.join(d, "w.wig"), os.path.join(d, "w.out")
                    open(inp, "w").write(w)
                    pre = None
                    if rng.random() < 0.4:  # --append onto an existing target wiggle
                        pre = random_wig(rng, o.sequences(o.genome_id(tgt)), sections=(1, 3), max_lines=60)
                        if "variableStep" in pre:  # (the reference reads those 0-based and writes 1-based: keep it simple)
                            pre = None
                    if os.path.exists(out):
                        os.remove(out)
                    if pre is not None:
                        open(out, "w").write(pre)
                    r = subprocess.run([ref] + (["--noDupes"] if nd else []) + (["--append"] if pre is not None else []) +
                                       [hal, src, inp, tgt, out], capture_output=True, text=True)
                    tot += 1
                    try:
                        got, err = o.wiggle_liftover(src, tgt, w, no_dupes=nd, preload_text=pre), None
                    except RuntimeError as e:
                        got, err = None, str(e)
                    if r.returncode != 0:
                        exc += 1
                        msg = r.stderr.strip().replace("hal exception caught: ", "")
                        if err != msg:
                            bad += 1
                            print("ERRDIFF", src, tgt, nd, repr(msg), repr(err))
                    elif got != open(out).read():
                        bad += 1
                        print("DIFF", src, tgt, nd, err)
                        open(os.path.join(d, f"bad{bad}.wig"), "w").write(w)
    print(tot, "cases,", exc, "reference exceptions,", bad, "mismatches", d if bad else "")


if __name__ == "__main__":
    main()
