"""Dev tool: a halWiggleLiftover binary of this repo (emulated or CUDA) against oracle/_ref/halWiggleLiftover."""
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
    cli, hal = sys.argv[1], sys.argv[2]
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    o = pyoracle.Oracle(hal)
    ref = os.path.join(ROOT, "oracle", "_ref", "halWiggleLiftover")
    d = tempfile.mkdtemp()
    rng = random.Random(11)
    tot = bad = exc = turn = 0
    for _ in range(rounds):
        for src in o.genomes:
            for tgt in o.genomes:
                if src == tgt:
                    continue
                for nd in (False, True):
                    w = random_wig(rng, o.sequences(o.genome_id(src)), disorder=rng.choice([0, 0, 0.02]))
                    inp, out, mine = os.path.join(d, "w.wig"), os.path.join(d, "w.out"), os.path.join(d, "m.out")
                    open(inp, "w").write(w)
                    pre = None
                    if rng.random() < 0.4:
                        pre = random_wig(rng, o.sequences(o.genome_id(tgt)), sections=(1, 3), max_lines=60)
                        if "variableStep" in pre:
                            pre = None
                    for f in (out, mine):
                        if os.path.exists(f):
                            os.remove(f)
                        if pre is not None:
                            open(f, "w").write(pre)
                    flags = (["--noDupes"] if nd else []) + (["--append"] if pre is not None else [])
                    r = subprocess.run([ref] + flags + [hal, src, inp, tgt, out], capture_output=True, text=True)
                    m = subprocess.run([cli] + flags + [hal, src, inp, tgt, mine], capture_output=True, text=True)
                    tot += 1
                    if r.returncode != 0:
                        exc += 1
                        if "Could not find correct child" in r.stderr:
                            turn += 1  # the reference's wrong turn at the MRCA: this build maps correctly; the oracle can too
                            try:
                                exp = o.wiggle_liftover(src, tgt, w, no_dupes=nd, preload_text=pre, correct_path=True)
                            except RuntimeError as e:  # e.g. a later "Coordinate out of order"
                                if m.returncode != 1 or m.stderr.strip() != "hal exception caught: " + str(e):
                                    bad += 1
                                    print("ERRDIFF(correct path)", src, tgt, flags, str(e), repr(m.stderr.strip()))
                                continue
                            if m.returncode != 0 or open(mine).read() != exp:
                                bad += 1
                                print("DIFF(correct path)", src, tgt, flags, m.stderr.strip()[:200])
                            continue
                        if m.returncode == 0 or m.stderr.strip() != r.stderr.strip():
                            bad += 1
                            print("ERRDIFF", src, tgt, flags, repr(r.stderr.strip()), repr(m.stderr.strip()))
                    elif m.returncode != 0 or open(mine).read() != open(out).read():
                        bad += 1
                        print("DIFF", src, tgt, flags, m.stderr.strip()[:200])
                        open(os.path.join(d, f"bad{bad}.wig"), "w").write(w)
    print(tot, "cases,", exc, "reference exceptions (", turn, "wrong-turn: checked against the oracle's correct-path mode),", bad, "mismatches", d if bad else "")


if __name__ == "__main__":
    main()
