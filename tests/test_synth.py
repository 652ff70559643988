"""halSynth (hal_b200/csrc/host/halsynth.cpp) writes the benchmark alignments: its files must be valid for the reference's
own halValidate, and with --branch 0 their content must be what the reference generator (createRandomGenome, driven by
oracle/gen/halTreeGen.cpp --mode randgen) produces for equal dimensions: the same liftover answers."""
import os
import random
import subprocess

import pytest

from conftest import ROOT, ref_bin

NEWICK = "(((L0,L1)A0,(L2,L3)A1)B0,((L4,L5)A2,(L6)A3)B1,(L7)B2)R;"


@pytest.fixture(scope="module")
def synth():
    from hal_b200 import build
    build.build()
    return os.path.join(ROOT, "hal_b200", "bin", "halSynth")


@pytest.mark.parametrize("branch,seg_len", [("0", 32), ("0.05", 32), ("0.3", 7)])
def test_halsynth_files_pass_the_reference_validator(synth, tmp_path, branch, seg_len):
    validate = ref_bin("halValidate")
    if validate is None:
        pytest.skip("oracle/_ref not built")
    hal = str(tmp_path / "s.hal")
    subprocess.check_call([synth, "--newick", NEWICK, "--segs", "3000", "--segLen", str(seg_len), "--branch", branch, "--seed", "7", hal])
    r = subprocess.run([validate, hal], capture_output=True, text=True)
    assert r.returncode == 0 and "File valid" in r.stdout, r.stdout + r.stderr


def test_halsynth_branch0_lifts_like_the_reference_generator(synth, tmp_path):
    gen, lift = ref_bin("halTreeGen"), ref_bin("halLiftover")
    if gen is None or lift is None:
        pytest.skip("oracle/_ref not built")
    newick, segs, seg_len = "((L0,L1)A0,(L2)A1)R;", 800, 16
    a, b = str(tmp_path / "treegen.hal"), str(tmp_path / "synth.hal")
    subprocess.check_call([gen, "--mode", "randgen", "--newick", newick, "--segs", str(segs), "--minLen", str(seg_len), "--maxLen", str(seg_len),
                           "--branch", "0", "--seed", "3", a])
    subprocess.check_call([synth, "--newick", newick, "--segs", str(segs), "--segLen", str(seg_len), "--branch", "0", "--seed", "3", b])
    rng = random.Random(5)
    glen = segs * seg_len
    lines = []
    for i in range(400):
        ln = rng.randint(1, 300)
        s = rng.randint(0, glen - ln)
        lines.append(f"L0_seq\t{s}\t{s + ln}\tn{i}\t0\t{rng.choice('+-')}")
    bed = tmp_path / "in.bed"
    bed.write_text("\n".join(lines) + "\n")
    outs = []
    for hal in (a, b):
        for tgt in ("L2", "R", "L1"):
            o = str(tmp_path / "o.bed")
            subprocess.check_call([lift, hal, "L0", str(bed), tgt, o])
            outs.append(open(o).read())
    assert outs[:3] == outs[3:] and len(outs[0]) > 0
