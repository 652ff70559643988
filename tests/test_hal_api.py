"""CPU tier: the read-only hal:: query surface (hal_b200/csrc/host/hal_api.hpp) -- single-base lifts written with the
reference's iterator vocabulary must agree with the oracle's --noDupes liftover, base by base."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def walker(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "hal_api_walk")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", out, os.path.join(ROOT, "tests", "cpp", "hal_api_walk.cpp"),
                           os.path.join(ROOT, "hal_b200", "csrc", "halmmap.cpp")])
    return out


@pytest.mark.parametrize("hal,src,tgt", [("varlen8.hal", "L0", "L3"), ("varlen8.hal", "L3", "L1"), ("varlen8.hal", "R", "L2"),
                                         ("varlen8.hal", "A2", "R"), ("varlen8.hal", "A0", "L1"), ("varlen8.hal", "L2", "L2"),
                                         ("refBedLiftoverTest.hal", "leaf3", "leaf1"), ("randgenSmallSeed0.hal", "Genome_3", "Genome_2")])
def test_iterator_walk_equals_oracle(walker, oracle_lib, hal, src, tgt):
    path = os.path.join(GOLDEN, hal)
    lines = subprocess.check_output([walker, path, src, tgt, "3000", "7"], text=True).splitlines()
    o = oracle_lib.Oracle(path)
    assert lines[0].startswith(f"# {src} {tgt} ")
    s, t = o.genome_id(src), o.genome_id(tgt)
    tseqs = o.sequences(t)
    pos = np.array([int(l.split()[0]) for l in lines[1:]], np.int64)
    r = o.liftover(s, t, pos, pos, None, no_dupes=True)
    off = r["offsets"]
    complement = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "n"}
    for i, l in enumerate(lines[1:]):
        f = l.split()
        n = int(off[i + 1] - off[i])
        if f[1] == "-":
            assert n == 0, l
        else:
            assert n == 1, l
            j = int(off[i])
            assert (tseqs[r["tgtSeq"][j]][0], int(r["start"][j]), chr(r["strand"][j])) == (f[1], int(f[2]), f[3]), l
            assert f[4] in complement  # getString of a 1-base slice is a nucleotide
