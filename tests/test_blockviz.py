"""The reference's blockViz C API (blockViz/inc/halBlockViz.h) re-implemented over the GPU context:
hal_b200/libhalBlockVizGpu.so + include/halgpu_blockviz.h.  A text driver of the API (tests/cpp/blockviz_cli.cpp) is linked once
against the reference's own implementation (oracle/_ref/blockVizCli) and once against this one; their outputs must be identical.

CPU tier: emulated library vs committed answers of the reference (tests/golden/blockviz) and vs the reference live.
GPU tier: the CUDA build."""
import ctypes
import json
import os
import random
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT, ref_bin

CASES = json.load(open(os.path.join(GOLDEN, "blockviz", "cases.json")))


def check_cases(cli, step=1):
    kinds = set()
    for c in CASES[::step] + [c for c in CASES if c["args"][0] != "blocks"]:
        r = subprocess.run([cli, os.path.join(GOLDEN, c["hal"])] + c["args"], capture_output=True, text=True)
        assert r.returncode == c["rc"], (c["args"], r.stdout, r.stderr)
        assert r.stdout == c["out"], c["args"]
        kinds.add(c["args"][0])
    assert kinds >= {"species", "chroms", "dna", "limits", "maxlod", "blocks", "maf", "meta"}
    assert any(c["hal"].endswith(".lod.txt") and c["args"][0] == "blocks" for c in CASES)  # level-of-detail list files
    assert any("\nD\t" in c["out"] for c in CASES) and any(c["args"][0] == "blocks" and c["args"][9] == "1" for c in CASES)


def test_blockviz_emulated_matches_reference_answers(emul_blockviz_cli):
    check_cases(emul_blockviz_cli, step=2)  # every other committed query (+ all non-"blocks" ones): keeps the CPU tier short


@pytest.mark.skipif(ref_bin("blockVizCli") is None, reason="oracle/_ref not built")
def test_blockviz_emulated_vs_reference_live(emul_blockviz_cli):
    import pyoracle
    hal = os.path.join(GOLDEN, "varlen8.hal")
    o = pyoracle.Oracle(hal)
    rng = random.Random(2)
    for _ in range(15):
        q, t = rng.choice(o.genomes), rng.choice(o.genomes)
        nm, _, ln = rng.choice(o.sequences(o.genome_id(t)))
        L = rng.randint(1, min(ln, 1500))
        a = rng.randint(0, ln - L)
        dup = rng.choice([0, 1, 2])
        rev = "1" if (dup < 2 and rng.random() < 0.3) else "0"
        adj = "1" if (rev == "0" and rng.random() < 0.5) else "0"
        args = ["blocks", q, t, nm, str(a), str(a + L), rev, str(rng.choice([0, 2])), str(dup), adj, "-"]
        r = subprocess.run([ref_bin("blockVizCli"), hal] + args, capture_output=True, text=True)
        m = subprocess.run([emul_blockviz_cli, hal] + args, capture_output=True, text=True)
        if r.returncode < 0:
            continue  # oracle/_ref is built with assertions on; a query that trips one of the reference's asserts has no answer
        assert (r.returncode, r.stdout) == (m.returncode, m.stdout), args


def test_blockviz_unsupported_calls_fail_with_a_message(emul_blockviz_cli):
    hal = os.path.join(GOLDEN, "varlen8.hal")
    r = subprocess.run([emul_blockviz_cli, hal, "maf", "L3", "L3_s0", "0", "100", "3", "1000", "1", "L0"], capture_output=True, text=True)
    assert r.returncode == 1 and "maxRefGap > 0" in r.stdout


def test_blockviz_library_exports_every_declared_symbol(product_lib):
    lib = os.path.join(ROOT, "hal_b200", "libhalBlockVizGpu.so")
    assert os.path.exists(lib), "hal_b200/build.py builds it next to libhalgpu.so"
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "halgpu_blockviz.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(hal[A-Z][A-Za-z_]+)\s*\(", text)))
    assert len(syms) >= 20
    L = ctypes.CDLL(lib)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/halgpu_blockviz.h but not exported"


@pytest.mark.gpu
def test_blockviz_cuda_matches_reference_answers():
    from hal_b200 import build
    build.build()
    check_cases(os.path.join(ROOT, "hal_b200", "bin", "blockVizCli"), step=5)  # (every query is a process + CUDA context)
