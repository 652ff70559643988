import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_intervals(genome_len, n, maxlen, seed, strands=b"+-."):
    rng = np.random.default_rng(seed)
    ln = rng.integers(1, min(maxlen, genome_len) + 1, n)
    gs = rng.integers(0, genome_len - ln + 1)
    return gs.astype(np.int64), (gs + ln - 1).astype(np.int64), rng.choice(np.frombuffer(strands, dtype=np.uint8), n)


def assert_same_as_oracle(off, recs, r):
    assert np.array_equal(off, r["offsets"]), "CSR offsets differ"
    for k, ok in (("start", "start"), ("end", "end"), ("src_start", "srcStart"), ("tgt_seq", "tgtSeq"),
                  ("strand", "strand"), ("src_strand", "srcStrand")):
        assert np.array_equal(recs[k], r[ok]), f"field {k} differs from the oracle"


def bed_to_batch(seqs, bed_text):
    """BED3..9 text -> (rows, gs, ge, strand) the way Liftover::visitLine prepares liftInterval
    (liftover/impl/halLiftover.cpp:46-70, halBlockLiftover.cpp:48-50)."""
    by_name = {n: (s, l) for (n, s, l) in seqs}
    rows, gs, ge, st = [], [], [], []
    strand = "+"
    for line in bed_text.split("\n"):
        if not line.strip():
            continue
        row = line.split("\t")
        bt = min(len(row), 12)
        if bt > 5:
            strand = row[5][0]
        if row[0] not in by_name:
            continue
        s0, e0 = int(row[1]), int(row[2])
        if e0 > by_name[row[0]][1]:
            continue
        rows.append((row, bt))
        gs.append(s0 + by_name[row[0]][0])
        ge.append(e0 - 1 + by_name[row[0]][0])
        st.append(ord(strand))
    return rows, np.array(gs, np.int64), np.array(ge, np.int64), np.array(st, np.uint8)


def batch_to_bed(rows, tseqs, off, recs):
    """Format like BedLine::write (liftover/impl/halBedLine.cpp:104-151) for BED3..6 (+extra columns)."""
    out = []
    for i, (row, bt) in enumerate(rows):
        for j in range(int(off[i]), int(off[i + 1])):
            cols = [tseqs[recs["tgt_seq"][j]][0], str(recs["start"][j]), str(recs["end"][j])]
            if bt > 3:
                cols.append(row[3])
            if bt > 4:
                cols.append(str(int(row[4])))
            if bt > 5:
                cols.append(chr(recs["strand"][j]))
            cols += row[bt:]
            out.append("\t".join(cols))
    return "\n".join(out) + ("\n" if out else "")
