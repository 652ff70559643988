"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The BED text layer mirrors the reference's BedLine::read/write for BED3..BED9 (+ pass-through extra
columns) -- liftover/impl/halBedLine.cpp:27-151 -- so oracle output can be compared byte for byte with
oracle/_ref/halLiftover.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF = os.path.join(HERE, "_ref")


def build(ref=False):
    """Compile the restatement (and, where /root/reference exists, the reference binaries)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "restate"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def _lib():
    if not os.path.exists(LIB):
        build()
    L = C.CDLL(LIB)
    L.oracle_open.restype = C.c_void_p
    L.oracle_open.argtypes = [C.c_char_p]
    L.oracle_close.argtypes = [C.c_void_p]
    for f in ("oracle_genome_name", "oracle_newick"):
        getattr(L, f).restype = C.c_char_p
    L.oracle_genome_name.argtypes = [C.c_void_p, C.c_int]
    L.oracle_newick.argtypes = [C.c_void_p]
    L.oracle_num_genomes.argtypes = [C.c_void_p]
    L.oracle_genome_id.argtypes = [C.c_void_p, C.c_char_p]
    L.oracle_genome_parent.argtypes = [C.c_void_p, C.c_int]
    for f in ("oracle_genome_length", "oracle_genome_num_top", "oracle_genome_num_bottom"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [C.c_void_p, C.c_int]
    L.oracle_num_sequences.argtypes = [C.c_void_p, C.c_int]
    L.oracle_seq_name.restype = C.c_char_p
    L.oracle_seq_name.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for f in ("oracle_seq_start", "oracle_seq_length"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.oracle_liftover.restype = C.c_int64
    L.oracle_liftover.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oracle_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.oracle_set_coalescence_limit.argtypes = [C.c_void_p, C.c_int]
    L.oracle_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.oracle_hal2maf.restype = C.c_void_p
    L.oracle_hal2maf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]
    L.oracle_column_liftover.restype = C.c_int64
    L.oracle_column_liftover.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_char]
    L.oracle_liftover_frags.restype = C.c_int64
    L.oracle_liftover_frags.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.oracle_fetch_frags.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.oracle_wiggle_liftover.restype = C.c_void_p
    L.oracle_wiggle_liftover.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p]
    L.oracle_last_error.restype = C.c_char_p
    L.oracle_last_error.argtypes = [C.c_void_p]
    L.oracle_depth.restype = C.c_int64
    L.oracle_depth.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_void_p, C.c_void_p]
    return L


class Oracle:
    def __init__(self, path):
        self.L = _lib()
        self.h = self.L.oracle_open(path.encode())
        if not self.h:
            raise RuntimeError("oracle_open failed: " + path)
        self.genomes = [self.L.oracle_genome_name(self.h, g).decode() for g in range(self.L.oracle_num_genomes(self.h))]

    def close(self):
        if self.h:
            self.L.oracle_close(self.h)
            self.h = None

    def genome_id(self, name):
        return self.L.oracle_genome_id(self.h, name.encode())

    def sequences(self, g):
        return [(self.L.oracle_seq_name(self.h, g, s).decode(), self.L.oracle_seq_start(self.h, g, s),
                 self.L.oracle_seq_length(self.h, g, s)) for s in range(self.L.oracle_num_sequences(self.h, g))]

    def genome_length(self, g):
        return self.L.oracle_genome_length(self.h, g)

    def liftover(self, src, tgt, gs, ge, strand=None, no_dupes=False, coalescence_limit=None):
        """gs/ge: genome-global inclusive int64 arrays.  Returns dict of numpy arrays (CSR by interval)."""
        gs = np.ascontiguousarray(gs, dtype=np.int64)
        ge = np.ascontiguousarray(ge, dtype=np.int64)
        n = len(gs)
        st = np.full(n, ord('+'), dtype=np.uint8) if strand is None else np.ascontiguousarray(strand, dtype=np.uint8)
        self.L.oracle_set_coalescence_limit(self.h, -1 if coalescence_limit is None else coalescence_limit)
        tot = self.L.oracle_liftover(self.h, src, tgt, int(no_dupes), n, gs.ctypes.data, ge.ctypes.data, st.ctypes.data)
        self.L.oracle_set_coalescence_limit(self.h, -1)
        if tot < 0:
            raise RuntimeError(self.L.oracle_last_error(self.h).decode())
        r = dict(offsets=np.zeros(n + 1, np.uint64), tgtSeq=np.zeros(tot, np.int32), start=np.zeros(tot, np.int64),
                 end=np.zeros(tot, np.int64), strand=np.zeros(tot, np.uint8), srcStart=np.zeros(tot, np.int64),
                 srcStrand=np.zeros(tot, np.uint8))
        self.L.oracle_fetch(self.h, *[r[k].ctypes.data for k in
                                      ("offsets", "tgtSeq", "start", "end", "strand", "srcStart", "srcStrand")])
        s = np.zeros(8, np.uint64)
        self.L.oracle_stats(self.h, s.ctypes.data)
        r["stats"] = dict(zip(("seeds", "visitsTop", "visitsBot", "visitBytes", "searchProbes", "rawFrags",
                               "refinedFrags", "outLines"), (int(x) for x in s)))
        return r

    def depth(self, ref, first, last, step=1, targets=(), count_dupes=False, no_ancestors=False, no_dupes=False):
        """halAlignmentDepth values for genome positions first..last (inclusive).  Returns (int32 array, visits)."""
        n = (last - first) // step + 1
        out = np.zeros(n, np.int32)
        t = np.ascontiguousarray(list(targets), dtype=np.int32)
        vis = C.c_uint64(0)
        got = self.L.oracle_depth(self.h, ref, first, last, step, t.ctypes.data if len(t) else None, len(t), int(count_dupes),
                                  int(no_ancestors), int(no_dupes), out.ctypes.data, C.byref(vis))
        assert got == n
        return out, vis.value

    def depth_wig(self, ref_name, **kw):
        """Text of `halAlignmentDepth <hal> <ref>` (whole genome; alignmentDepth/halAlignmentDepth.cpp:215-346)."""
        g = self.genome_id(ref_name)
        step = kw.get("step", 1)
        out = []
        for (name, start, length) in self.sequences(g):
            if length == 0:
                continue
            d, _ = self.depth(g, start, start + length - 1, **kw)
            out.append(f"fixedStep chrom={name} start=1 step={step}\n" + "".join(f"{x}\n" for x in d))
        return "".join(out)

    def column_liftover(self, src, tgt, gs, ge, strand="+", no_dupes=False):
        """hal::ColumnLiftover::liftInterval for ONE interval: list of (tgtSeq, start, end, strand), forward-strand runs first,
        each group by (sequence index, start).  The reference orders sequences by pointer value: compare after sorting."""
        k = self.L.oracle_column_liftover(self.h, src, tgt, int(no_dupes), gs, ge, strand.encode())
        r = dict(offsets=np.zeros(2, np.uint64), tgtSeq=np.zeros(k, np.int32), start=np.zeros(k, np.int64), end=np.zeros(k, np.int64),
                 strand=np.zeros(k, np.uint8), srcStart=np.zeros(k, np.int64), srcStrand=np.zeros(k, np.uint8))
        self.L.oracle_fetch(self.h, *[r[x].ctypes.data for x in ("offsets", "tgtSeq", "start", "end", "strand", "srcStart", "srcStrand")])
        return [(int(r["tgtSeq"][j]), int(r["start"][j]), int(r["end"][j]), chr(r["strand"][j])) for j in range(k)]

    def hal2maf(self, ref_name, ref_seq=None, start=0, length=0, targets=(), no_dupes=False, no_ancestors=False,
                only_orthologs=False, only_sequence_names=False, keep_empty_ref_blocks=False, max_block_len=0, unique=False):
        """MAF text of `hal2maf --refGenome ref [--refSequence s --start a --length n] ...` (bytes)."""
        g = self.genome_id(ref_name)
        si = -1
        if ref_seq is not None:
            si = [n for (n, _, _) in self.sequences(g)].index(ref_seq)
        t = np.ascontiguousarray([self.genome_id(x) for x in targets], dtype=np.int32)
        n = C.c_uint64(0)
        p = self.L.oracle_hal2maf(self.h, g, si, start, length, t.ctypes.data if len(t) else None, len(t), int(no_dupes),
                                  int(no_ancestors), int(only_orthologs), int(only_sequence_names), int(keep_empty_ref_blocks),
                                  max_block_len, int(unique), C.byref(n))
        if not p:
            raise RuntimeError("oracle_hal2maf failed")
        return C.string_at(p, n.value)

    def liftover_frags(self, src, tgt, gs, ge, no_dupes=False, coalescence_limit=None):
        """Per '+' interval the mapped fragments [(sLo, tLo, length, tRev)] (source forward; target reversed when tRev)."""
        gs = np.ascontiguousarray(gs, dtype=np.int64)
        ge = np.ascontiguousarray(ge, dtype=np.int64)
        self.L.oracle_set_coalescence_limit(self.h, -1 if coalescence_limit is None else coalescence_limit)
        k = self.L.oracle_liftover_frags(self.h, src, tgt, int(no_dupes), len(gs), gs.ctypes.data, ge.ctypes.data)
        self.L.oracle_set_coalescence_limit(self.h, -1)
        off = np.zeros(len(gs) + 1, np.uint64)
        s, t, ln, rv = np.zeros(k, np.int64), np.zeros(k, np.int64), np.zeros(k, np.int64), np.zeros(k, np.uint8)
        self.L.oracle_fetch_frags(self.h, off.ctypes.data, s.ctypes.data, t.ctypes.data, ln.ctypes.data, rv.ctypes.data)
        return [[(int(s[j]), int(t[j]), int(ln[j]), bool(rv[j])) for j in range(int(off[i]), int(off[i + 1]))] for i in range(len(gs))]

    def wiggle_liftover(self, src_name, tgt_name, wig_text, no_dupes=False, preload_text=None, correct_path=False):
        """Output text of halWiggleLiftover (str); raises RuntimeError carrying the reference's exception message.
        correct_path: do not reproduce the reference's wrong turn at the MRCA (oracle/restate/wiggle.cpp)."""
        n = C.c_uint64(0)
        p = self.L.oracle_wiggle_liftover(self.h, self.genome_id(src_name), self.genome_id(tgt_name), int(no_dupes), wig_text.encode(),
                                          None if preload_text is None else preload_text.encode(), int(correct_path), C.byref(n))
        if not p:
            raise RuntimeError(self.L.oracle_last_error(self.h).decode())
        return C.string_at(p, n.value).decode()

    def liftover_bed(self, src_name, tgt_name, bed_text, no_dupes=False, coalescence_limit=None):
        """BED3..BED9 text in -> text out, formatted as halLiftover would (no BED12 regrouping / PSL)."""
        src, tgt = self.genome_id(src_name), self.genome_id(tgt_name)
        sseq = {n: (s, l) for (n, s, l) in self.sequences(src)}
        tseq = self.sequences(tgt)
        rows, gs, ge, st = [], [], [], []
        strand = '+'
        for line in bed_text.split("\n"):
            if not line.strip():
                continue
            row = line.split("\t")
            bt = min(len(row), 12)
            if bt > 9:
                raise ValueError("BED12 not handled by the python oracle front end")
            if bt > 5:
                strand = row[5][0]
            if row[0] not in sseq:
                continue
            s0, e0 = int(row[1]), int(row[2])
            if e0 > sseq[row[0]][1]:
                continue
            rows.append((row, bt, strand))
            gs.append(s0 + sseq[row[0]][0])
            ge.append(e0 - 1 + sseq[row[0]][0])
            st.append(ord(strand))
        r = self.liftover(src, tgt, gs, ge, st, no_dupes,
                          None if coalescence_limit is None else self.genome_id(coalescence_limit))
        out = []
        off = r["offsets"]
        for i, (row, bt, strand) in enumerate(rows):
            for j in range(int(off[i]), int(off[i + 1])):
                cols = [tseq[r["tgtSeq"][j]][0], str(r["start"][j]), str(r["end"][j])]
                if bt > 3:
                    cols.append(row[3])
                if bt > 4:
                    cols.append(str(int(row[4])))
                if bt > 5:
                    cols.append(chr(r["strand"][j]))
                if bt > 6:
                    ts, te = int(row[6]), (int(row[7]) if bt > 7 else 0)
                    if ts != 0 or te != 0:
                        ts, te = int(r["start"][j]), int(r["end"][j])
                    cols.append(str(ts))
                    if bt > 7:
                        cols.append(str(te))
                if bt > 8:
                    rgb = [int(x) for x in row[8].split(",")]
                    rgb = (rgb + [rgb[0]] * 3)[:3] if len(rgb) == 1 else (rgb + [rgb[0]])[:3]
                    cols.append(",".join(str(x) for x in rgb))
                cols += row[bt:]
                out.append("\t".join(cols))
        return "\n".join(out) + ("\n" if out else ""), r["stats"]
