/* TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle for ctypes (tests/, bench.py cpu_baseline). */
#include "columns.h"
#include "liftover.h"
#include "maf.h"
#include "wiggle.h"
#include <algorithm>
#include <cstdio>
#include <memory>

using namespace oracle;

struct OracleHandle {
    HalView view;
    int coal = -1; /* coalescence limit of the following oracle_liftover calls (-1: MRCA) */
    std::vector<uint64_t> fragOffsets;
    std::vector<Frag> frags;
    std::vector<uint64_t> offsets;
    std::vector<OutLine> lines;
    Stats stats;
    std::string err;
    std::string maf;
};

extern "C" {

void *oracle_open(const char *path) {
    std::unique_ptr<OracleHandle> h(new OracleHandle);
    try {
        h->view.open(path);
    } catch (std::exception &e) {
        fprintf(stderr, "oracle_open: %s\n", e.what());
        return nullptr;
    }
    return h.release();
}
void oracle_close(void *hp) { delete (OracleHandle *)hp; }
int oracle_num_genomes(void *hp) { return (int)((OracleHandle *)hp)->view.genomes.size(); }
const char *oracle_genome_name(void *hp, int g) { return ((OracleHandle *)hp)->view.genomes[g].name.c_str(); }
int oracle_genome_id(void *hp, const char *name) { return ((OracleHandle *)hp)->view.genomeId(name); }
int oracle_genome_parent(void *hp, int g) { return ((OracleHandle *)hp)->view.genomes[g].parent; }
int64_t oracle_genome_length(void *hp, int g) { return ((OracleHandle *)hp)->view.genomes[g].len; }
int64_t oracle_genome_num_top(void *hp, int g) { return ((OracleHandle *)hp)->view.genomes[g].numTop; }
int64_t oracle_genome_num_bottom(void *hp, int g) { return ((OracleHandle *)hp)->view.genomes[g].numBot; }
int oracle_num_sequences(void *hp, int g) { return (int)((OracleHandle *)hp)->view.genomes[g].seqs.size(); }
const char *oracle_seq_name(void *hp, int g, int s) { return ((OracleHandle *)hp)->view.genomes[g].seqs[s].name.c_str(); }
int64_t oracle_seq_start(void *hp, int g, int s) { return ((OracleHandle *)hp)->view.genomes[g].seqs[s].start; }
int64_t oracle_seq_length(void *hp, int g, int s) { return ((OracleHandle *)hp)->view.genomes[g].seqs[s].length; }
const char *oracle_newick(void *hp) { return ((OracleHandle *)hp)->view.newick.c_str(); }

/* Lift n intervals given in genome-global inclusive coordinates.  Returns the total number of output
 * lines (kept inside the handle until the next call); fetch them with oracle_fetch. */
int64_t oracle_liftover(void *hp, int src, int tgt, int noDupes, int64_t n, const int64_t *gs, const int64_t *ge,
                        const char *strand) {
    OracleHandle *h = (OracleHandle *)hp;
    Plan plan;
    try {
        plan = makePlan(h->view, src, tgt, h->coal);
    } catch (std::exception &e) {
        h->err = e.what();
        return -1;
    }
    h->offsets.assign(1, 0);
    h->lines.clear();
    h->stats = Stats();
    for (int64_t i = 0; i < n; i++) {
        liftInterval(h->view, plan, !noDupes, gs[i], ge[i], strand ? strand[i] : '+', h->lines, &h->stats);
        h->offsets.push_back(h->lines.size());
    }
    return (int64_t)h->lines.size();
}

/* halLiftover --coalescenceLimit for the following oracle_liftover calls (genome id, -1 = MRCA) */
void oracle_set_coalescence_limit(void *hp, int genome) { ((OracleHandle *)hp)->coal = genome; }

void oracle_fetch(void *hp, uint64_t *offsets, int32_t *tgtSeq, int64_t *start, int64_t *end, char *strand,
                  int64_t *srcStart, char *srcStrand) {
    OracleHandle *h = (OracleHandle *)hp;
    for (size_t i = 0; i < h->offsets.size(); i++) offsets[i] = h->offsets[i];
    for (size_t i = 0; i < h->lines.size(); i++) {
        const OutLine &o = h->lines[i];
        tgtSeq[i] = o.tgtSeq; start[i] = o.start; end[i] = o.end; strand[i] = o.strand;
        srcStart[i] = o.srcStart; srcStrand[i] = o.srcStrand;
    }
}

/* stats of the last oracle_liftover call: seeds, visitsTop, visitsBot, visitBytes, searchProbes, rawFrags,
 * refinedFrags, outLines */
void oracle_stats(void *hp, uint64_t *out8) {
    const Stats &s = ((OracleHandle *)hp)->stats;
    out8[0] = s.seeds; out8[1] = s.visitsTop; out8[2] = s.visitsBot; out8[3] = s.visitBytes;
    out8[4] = s.searchProbes; out8[5] = s.rawFrags; out8[6] = s.refinedFrags; out8[7] = s.outLines;
}

/* halAlignmentDepth per-base values for reference positions first..last (genome coordinates, inclusive) every `step`.
 * targets: genome ids (nt == 0: all).  Returns the number of values written; visits (optional) receives the landing count. */
int64_t oracle_depth(void *hp, int ref, int64_t first, int64_t last, int64_t step, const int *targets, int nt, int countDupes,
                     int noAncestors, int noDupes, int32_t *out, uint64_t *visits) {
    OracleHandle *h = (OracleHandle *)hp;
    std::vector<int> t(targets, targets + nt);
    ColumnOpts o = makeColumnOpts(h->view, ref, t, noDupes != 0, noAncestors != 0, false);
    std::vector<ColRow> rows;
    int64_t n = 0;
    uint64_t vis = 0;
    for (int64_t p = first; p <= last; p += step) {
        column(h->view, ref, p, o, rows, &vis);
        out[n++] = (int32_t)depthOf(h->view, rows, countDupes != 0);
    }
    if (visits) *visits = vis;
    return n;
}

/* hal2maf text for one reference genome (default flags + noDupes/noAncestors/onlyOrthologs/targets).  Returns a pointer to
 * the text kept inside the handle until the next call; *len receives its length.  NULL on error. */
const char *oracle_hal2maf(void *hp, int ref, int refSeq, int64_t start, int64_t length, const int *targets, int nt, int noDupes,
                           int noAncestors, int onlyOrthologs, int onlySequenceNames, int keepEmptyRefBlocks, int64_t maxBlockLen,
                           int unique, uint64_t *len) {
    OracleHandle *h = (OracleHandle *)hp;
    try {
        MafOpts o;
        std::vector<int> t(targets, targets + nt);
        o.col = makeColumnOpts(h->view, ref, t, noDupes != 0, noAncestors != 0, onlyOrthologs != 0);
        o.fullNames = !onlySequenceNames;
        o.keepEmptyRefBlocks = keepEmptyRefBlocks != 0;
        if (maxBlockLen > 0) o.maxBlockLen = maxBlockLen;
        o.unique = unique != 0;
        h->maf.clear();
        hal2maf(h->view, ref, refSeq, start, length, o, h->maf);
    } catch (std::exception &e) {
        fprintf(stderr, "oracle_hal2maf: %s\n", e.what());
        return nullptr;
    }
    *len = h->maf.size();
    return h->maf.c_str();
}

/* ColumnLiftover::liftInterval restated (liftover/impl/halColumnLiftover.cpp:21-92) for one interval [gs,ge] of src:
 * union of the target-genome bases of every column, split by (sequence, strand), merged into maximal runs.  With
 * unique=true the reference skips columns whose reference base was already seen as a paralog of an earlier column; the
 * union over all columns is the same set.  Output (kept in the handle): one line per run, forward-strand runs first, each
 * group by sequence index then by position (the reference orders sequences by POINTER value, so compare after sorting).
 * Returns the number of lines; fetch with oracle_fetch (srcStart = -1, srcStrand as strand). */
int64_t oracle_column_liftover(void *hp, int src, int tgt, int noDupes, int64_t gs, int64_t ge, char strand) {
    OracleHandle *h = (OracleHandle *)hp;
    ColumnOpts o = makeColumnOpts(h->view, src, std::vector<int>{tgt}, noDupes != 0, false, false);
    std::vector<ColRow> rows;
    std::vector<std::pair<int64_t, int>> hits[2]; /* (pos, seq) per strand */
    const GenomeView &T = h->view.genomes[tgt];
    for (int64_t p = gs; p <= ge; ++p) {
        column(h->view, src, p, o, rows);
        for (const ColRow &r : rows) {
            if (r.genome != tgt) continue;
            bool rev = r.rev != (strand == '-'); /* reverseStrand iterator flips every row */
            hits[rev ? 1 : 0].push_back(std::make_pair(r.pos, T.seqOf(r.pos)));
        }
    }
    h->offsets.assign(1, 0);
    h->lines.clear();
    for (int s = 0; s < 2; ++s) {
        auto &v = hits[s];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        for (size_t i = 0; i < v.size();) {
            size_t j = i + 1;
            while (j < v.size() && v[j].first == v[j - 1].first + 1 && v[j].second == v[i].second) ++j;
            OutLine l;
            l.tgtSeq = v[i].second;
            l.start = v[i].first - T.seqs[l.tgtSeq].start;
            l.end = v[j - 1].first + 1 - T.seqs[l.tgtSeq].start;
            l.strand = strand == '.' ? '.' : (s ? '-' : '+');
            l.srcStart = -1; l.srcStrand = l.strand; l.nFrag = 1;
            h->lines.push_back(l);
            i = j;
        }
    }
    h->offsets.push_back(h->lines.size());
    return (int64_t)h->lines.size();
}

/* The mapped fragments (after insertAndBreakOverlaps, as handed to extractSegment) of n '+' intervals: returns their total
 * number; oracle_fetch_frags copies offsets[n+1] and per fragment sLo, tLo, length, tRev. */
int64_t oracle_liftover_frags(void *hp, int src, int tgt, int noDupes, int64_t n, const int64_t *gs, const int64_t *ge) {
    OracleHandle *h = (OracleHandle *)hp;
    Plan plan = makePlan(h->view, src, tgt, h->coal);
    h->fragOffsets.assign(1, 0);
    h->frags.clear();
    std::vector<OutLine> lines;
    for (int64_t i = 0; i < n; i++) {
        lines.clear();
        liftInterval(h->view, plan, !noDupes, gs[i], ge[i], '+', lines, nullptr, &h->frags);
        h->fragOffsets.push_back(h->frags.size());
    }
    return (int64_t)h->frags.size();
}
void oracle_fetch_frags(void *hp, uint64_t *offsets, int64_t *sLo, int64_t *tLo, int64_t *len, uint8_t *tRev) {
    OracleHandle *h = (OracleHandle *)hp;
    for (size_t i = 0; i < h->fragOffsets.size(); i++) offsets[i] = h->fragOffsets[i];
    for (size_t i = 0; i < h->frags.size(); i++) {
        const Frag &f = h->frags[i];
        sLo[i] = f.sLo; tLo[i] = f.tLo; len[i] = f.sHi - f.sLo + 1; tRev[i] = f.tRev ? 1 : 0;
    }
}

/* halWiggleLiftover text -> text (oracle/restate/wiggle.cpp).  Returns the output text (kept in the handle) and its length,
 * or NULL with the reference's exception message in the handle (oracle_last_error). */
const char *oracle_wiggle_liftover(void *hp, int src, int tgt, int noDupes, const char *inText, const char *preloadText, int correctPath,
                                   uint64_t *len) {
    OracleHandle *h = (OracleHandle *)hp;
    try {
        std::string pre = preloadText ? preloadText : "";
        h->maf = wiggleLiftover(h->view, src, tgt, !noDupes, inText, preloadText ? &pre : nullptr, correctPath != 0);
    } catch (std::exception &e) {
        h->err = e.what();
        return nullptr;
    }
    *len = h->maf.size();
    return h->maf.c_str();
}
const char *oracle_last_error(void *hp) { return ((OracleHandle *)hp)->err.c_str(); }

} // extern "C"
