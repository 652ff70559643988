/* TEST INFRASTRUCTURE ONLY -- CPU oracle (array-level restatement of the reference liftover walk). */
#ifndef ORACLE_LIFTOVER_H
#define ORACLE_LIFTOVER_H
#include "halview.h"

namespace oracle {

/* One mapped fragment: forward-coordinate inclusive extents of the source and target pieces,
 * their traversal directions, and the target segment containing [tLo,tHi]. */
struct Frag {
    int64_t sLo, sHi, tLo, tHi;
    int64_t idx;   /* array index of the target-side segment */
    int64_t order; /* generation order, tie-break for "first wins" de-duplication */
    bool sRev, tRev, top;
};

/* One output interval of BlockLiftover::liftInterval (liftover/impl/halBlockLiftover.cpp:82-105) */
struct OutLine {
    int32_t tgtSeq;
    int64_t start, end; /* sequence-relative, end exclusive */
    char strand;
    int64_t srcStart;   /* genome-global source start (BedLine::_srcStart) */
    char srcStrand;
    int32_t nFrag;      /* fragments merged into the line (PSL blocks) */
};

struct Stats {
    uint64_t seeds = 0, visitsTop = 0, visitsBot = 0, visitBytes = 0, searchProbes = 0, rawFrags = 0, refinedFrags = 0, outLines = 0;
};

struct Plan {
    int src, tgt, mrca;
    std::vector<int> up;   /* src ... mrca */
    std::vector<int> down; /* mrca ... tgt */
    /* halLiftover --coalescenceLimit: mrca ... the child of the limit genome on the way up (empty when the limit is the
     * MRCA, the default): the genomes whose paralogy rings mapRecursiveParalogies walks (halSegmentMapper.cpp:525-576) */
    std::vector<int> para;
};

/* coal: coalescence limit genome, -1 = the MRCA.  Throws when it is not the MRCA or one of its ancestors (the reference
 * then runs off the root: "Hit root genome when attempting to map paralogies", halSegmentMapper.cpp:543-545). */
Plan makePlan(const HalView &v, int src, int tgt, int coal = -1);

/* Lift one interval [gs,ge] (genome-global inclusive) with BED strand ('+','-','.').
 * Appends the reference-ordered output lines (stable by srcStart) to out; optionally the
 * fragments of each line to fragsOut (run order).
 * runSizesOut (optional) receives, per extracted run in extraction order (BEFORE the stable sort of the lines), the
 * number of fragments it contributed to fragsOut. */
void liftInterval(const HalView &v, const Plan &plan, bool dupes, int64_t gs, int64_t ge, char strand,
                  std::vector<OutLine> &out, Stats *stats, std::vector<Frag> *fragsOut = nullptr,
                  std::vector<size_t> *runSizesOut = nullptr);

} // namespace oracle
#endif
