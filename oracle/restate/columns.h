/* TEST INFRASTRUCTURE ONLY -- CPU oracle: column walk for the default
 * hal2maf / halAlignmentDepth flags (unique=false, maxRefGap=0: every reference base is an independent query). */
#ifndef ORACLE_COLUMNS_H
#define ORACLE_COLUMNS_H
#include "halview.h"

namespace oracle {

struct ColRow {
    int32_t genome;
    int64_t pos;   /* forward genome coordinate */
    bool rev;
};

struct ColumnOpts {
    bool noDupes = false, noAncestors = false, onlyOrthologs = false;
    std::vector<char> inScope;   /* per genome: traversal allowed (spanning tree of targets + reference); empty = all */
    std::vector<char> isTarget;  /* per genome: rows reported; empty = all */
};

/* rows of the alignment column of reference base p, in DFS discovery order (reference row first) */
void column(const HalView &v, int ref, int64_t p, const ColumnOpts &o, std::vector<ColRow> &rows, uint64_t *visits = nullptr);

/* halAlignmentDepth::printSequence per-base count (alignmentDepth/halAlignmentDepth.cpp:262-280) */
int64_t depthOf(const HalView &v, const std::vector<ColRow> &rows, bool countDupes);

ColumnOpts makeColumnOpts(const HalView &v, int ref, const std::vector<int> &targets, bool noDupes, bool noAncestors,
                          bool onlyOrthologs);

} // namespace oracle
#endif
