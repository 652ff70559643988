/* TEST INFRASTRUCTURE ONLY -- CPU oracle: hal2maf block state machine. */
#ifndef ORACLE_MAF_H
#define ORACLE_MAF_H
#include "columns.h"
#include <string>

namespace oracle {

struct MafOpts {
    ColumnOpts col;
    bool fullNames = true;            /* !--onlySequenceNames */
    bool keepEmptyRefBlocks = false;
    int64_t maxBlockLen = 1000;       /* MafBlock::defaultMaxLength */
    bool unique = false;              /* hal2maf --unique: only columns whose left-most reference-genome base is the column's own */
};

/* hal2maf for one reference genome: refSeq < 0 -> every sequence (one convertSequence call each, shared MafBlock);
 * start/length as in the CLI (length 0 = to the end of the sequence).  Appends the MAF text to out. */
void hal2maf(const HalView &v, int ref, int refSeq, int64_t start, int64_t length, const MafOpts &o, std::string &out);

} // namespace oracle
#endif
