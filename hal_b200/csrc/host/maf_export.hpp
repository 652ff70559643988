// GpuMafExport -- host-side mirror of hal::MafExport + hal::MafBlock (maf/inc/halMafExport.h:30-37,
// maf/inc/halMafBlock.h:117) over the column-run ABI of include/halgpu.h.
//
// The reference advances a ColumnIterator one base at a time and asks MafBlock::canAppendColumn for every
// column (maf/impl/halMafExport.cpp:46-81).  Here the GPU returns maximal runs of collinear columns
// (halgpu_column_runs); the block state machine -- entries multimap, lingering blank entries (lastUsed <= 10),
// ColumnMap keys that persist until defragment() at flush #0, #1000, ... -- is run
// exactly for the first column of every run and in bulk for the rest of it (inside a run canAppendColumn can
// only fail on maxBlockLen).  Row text is decoded on the host from the mapped file's packed DNA.
#pragma once
#include "../../../include/halgpu.h"
#include <cstdint>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

namespace halgpu {

class GpuMafExport {
  public:
    explicit GpuMafExport(halgpu_ctx *ctx);
    ~GpuMafExport();
    // same meaning as MafExport::convertSequence(mafStream, alignment, seq, startPosition, length, targets)
    void convertSequence(std::ostream &mafStream, int refGenome, int refSequence, int64_t startPosition, uint64_t length,
                         const std::vector<int> &targets);
    void setNoDupes(bool v) { _noDupes = v; }
    void setNoAncestors(bool v) { _noAncestors = v; }
    void setUcscNames(bool v) { _ucscNames = v; }
    void setAppend(bool v) { _append = v; }
    void setMaxBlockLength(int64_t v) { _maxLength = v <= 0 ? INT64_MAX : v; }
    // MafBlock::setMaxLength as it is (maf/inc/halMafBlock.h:132-134): 0 makes canAppendColumn fail on every column, i.e.
    // one-column blocks in a release build of the reference (its assert build aborts) -- what halGetMAF's 0 means
    void setMaxBlockLengthRaw(int64_t v) { _maxLength = v; }
    void setOnlyOrthologs(bool v) { _onlyOrthologs = v; }
    void setKeepEmptyRefBlocks(bool v) { _keepEmptyRefBlocks = v; }
    void setUnique(bool v) { _unique = v; } // MafExport::setUnique (maf/inc/halMafExport.h:51-53)
    size_t chunkColumns = 8u << 20; // columns per halgpu_column_runs call
    unsigned formatThreads = defaultFormatThreads(); // threads that decode and print the rows of finished blocks
    static unsigned defaultFormatThreads();
    size_t queueBytes = (size_t)64 << 20;            // finished blocks are formatted and written once about this much text is queued
    // totals
    uint64_t columns = 0, runs = 0, blocks = 0;
    double gpuSeconds = 0;          // halgpu_column_runs calls
    double blockerSeconds = 0;      // the sequential block state machine over the runs
    double textSeconds() const;     // halgpu_maf_text calls (row prefixes, device text, copy back) or the host formatter
    double writeSeconds() const;    // ostream writes of the finished text

  private:
    struct Impl;
    std::unique_ptr<Impl> _impl;
    halgpu_ctx *_ctx;
    bool _noDupes = false, _noAncestors = false, _ucscNames = true, _append = false, _onlyOrthologs = false,
         _keepEmptyRefBlocks = false, _unique = false;
    int64_t _maxLength = 1000; // MafBlock::defaultMaxLength
};

} // namespace halgpu
