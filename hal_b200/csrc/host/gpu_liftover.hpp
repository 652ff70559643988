// GpuBlockLiftover -- host-side mirror of hal::Liftover / hal::BlockLiftover (liftover/inc/halLiftover.h:20-44,
// liftover/inc/halBlockLiftover.h:19-26) over the C ABI of include/halgpu.h.
//
// Same call shape as Liftover::convert(alignment, srcGenome, istream*, tgtGenome, ostream*, bedType, traverseDupes,
// outPSL, outPSLWithName, coalescenceLimit); genomes are indices of the staged context instead of Genome*.
// The reference lifts line by line (BedScanner::scan -> visitLine -> liftInterval); here lines are read in
// batches, every BED interval (or BED12 block) of a batch becomes one element of a single halgpu_liftover call,
// and the per-line post-processing (BED12 regrouping, thick/blocks clean-up, stable order by source start,
// halLiftover.cpp:72-92) runs on the host over the returned records.  Output text is byte-identical.
#pragma once
#include "../../../include/halgpu.h"
#include "bed.hpp"
#include <cstddef>
#include <iosfwd>
#include <set>
#include <string>

namespace halgpu {

class GpuBlockLiftover {
  public:
    explicit GpuBlockLiftover(halgpu_ctx *ctx) : _ctx(ctx) {}
    void convert(int srcGenome, std::istream *inBed, int tgtGenome, std::ostream *outBed, int bedType = 0,
                 bool traverseDupes = true, bool outPSL = false, bool outPSLWithName = false, int coalescenceLimit = -1);
    size_t batchLines = 1u << 20; // BED lines per GPU call on the serial text path
    // The input is consumed in blocks of whole lines of about this many bytes.  A block whose lines are plain BED3..BED9
    // of one width goes through the multi-threaded text layer (bed_fast.hpp) and one GPU call; any other block through
    // the serial BedLine code in batches of batchLines.  textThreads == 0 disables the fast path (env HALGPU_TEXT_THREADS
    // overrides it).
    size_t blockBytes = 24u << 20; // (small enough for the read / lift / write pipeline of the fast path to overlap)
    unsigned textThreads = defaultTextThreads();
    static unsigned defaultTextThreads();
    // lift with hal::ColumnLiftover::liftInterval semantics (liftover/inc/halColumnLiftover.h:19-26) instead of
    // BlockLiftover's: the reference compiles that class into libHalLiftover but no CLI instantiates it
    bool columnLiftover = false;
    // totals of the last convert()
    size_t linesIn = 0, intervalsLifted = 0, linesOut = 0;
    double gpuSeconds = 0, textSeconds = 0, writeSeconds = 0, parseSeconds = 0, readSeconds = 0; // halgpu_liftover calls / fast-path parse+format / ostream writes
    size_t fastLines = 0;                                     // input lines that took the multi-threaded text path

  private:
    halgpu_ctx *_ctx;
    std::set<std::string> _missed;
};

} // namespace halgpu
