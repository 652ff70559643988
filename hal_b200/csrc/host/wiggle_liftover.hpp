// GpuWiggleLiftover -- host-side mirror of hal::WiggleLiftover (liftover/inc/halWiggleLiftover.h:21-33) over the C ABI.
//
// Same call shape: preloadOutput(tgtGenome, istream*) then convert(srcGenome, istream*, tgtGenome, ostream*,
// traverseDupes, unique); genomes are indices of the staged context.  The reference scans the wiggle file line by line
// (WiggleScanner::scan, liftover/impl/halWiggleScanner.cpp:39-69), maps the lines that fall into the current source
// segment as one batch (WiggleLiftover::visitLine / mapSegment, halWiggleLiftover.cpp:72-131) and keeps the maximum per
// target base in WiggleTiles<double>.  Here the scanner (same dialect, same quirks, same messages) turns the whole
// file into runs of consecutive source bases with their values, ONE halgpu_wiggle_liftover call maps and scatters
// them on the GPU, and write() prints the bases that hold a value exactly like WiggleLiftover::write (:160-198).
//
// The batch structure of the reference does not change any value (max is commutative and idempotent); it only decides
// when "Coordinate out of order" is raised, so the host replays just that state machine over the segment boundaries.
//
// One deliberate difference: the reference hands halMapSegment the src/tgt SPANNING tree as genomesOnPath
// (halWiggleLiftover.cpp:51-54), which makes mapRecursiveDown turn into the source-side child of the MRCA whenever that
// child precedes the target-side one, and then throw "Could not find correct child that leads from <src> to <tgt>".
// This implementation always walks the correct path (the one halLiftover uses) and produces the lifted values there.
#pragma once
#include "../../../include/halgpu.h"
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace halgpu {

class GpuWiggleLiftover {
  public:
    explicit GpuWiggleLiftover(halgpu_ctx *ctx) : _ctx(ctx) {}
    // --append: load an existing target wiggle so that it is merged with the lifted values (WiggleLiftover::preloadOutput)
    void preloadOutput(int tgtGenome, std::istream *inputFile);
    void convert(int srcGenome, std::istream *inputFile, int tgtGenome, std::ostream *outputFile, bool traverseDupes = true,
                 bool unique = false);

    static const double DefaultValue; // 0.0 (halWiggleLiftover.cpp:17)
    // consecutive single-base lines are packed into runs of at most this many bases (one warp maps one run)
    size_t maxRunBases = 2048;
    // a line whose span reaches this many bases is sent as one run with a single value instead of per-base copies
    int64_t singleValueSpan = 64;
    // threads of the strict multi-threaded pre-parse of the data lines and of the output formatter (0: serial scanner only;
    // env HALGPU_TEXT_THREADS overrides)
    unsigned textThreads = defaultTextThreads();
    static unsigned defaultTextThreads();
    bool fastParsed = false; // the last convert() took the multi-threaded pre-parse
    // totals of the last convert()
    size_t linesIn = 0, runs = 0, basesIn = 0, basesOut = 0;
    double parseSeconds = 0, gpuSeconds = 0, writeSeconds = 0;
    float kernelMs = 0;

  private:
    halgpu_ctx *_ctx;
    std::vector<int64_t> _prePos;
    std::vector<double> _preVal;
};

} // namespace halgpu
