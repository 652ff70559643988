// halSynteny on the GPU context -- host-side mirror of hal::Hal2Psl (synteny/inc/hal2psl.h:22-33), dag_merge
// (synteny/inc/psl_merger.h) and psl_io (synteny/inc/psl_io.h) over the C ABI.
//
// Hal2Psl::convert2psl lifts every query chromosome as ONE interval [0, length) through BlockLiftover::liftInterval
// (synteny/impl/hal2psl.cpp:20-53): halMapSegment for each of its source segments, ONE MappedSegmentSet for the whole
// chromosome (insertAndBreakOverlaps against everything mapped so far) and one extractSegment sweep.  The mapping is the
// expensive, pointer-chasing part and runs on the GPU exactly like for halLiftover: the chromosome is cut at source-segment
// boundaries into windows of a few dozen segments, every window is one interval of a halgpu_liftover call with
// HALGPU_RAW_FRAGMENTS, and the kernel returns the mapped fragments.  What makes the chromosome one interval -- the common
// refinement of ALL target extents and the merge sweep over the whole sorted set -- is a sort/scan over the fragment list,
// done here on the host (O(F log F) for F fragments; the per-warp phase 2 of the kernel is the same algorithm at
// interval scale).  Every merged line becomes one PslBlock (Hal2Psl::makeUpPsl, hal2psl.cpp:55-91), dag_merge chains them.
#pragma once
#include "../../../include/halgpu.h"
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace halgpu {

struct PslBlock { // synteny/inc/psl.h:10-38
    uint64_t qStart = 0, qEnd = 0, tStart = 0, tEnd = 0, size = 0;
    std::string strand, qName, tName;
    uint64_t qSize = 0, tSize = 0;
};

class GpuHal2Psl {
  public:
    explicit GpuHal2Psl(halgpu_ctx *ctx) : _ctx(ctx) {}
    // Hal2Psl::convert2psl(alignment, srcGenome, tgtGenome, srcChrom); srcChrom "\"\"" (the CLI's default string) = every
    // chromosome of the query genome
    std::vector<PslBlock> convert2psl(int srcGenome, int tgtGenome, const std::string &srcChrom);
    size_t segmentsPerInterval = 48; // source segments per GPU interval (one warp; its fragment list lives in shared memory)
    // totals
    size_t intervals = 0, fragments = 0, refined = 0, lines = 0;
    double gpuSeconds = 0, hostSeconds = 0;

  private:
    halgpu_ctx *_ctx;
};

// dag_merge (synteny/impl/psl_merger.cpp:108-136)
std::vector<std::vector<PslBlock>> dagMerge(const std::vector<PslBlock> &blocks, uint64_t minBlockBreath, uint64_t maxAnchorDistance);
// psl_io::write_psl (synteny/impl/psl_io.cpp:85-90) and psl_io::get_blocks_set (:20-28)
void writePsl(const std::vector<std::vector<PslBlock>> &mergedBlocks, std::ostream &os);
std::vector<PslBlock> readPslBlocks(const std::string &pslPath);

} // namespace halgpu
