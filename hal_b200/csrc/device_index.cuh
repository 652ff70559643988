// Device-resident index layout (HBM) and the per-launch parameter blocks.
//
// The file's AoS records (top 40 B, bottom 8*(2+nc)+roundup8(nc) B) are re-packed
// once at staging time into sector-friendly arrays; the file format itself is untouched.
//   TopRec   32 B  == one DRAM sector: start | parent<<1|reversed | bottomParseIndex | nextParalogyIndex
//   BotCore  16 B : start | topParseIndex          (what every walk needs from a bottom record)
//   childEnc  8 B : child<<1|reversed, one array per child slot (a walk only ever follows ONE slot per level,
//                   so the other slots' columns are never fetched -- AoS bottoms of an 8-child genome are 88 B)
// Both arrays keep the file's +1 sentinel record so that length(i) = start(i+1) - start(i)
// (api/mmap_impl/mmapTopSegment.h:78-80).
#pragma once
#include <cstdint>

#if defined(HALGPU_SIMT_EMUL)
#include "simt_emul.h"
#else
#include <cuda_runtime.h>
#endif

#include "../../include/halgpu.h"

namespace halgpu {

struct alignas(32) TopRec {
    int64_t start;
    int64_t parentEnc; // -1: no parent; else (parentIndex << 1) | parentReversed
    int64_t botParse;  // bottomParseIndex (-1 for leaves)
    int64_t nextPara;  // nextParalogyIndex (-1: none)
};

struct alignas(16) BotCore {
    int64_t start;
    int64_t topParse; // topParseIndex (-1 for the root)
};

// One genome on the src -> mrca -> tgt path.  Entry p describes genome path[p] and the transition p -> p+1.
struct PathStep {
    const TopRec *top;
    const BotCore *bot;
    const int64_t *child; // down transitions: childEnc column of path[p] for the slot of path[p+1];
                          // up transitions: childEnc column of path[p+1] (the parent) for the slot of path[p]
    int64_t numTop, numBot;
    int32_t up;    // 1: transition p -> p+1 goes to the parent
    int32_t flags; // STEP_* (only set on paths planned with a coalescence limit above the MRCA)
    int32_t jump;  // STEP_PARA: path position at which this genome's paralogs start their way back down to the MRCA
    int32_t pad;
};
// halLiftover --coalescenceLimit (mapRecursiveParalogies, api/impl/halSegmentMapper.cpp:525-576): between the upward and the
// downward part of the path sit the genomes from the MRCA up to the child of the limit (STEP_PARA entries, walked upward),
// followed by the same genomes walked back down WITHOUT following paralogy rings (STEP_NODUPES entries).
enum : int32_t {
    STEP_PARA = 1,      // at this genome: every fragment (as top pieces) forks into (a) itself + its paralogy ring, which
                        // continue at `jump`, and (b) itself mapped to the parent, unless STEP_PARA_LAST
    STEP_NODUPES = 2,   // downward transition that does not start ring walks (mapRecursiveDown with doDupes = false)
    STEP_PARA_LAST = 4  // the parent of this genome is the coalescence limit: nothing goes further up
};

struct LiftParams {
    const PathStep *steps;
    int32_t P;      // genomes on the path (>= 1)
    int32_t dupes;  // !--noDupes
    int32_t columnMerge;     // HALGPU_COLUMN_LIFTOVER: phase 2 of hal::ColumnLiftover instead of BlockLiftover's
    int32_t upCanonicalOnly; // ... whose noDupes mode lets only canonical paralogs map up
    // seeds come from the source genome's top array if it has one, else its bottom array
    // (liftover/impl/halBlockLiftover.cpp:24-30)
    int32_t srcIsTop;
    int32_t srcShift;
    int64_t srcN;
    int64_t srcLen;            // length of the source genome (inputs are validated on the device)
    const uint32_t *srcBucket; // srcBucket[b] = index of the segment containing position b << srcShift
    int64_t srcNumBuckets;
    // target genome sequence starts (numSeq + 1 entries, last = genome length)
    const int64_t *tgtSeqStart;
    int32_t tgtNumSeq;
    int32_t pad0;
    // batch
    int64_t n;               // work items of this launch
    const uint32_t *work;    // optional: work[w] = interval id (sorted order or retry subset); NULL: identity
    const int64_t *gs, *ge;  // genome-global inclusive
    const uint8_t *strand;   // may be NULL
    // outputs
    uint32_t *outCount;      // per interval
    uint64_t *outOffset;     // per interval: first record in pool
    uint32_t *status;        // per interval: ST_*
    halgpu_lift_rec *pool;
    uint32_t *pslPool;        // optional (HALGPU_PSL): 4 counters per pool record (matches, misMatches, repMatches, nCount), zeroed
    const uint8_t *srcDna, *tgtDna; // packed nibbles of the source / target genome (PSL only)
    unsigned long long *poolCursor;
    uint64_t poolCap;
    // wiggle mode (halgpu_wiggle_liftover): no result lists; every mapped fragment scatters the source values of its
    // bases into wigKeys[target base] with atomicMax as soon as it reaches the target genome
    unsigned long long *wigKeys; // NULL: normal liftover.  One order-preserving key per target base (wigKey below)
    const int64_t *wigValOff;    // per interval: >= 0 index of its first per-base value in wigVals; < 0: ~index of its single value
    const double *wigVals;
    // scratch: lists live in dynamic shared memory unless gscratch != NULL
    int32_t listCap;   // fragments per list (two lists per warp)
    int32_t frameCap;  // work-pool frames per warp
    uint8_t *gscratch;
    uint64_t gscratchPerWarp;
};

struct GenomeTab { // one per genome, device resident
    const TopRec *top;
    const BotCore *bot;
    const int64_t *child;       // nc columns of numBot entries
    const int32_t *childGenome; // nc entries
    const uint32_t *topBucket, *botBucket;
    int64_t numTop, numBot;
    int32_t nc, parent, slot, topShift, botShift;
    uint8_t inScope, isTarget, pad[2];
    const int64_t *seqStart; // numSeq + 1 entries
    int32_t numSeq;
    int32_t nameRank;        // rank of the genome name in byte order (ColumnIterator::SequenceLess, halColumnIterator.h:45-50)
};


struct ColRowRec { // 16 B
    int64_t pos;    // forward genome coordinate
    int32_t seq;    // sequence index within the genome
    int16_t genome;
    uint8_t rev;
    uint8_t pad;
};


// Wiggle values are kept as keys whose unsigned order is the order of the doubles (sign bit flipped for v >= 0, all bits
// for v < 0).  WIG_UNSET (== the key of -0.0) marks a base nothing was written to; WiggleLiftover::mapFragments takes
// max(value, WiggleTiles::get()) where get() is 0.0 until the base is set (liftover/impl/halWiggleLiftover.cpp:150-153,
// liftover/inc/halWiggleTiles.h:96-107), so the first write to an unset base stores max(v, +0.0).
static const unsigned long long WIG_UNSET = 0x7fffffffffffffffull;
static const unsigned long long WIG_ZERO = 0x8000000000000000ull;

enum : uint32_t { ST_OK = 0, ST_SCRATCH_OVERFLOW = 1, ST_POOL_FULL = 2, ST_BAD_INPUT = 3 };

} // namespace halgpu
