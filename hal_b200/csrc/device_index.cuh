// Device-resident index layout (HBM) and the per-launch parameter blocks.
//
// The file's AoS records (top 40 B, bottom 8*(2+nc)+roundup8(nc) B) are re-packed
// once at staging time into sector-friendly arrays; the file format itself is untouched.
//   TopRec   32 B  == one DRAM sector: start | parent link | bottomParseIndex | nextParalogyIndex
//   BotCore  16 B : start | topParseIndex          (what every walk needs from a bottom record)
//   childEnc  8 B : child link, one array per child slot (a walk only ever follows ONE slot per level,
//                   so the other slots' columns are never fetched -- AoS bottoms of an 8-child genome are 88 B)
// A link (parent link of a top record, child link of a bottom record) is -1 for "none", else
//   bit 0      reversed
//   bits 1-32  index of the homologous segment in the other genome
//   bits 33-62 collinear run, in bases, capped at LINK_RUN_CAP: counted from the START of this segment, how far the
//              vertical map stays ONE affine map -- the following segments all have links, the same orientation and
//              indices advancing by +1 (forward) / -1 (reversed).  For child links the run is additionally 0 when the
//              landing top segment carries a paralogy ring and stops before the first following segment whose landing
//              top does (mapSelf fans out there, api/impl/halSegmentMapper.cpp:263-288).  Computed once at staging
//              (stage_kernels.cuh: linkRunKernel); the file format has no such field.
// Beside every link array sits an xlate array (topX / childX): the constant c of the affine map of that link's run,
// position q -> q + c (forward) or c - q (reversed) in the other genome (stage_kernels.cuh: linkXlateKernel).
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

static const int64_t LINK_RUN_CAP = (1ll << 30) - 1;
__host__ __device__ inline int64_t linkIdx(int64_t e) { return (e >> 1) & 0xffffffffll; }
__host__ __device__ inline bool linkRev(int64_t e) { return (e & 1) != 0; }
__host__ __device__ inline int64_t linkRun(int64_t e) { return e >> 33; } // e >= 0
__host__ __device__ inline int64_t makeLink(int64_t idx, bool rev, int64_t run) {
    return (run << 33) | (idx << 1) | (rev ? 1 : 0);
}

struct alignas(32) TopRec {
    int64_t start;
    int64_t parentEnc; // -1: no parent; else a link (see above) to the parent genome's bottom segment
    int64_t botParse;  // bottomParseIndex (-1 for leaves)
    int64_t nextPara;  // nextParalogyIndex (-1: none)
};

struct alignas(16) BotCore {
    int64_t start;
    int64_t topParse; // topParseIndex (-1 for the root)
};

// fastLiftKernel's index of one vertical transition: per position bucket (the buckets of topBucket / botBucket) everything
// a hop needs from the segment that holds the bucket's first base, in ONE 32-byte sector -- so a hop whose interval lies in
// that segment's collinear run costs a single memory round trip.
struct alignas(32) FastRec {
    int64_t start; // start of that segment
    int64_t link;  // its vertical link (reversed flag, index, run)
    int64_t xlate; // the run's translation constant
    int64_t idx;   // the segment's index (where the exact search resumes when the run does not cover the interval)
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
    // position -> segment index tables of this genome (fastLiftKernel re-locates a fragment after every hop)
    const uint32_t *topBucket, *botBucket;
    int32_t topShift, botShift;
    // translation constants of the transition p -> p+1, one per segment of the array the hop leaves (tops for an up
    // transition, this slot's bottoms for a down transition): a position q inside the segment's collinear run maps to
    // q + xlate (forward link) or xlate - q (reversed link) without touching the other genome's records
    const int64_t *xlate;
    const FastRec *fast; // the transition's bucket-indexed hop records (same buckets as topBucket / botBucket)
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
    const uint32_t *work;    // optional: work[w] = interval id (retry subsets, the complex list); NULL: identity
    const unsigned long long *work64; // optional, instead of work: the sorted batch, interval id in the low 32 bits
    const unsigned long long *nDev; // optional: the number of work items lives on the device (complex list length)
    const int64_t *gs, *ge;  // genome-global inclusive
    const uint8_t *strand;   // may be NULL
    // outputs
    unsigned long long *outLoc; // per interval: (first record in pool << LOC_COUNT_BITS) | number of records; 0 until written
    uint32_t *status;        // per interval: ST_* (zeroed == ST_OK by the engine; only failures are written)
    unsigned long long *failCount; // 5 counters indexed by ST_*: intervals that ended in that state (ST_OK is not counted)
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
    int32_t seedTile;  // 1: stage every interval's run of source top records into shared memory with one cp.async.bulk (TMA)
    int32_t pad1;
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

enum : uint32_t { ST_OK = 0, ST_SCRATCH_OVERFLOW = 1, ST_POOL_FULL = 2, ST_BAD_INPUT = 3, ST_REDO_EXACT = 4 /* fused walk (LIFT_FUSE) only */ };
#define HG_LOC_COUNT_BITS 25 // an interval yields at most 2^24 records (engine.cu: listCap ladder)

// fastLiftKernel (liftover_kernel.cuh): one LANE per interval.  An interval whose whole source range maps through every
// genome of the path as ONE affine piece without meeting a paralogy ring yields exactly one output line; everything
// else is appended to the complex list and walked by liftoverKernel (one warp per interval).
struct FastParams {
    const PathStep *steps;
    int32_t P;
    int32_t srcIsTop;
    int64_t srcLen;
    const int64_t *tgtSeqStart;
    int32_t tgtNumSeq;
    int32_t tileGrab;                     // a warp takes this many tiles per atomicAdd on tileCursor (>= 1)
    int64_t n;
    const int64_t *gs, *ge;               // input order
    const uint8_t *strand;                // may be NULL
    const unsigned long long *sortedGs;   // optional: gs in visiting order ...
    const unsigned long long *sortedVal;  // ... with (interval id | min(length, 2^32 - 1) << 32)
    const unsigned long long *sortedKey;  // or (source genomes shorter than 2^32): source start << 32 | interval id, in visiting order
    unsigned long long *tileCursor;       // next tile of 32 work items
    halgpu_lift_rec *pool;                // a finished interval's one line goes to pool[interval id] (outLoc is preset to that)
    uint32_t *complexList;                // interval ids left to liftoverKernel
    unsigned long long *complexCount;
};

} // namespace halgpu
