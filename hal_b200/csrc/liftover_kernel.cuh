// liftover_kernel.cuh -- the per-interval liftover walk, one warp per source interval.
//
// Replaces, for one BED interval, the reference call chain
//   BlockLiftover::liftInterval            liftover/impl/halBlockLiftover.cpp:46-113
//     SegmentIterator::toSite / toRight    api/impl/halSegmentIterator.cpp:240-299, 208-238      (seeds)
//     halMapSegment -> mapSource           api/impl/halSegmentMapper.cpp:578-670
//       mapUp / mapDown / mapSelf          api/impl/halSegmentMapper.cpp:25-80, 128-186, 263-288
//       insertAndBreakOverlaps             api/impl/halSegmentMapper.cpp:475-520                  (refinement)
//     BlockMapper::extractSegment          liftover/impl/halBlockMapper.cpp:331-394               (merge)
//   + stable sort by source start          liftover/impl/halLiftover.cpp:90
//
// Phase 1 (map): every lane owns one mapped fragment and advances it one genome per iteration along the
//   src -> mrca -> tgt path.  Lanes that run dry take the next source segment of the interval (a "seed");
//   when a fragment fans out (it straddles a parse boundary, or lands on a paralogy ring) the extra piece
//   is pushed as a 48-byte frame to a per-warp work pool that any idle lane pops -- pushes, pops and
//   emits are compacted with __ballot_sync/popc, so list order is deterministic.  With collinear data
//   all 32 lanes stay in lock step on adjacent records: every hop is one coalesced 1 KB read.
// Phase 2 (reduce): the warp sorts the interval's fragments by target, cuts overlapping target extents
//   to their common refinement, merges collinear neighbours into output lines and writes 32-byte
//   records to the output pool (one atomicAdd per interval).
//
// The same source compiles for the host under HALGPU_SIMT_EMUL (tests/simt: 32 threads emulate a warp)
// so that the control flow can be checked against the CPU oracle without a GPU.  That build is a test
// harness only; the product library contains the sm_100a build alone.
#pragma once
#include "device_index.cuh"

namespace halgpu {

#define HG_FULL 0xffffffffu

struct Frag { // 32 B
    int64_t sLo, tLo, len;
    int64_t meta; // bit0 sRev, bit1 tRev, bit2 kindTop, bits 3.. : segment index (phase 1) / sequence id (phase 2)
};

static_assert(sizeof(Frag) == sizeof(halgpu_frag) && sizeof(Frag) == sizeof(halgpu_lift_rec), "HALGPU_RAW_FRAGMENTS hands Frag records out as halgpu_frag");

struct Frame { // 48 B
    int64_t sLo, tLo, len, meta;
    int64_t aux;  // PARSE: index of the next segment of the other array; RING: first ring member
    int32_t p;    // path position
    int32_t type; // 0 PARSE, 1 RING
};

__device__ __forceinline__ int64_t ldS(const int64_t *p) {
    return (int64_t)__ldg(reinterpret_cast<const long long *>(p));
}
__device__ __forceinline__ TopRec ldTop(const TopRec *p) {
    const longlong2 *q = reinterpret_cast<const longlong2 *>(p);
    longlong2 a = __ldg(q), b = __ldg(q + 1);
    TopRec r;
    r.start = a.x; r.parentEnc = a.y; r.botParse = b.x; r.nextPara = b.y;
    return r;
}
__device__ __forceinline__ BotCore ldBot(const BotCore *p) {
    longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(p));
    BotCore r;
    r.start = a.x; r.topParse = a.y;
    return r;
}
__device__ __forceinline__ int64_t topStart(const TopRec *t, int64_t i) { return ldS(&t[i].start); }
__device__ __forceinline__ int64_t botStart(const BotCore *b, int64_t i) { return ldS(&b[i].start); }

// largest i in [i0, N) with start(i) <= pos, given start(i0) <= pos: gallop then bisect
template <bool TOP>
__device__ __forceinline__ int64_t searchFrom(const void *arr, int64_t i0, int64_t N, int64_t pos) {
    auto st = [&](int64_t i) -> int64_t {
        return TOP ? topStart(static_cast<const TopRec *>(arr), i) : botStart(static_cast<const BotCore *>(arr), i);
    };
    int64_t lo = i0, step = 1, hi;
    while (true) {
        hi = lo + step;
        if (hi >= N) { hi = N; break; }
        if (st(hi) <= pos) { lo = hi; step <<= 1; } else break;
    }
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (st(mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// the part of a fragment whose target extent is [a,b]  (sub-range rule: replaces the parse-back-and-
// difference trick of halSegmentMapper.cpp:52-62,157-167 and MappedSegment::slice)
__device__ __forceinline__ void subRange(int64_t &sLo, int64_t &tLo, int64_t &len, bool sRev, bool tRev, int64_t a,
                                         int64_t b) {
    const int64_t u = tRev ? (tLo + len - 1 - b) : (a - tLo);
    const int64_t m = b - a + 1;
    sLo = sRev ? (sLo + len - u - m) : (sLo + u);
    tLo = a;
    len = m;
}

__device__ __forceinline__ unsigned long long wigKey(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double wigUnkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// value = max(v, what the base holds); an unset base holds +0.0 (see WIG_UNSET)
__device__ __forceinline__ void wigRaise(unsigned long long *p, double v) {
    const unsigned long long k = wigKey(v);
    if (k > WIG_UNSET) {
        atomicMax(p, k);
    } else if (atomicMax(p, k) == WIG_UNSET) {
        atomicMax(p, WIG_ZERO);
    }
}

__device__ __forceinline__ int lanePrefix(unsigned mask, int lane) { return __popc(mask & ((1u << lane) - 1u)); }

// index of the sequence containing genome position pos (replaces Genome::getSequenceBySite,
// api/mmap_impl/mmapGenomeSiteMap.cpp:99-113)
__device__ __forceinline__ int seqOf(const int64_t *seqStart, int nseq, int64_t pos) {
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ldS(&seqStart[mid]) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// key order of MappedSegmentLess (api/impl/halMappedSegment.cpp:36-43): target (lo, hi) then source (lo, hi);
// hi is implied by len.  Flags break the (impossible in a valid HAL) exact-coordinate tie deterministically.
__device__ __forceinline__ bool fragLess(const Frag &a, const Frag &b) {
    if (a.tLo != b.tLo) return a.tLo < b.tLo;
    if (a.len != b.len) return a.len < b.len;
    if (a.sLo != b.sLo) return a.sLo < b.sLo;
    return (a.meta & 3) < (b.meta & 3);
}
__device__ __forceinline__ bool fragSameCoords(const Frag &a, const Frag &b) {
    return a.tLo == b.tLo && a.len == b.len && a.sLo == b.sLo;
}

// MappedSegment::canMergeRightWith (api/impl/halMappedSegment.cpp:109-161) + the same-sequence test of
// BlockMapper::extractSegment (liftover/impl/halBlockMapper.cpp:364-371); cut sets handled by the caller
__device__ __forceinline__ bool canMergeRight(const Frag &p, const Frag &q) {
    if (((p.meta ^ q.meta) & 3) != 0) return false;      // same target strand, same source strand
    if ((p.meta >> 3) != (q.meta >> 3)) return false;    // same target sequence
    if (q.tLo - (p.tLo + p.len - 1) != 1) return false;
    const bool sRev = p.meta & 1, tRev = (p.meta >> 1) & 1;
    if (sRev == tRev) return q.sLo - (p.sLo + p.len - 1) == 1;
    return p.sLo - (q.sLo + q.len - 1) == 1;
}

// stable rank sort of src[0..m) into dst by fragLess (ties keep list order).  The rank of an element is the number of
// elements before it in that order; nearly all pairs are decided by the target start alone, so the inner loop reads just
// that field (one shared-memory broadcast) and falls back to the full comparison only on equal starts.
__device__ __forceinline__ void warpRankSort(const Frag *src, Frag *dst, int m, int lane) {
    for (int i = lane; i < m; i += 32) {
        const Frag me = src[i];
        const int64_t myT = me.tLo;
        int r = 0;
        for (int j = 0; j < m; ++j) {
            const int64_t ot = src[j].tLo;
            r += ot < myT ? 1 : 0;
            // (j == i never counts, and testing for it keeps the element's comparison with ITSELF -- one lane of the warp in every
            // iteration -- off the full-comparison path: round 2's profile had 27 % of the walk's instructions there)
            if (ot == myT && j != i) {
                const Frag o = src[j];
                r += (fragLess(o, me) || (!fragLess(me, o) && j < i)) ? 1 : 0;
            }
        }
        dst[r] = me;
    }
    __syncwarp();
}

// per-interval bookkeeping: only failures touch the status array and the failure counters (the engine zeroes both)
__device__ __forceinline__ void liftFail(const LiftParams &P, uint32_t item, uint32_t st) {
    P.status[item] = st;
    P.outLoc[item] = 0ull;
    atomicAdd(P.failCount + st, 1ull);
}
__device__ __forceinline__ void liftDone(const LiftParams &P, uint32_t item, unsigned long long base, uint32_t count) {
    P.status[item] = ST_OK;
    P.outLoc[item] = (base << HG_LOC_COUNT_BITS) | (unsigned long long)count;
}

struct WarpScratch {
    Frag *listA, *listB;
    Frame *frames;
    // optional seed tile (LiftParams::seedTile): the interval's run of source top records, staged into shared memory by one
    // cp.async.bulk (TMA, 1-D) per interval and completed through an mbarrier
    TopRec *tile;
    uint32_t tileBar, tilePhase;
};

#define HG_SEED_TILE 40 // records per tile: covers a 1.2 kb interval of 32-bp segments plus the successor record

#if !defined(HALGPU_SIMT_EMUL)
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tileBarInit(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one lane: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void tileLoad(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tileWait(uint32_t bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    }
}
#endif

// index of the segment of genome st that holds position pos: bucket table + gallop
template <bool TOP> __device__ __forceinline__ int64_t locateSeg(const PathStep &st, int64_t pos) {
    if (TOP) return searchFrom<true>(st.top, (int64_t)__ldg(&st.topBucket[pos >> st.topShift]), st.numTop, pos);
    return searchFrom<false>(st.bot, (int64_t)__ldg(&st.botBucket[pos >> st.botShift]), st.numBot, pos);
}

// MODE: the wiggle mode, the coalescence-limit path and the raw-fragment mode are separate instantiations so that the
// default BED path's code and register allocation are untouched by them
enum : int { LIFT_BED = 0, LIFT_WIG = 1, LIFT_COAL = 2, LIFT_RAW = 4, LIFT_RAW_COAL = 6, LIFT_TILE = 8, LIFT_FUSE = 16 }; // COAL, RAW, TILE and FUSE are bits
template <int MODE>
__device__ __forceinline__ void liftOneInterval(const LiftParams &P, WarpScratch &ws, uint32_t item, int lane) {
    constexpr bool WIG = MODE == LIFT_WIG, COAL = (MODE & LIFT_COAL) != 0, RAW = (MODE & LIFT_RAW) != 0, TILE = (MODE & LIFT_TILE) != 0;
    // FUSE: fragments follow whole collinear runs of the vertical links (device_index.cuh) instead of being cut at every segment
    // boundary.  A fused fragment is the union of reference fragments that are images of one affine map and adjacent on both
    // sides; if the interval's fused fragments are pairwise disjoint in the target (and none crosses a target sequence
    // boundary) so are the reference's, extractSegment merges exactly the chains of neighbours that merge here, and the output
    // lines are identical.  Phase 2 checks that; an interval that fails the check is flagged ST_REDO_EXACT and re-walked piece by
    // piece by the plain instantiation.
    constexpr bool FUSE = (MODE & LIFT_FUSE) != 0;
    const int64_t gs = ldS(&P.gs[item]), ge = ldS(&P.ge[item]);
    const uint8_t bedStrand = P.strand ? P.strand[item] : (uint8_t)'+';
    const bool flip = bedStrand == '-';
    if (gs < 0 || ge < gs || ge >= P.srcLen) { // reported by the engine as an error; nothing is read for this item
        if (lane == 0) liftFail(P, item, ST_BAD_INPUT);
        return;
    }
    const PathStep *steps = P.steps;
    const int np = P.P;
    const int listCap = P.listCap, frameCap = P.frameCap;
    Frag *listA = ws.listA, *listB = ws.listB;
    Frame *frames = ws.frames;

    // ---- seeds: first / last source segment overlapping [gs, ge] (toSite + toRight) ----
    const void *srcArr = P.srcIsTop ? (const void *)steps[0].top : (const void *)steps[0].bot;
    int64_t mySeg = 0;
    if (lane < 2) {
        const int64_t pos = lane == 0 ? gs : ge;
        const int64_t i0 = (int64_t)__ldg(&P.srcBucket[pos >> P.srcShift]);
        mySeg = P.srcIsTop ? searchFrom<true>(srcArr, i0, P.srcN, pos) : searchFrom<false>(srcArr, i0, P.srcN, pos);
    }
    int64_t nextSeg = __shfl_sync(HG_FULL, mySeg, 0);
    const int64_t lastSeg = __shfl_sync(HG_FULL, mySeg, 1);
    // seed tile: records [tileFirst, tileFirst + tileN) of the source top array in shared memory
    const TopRec *tile = nullptr;
    int64_t tileFirst = 0;
    int tileN = 0;
#if !defined(HALGPU_SIMT_EMUL)
    if (TILE && ws.tile != nullptr && P.srcIsTop) {
        int64_t cnt = lastSeg - nextSeg + 2;
        if (cnt > HG_SEED_TILE) cnt = HG_SEED_TILE;
        if (lane == 0) tileLoad(smemAddr(ws.tile), &steps[0].top[nextSeg], (uint32_t)cnt * (uint32_t)sizeof(TopRec), ws.tileBar);
        tileWait(ws.tileBar, ws.tilePhase);
        ws.tilePhase ^= 1u;
        tile = ws.tile; tileFirst = nextSeg; tileN = (int)cnt;
    }
#endif

    // FUSE: seeds are whole runs of the source genome's first transition (parent links of its tops when the path starts
    // upward, child links of its bottoms when it starts downward from a genome without tops)
    const bool seedFuse = FUSE && np > 1 && ((P.srcIsTop != 0) == (steps[0].up != 0));
    int64_t seedCovered = -1; // FUSE: last base covered by the runs of the seeds handed out so far

    // ---- phase 1 ----
    bool valid = false;
    int64_t sLo = 0, tLo = 0, len = 0, idx = 0, cursor = -1, ringFirst = -1;
    bool sRev = false, tRev = false, kindTop = false;
    int p = 0;
    int listCount = 0, poolCount = 0;
    bool overflow = false;

    while (true) {
        // 1. idle lanes take work: pool frames first (LIFO), then fresh seeds
        const unsigned idle = __ballot_sync(HG_FULL, !valid);
        if (idle) {
            const int rank = lanePrefix(idle, lane);
            const int nIdle = __popc(idle);
            const int takePool = nIdle < poolCount ? nIdle : poolCount;
            if (!valid && rank < takePool) {
                const Frame f = frames[poolCount - 1 - rank];
                sLo = f.sLo; tLo = f.tLo; len = f.len;
                sRev = f.meta & 1; tRev = (f.meta >> 1) & 1; kindTop = (f.meta >> 2) & 1; idx = f.meta >> 3;
                p = f.p;
                if (f.type == 0) { cursor = f.aux; ringFirst = -1; } else { cursor = -1; ringFirst = f.aux; }
                valid = true;
            }
            poolCount -= takePool;
            int64_t avail = lastSeg - nextSeg + 1;
            if (avail < 0) avail = 0;
            const int want = nIdle - takePool;
            const int takeSeeds = (int64_t)want < avail ? want : (int)avail;
            if (seedFuse) {
                const bool cand = !valid && rank >= takePool && rank - takePool < takeSeeds;
                const int64_t seg = nextSeg + (rank - takePool);
                int64_t s0 = 0, runEnd = -1;
                if (cand) {
                    int64_t s1, link;
                    if (P.srcIsTop) {
                        const longlong2 h = __ldg(reinterpret_cast<const longlong2 *>(&steps[0].top[seg]));
                        s0 = h.x; link = h.y; s1 = topStart(steps[0].top, seg + 1);
                    } else {
                        s0 = botStart(steps[0].bot, seg); s1 = botStart(steps[0].bot, seg + 1); link = ldS(&steps[0].child[seg]);
                    }
                    runEnd = s1 - 1;
                    if (link >= 0 && s0 + linkRun(link) - 1 > runEnd) runEnd = s0 + linkRun(link) - 1;
                }
                // a segment opens a seed unless an earlier segment's run already covers it: running maximum of the run ends
                int64_t mx = runEnd;
                for (int d = 1; d < 32; d <<= 1) {
                    const int64_t v = __shfl_up_sync(HG_FULL, mx, d);
                    if (lane >= d && v > mx) mx = v;
                }
                int64_t before = __shfl_up_sync(HG_FULL, mx, 1);
                if (lane == 0 || before < seedCovered) before = seedCovered;
                const int64_t all = __shfl_sync(HG_FULL, mx, 31);
                if (all > seedCovered) seedCovered = all;
                if (cand && s0 > before) {
                    const int64_t a = gs > s0 ? gs : s0, b = ge < runEnd ? ge : runEnd;
                    sLo = a; tLo = a; len = b - a + 1;
                    sRev = flip; tRev = flip; kindTop = P.srcIsTop != 0; idx = seg;
                    p = 0; cursor = -1; ringFirst = -1;
                    valid = true;
                }
            } else if (!valid && rank >= takePool && rank - takePool < takeSeeds) {
                const int64_t seg = nextSeg + (rank - takePool);
                int64_t s0, s1;
                if (TILE && tile != nullptr && seg + 1 - tileFirst < tileN) {
                    s0 = tile[seg - tileFirst].start; s1 = tile[seg + 1 - tileFirst].start;
                } else if (P.srcIsTop) {
                    s0 = topStart(steps[0].top, seg); s1 = topStart(steps[0].top, seg + 1);
                } else {
                    s0 = botStart(steps[0].bot, seg); s1 = botStart(steps[0].bot, seg + 1);
                }
                const int64_t a = gs > s0 ? gs : s0, b = ge < s1 - 1 ? ge : s1 - 1;
                sLo = a; tLo = a; len = b - a + 1;
                sRev = flip; tRev = flip; kindTop = P.srcIsTop != 0; idx = seg;
                p = 0; cursor = -1; ringFirst = -1;
                valid = true;
            }
            nextSeg += takeSeeds;
            __syncwarp();
        }
        if (!__any_sync(HG_FULL, valid)) break;

        // 2. paralogy ring: queue the next member (mapSelf, halSegmentMapper.cpp:265-288:
        //    do { emit(cur); if (hasNext) toNext; } while (cur.hasNext && cur != first))
        Frame push;
        bool doPush = false;
        if (valid && ringFirst >= 0) {
            const TopRec *top = steps[p].top;
            const TopRec rm = ldTop(&top[idx]);
            const int64_t nx = rm.nextPara;
            if (nx >= 0) {
                const TopRec rn = ldTop(&top[nx]);
                if (rn.nextPara >= 0 && nx != ringFirst) {
                    const int64_t L = topStart(top, idx + 1) - rm.start;
                    const int64_t off = tLo - rm.start;
                    const bool fl = ((rn.parentEnc ^ rm.parentEnc) & 1) != 0; // toNextParalogy, halTopSegmentIterator.cpp:99-107
                    push.sLo = sLo;
                    push.tLo = fl ? rn.start + L - off - len : rn.start + off;
                    push.len = len;
                    push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)((tRev != fl) ? 1 : 0) << 1) | (1ll << 2) | (nx << 3);
                    push.aux = ringFirst;
                    push.p = p;
                    push.type = 1;
                    doPush = true;
                }
            }
            ringFirst = -1;
        }
        {
            const unsigned pm = __ballot_sync(HG_FULL, doPush);
            if (pm) {
                const int slot = poolCount + lanePrefix(pm, lane);
                if (doPush && slot < frameCap) frames[slot] = push;
                poolCount += __popc(pm);
                if (poolCount > frameCap) overflow = true;
            }
        }

        // 3. one genome hop; a fragment that straddles a parse boundary leaves its remainder in the pool
        doPush = false;
        bool landedDown = false;
        if (valid && p < np - 1) {
            const PathStep st = steps[p];
            const int64_t tHi = tLo + len - 1;
            if (COAL && (st.flags & STEP_PARA)) {
                // mapRecursiveParalogies (halSegmentMapper.cpp:525-576) at one genome between the MRCA and the limit
                if (!kindTop) { // mapSelf / mapUp of a bottom fragment: one top piece at a time (toParseUp), remainder to the pool
                    int64_t t = cursor;
                    if (t < 0) t = searchFrom<true>(st.top, ldBot(&st.bot[idx]).topParse, st.numTop, tLo);
                    const int64_t tEnd = topStart(st.top, t + 1) - 1;
                    if (tEnd < tHi) {
                        push.sLo = sLo; push.tLo = tLo; push.len = len;
                        subRange(push.sLo, push.tLo, push.len, sRev, tRev, tEnd + 1, tHi);
                        push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1) | (idx << 3);
                        push.aux = t + 1; push.p = p; push.type = 0;
                        doPush = true;
                        subRange(sLo, tLo, len, sRev, tRev, tLo, tEnd);
                    }
                    kindTop = true; idx = t; cursor = -1; // the fork happens in the next iteration, as a top piece
                } else {
                    const TopRec r = ldTop(&st.top[idx]);
                    if (!(st.flags & STEP_PARA_LAST) && r.parentEnc >= 0) {
                        // (b) mapUp(original, doDupes = true): a copy continues in the parent genome, one level further up
                        const int64_t L = topStart(st.top, idx + 1) - r.start;
                        const int64_t pi = linkIdx(r.parentEnc);
                        const bool fl = (r.parentEnc & 1) != 0;
                        const int64_t ps = botStart(steps[p + 1].bot, pi);
                        const int64_t off = tLo - r.start;
                        push.sLo = sLo; push.len = len;
                        push.tLo = fl ? ps + L - off - len : ps + off;
                        push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)((tRev != fl) ? 1 : 0) << 1) | (pi << 3);
                        push.aux = -1; push.p = p + 1; push.type = 0;
                        doPush = true;
                    }
                    // (a) mapSelf: this top and every member of its paralogy ring head back down to the MRCA
                    landedDown = r.nextPara >= 0;
                    p = st.jump;
                }
            } else if (FUSE && st.up) {
                // the top segment holding tLo, then as far as its parent link's run reaches
                int64_t t;
                if (kindTop) t = idx >= 0 ? idx : locateSeg<true>(st, tLo);
                else if (cursor >= 0) t = cursor;
                else if (idx >= 0) t = searchFrom<true>(st.top, ldBot(&st.bot[idx]).topParse, st.numTop, tLo);
                else t = locateSeg<true>(st, tLo);
                cursor = -1;
                const TopRec r = ldTop(&st.top[t]);
                const int64_t tNext = topStart(st.top, t + 1);
                int64_t covEnd = tNext - 1;
                if (r.parentEnc >= 0 && r.start + linkRun(r.parentEnc) - 1 > covEnd) covEnd = r.start + linkRun(r.parentEnc) - 1;
                if (covEnd < tHi) {
                    push.sLo = sLo; push.tLo = tLo; push.len = len;
                    subRange(push.sLo, push.tLo, push.len, sRev, tRev, covEnd + 1, tHi);
                    push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1) | (1ll << 2) | (int64_t)(~7ull); // a top piece, segment unknown (-1)
                    push.aux = -1; push.p = p; push.type = 0;
                    doPush = true;
                    subRange(sLo, tLo, len, sRev, tRev, tLo, covEnd);
                }
                if (r.parentEnc < 0) {
                    valid = false;
                } else {
                    const int64_t L = tNext - r.start;
                    const int64_t pi = linkIdx(r.parentEnc);
                    const bool fl = (r.parentEnc & 1) != 0;
                    const int64_t ps = botStart(steps[p + 1].bot, pi);
                    const int64_t off = tLo - r.start;
                    tLo = fl ? ps + L - off - len : ps + off;
                    tRev = tRev != fl;
                    kindTop = false; idx = (!fl || off + len <= L) ? pi : -1; ++p; // (a reversed run lands in earlier segments)
                }
            } else if (FUSE) {
                // the bottom segment holding tLo, then as far as its child link's run reaches (0 when the landing top has a ring)
                int64_t b;
                if (!kindTop) b = idx >= 0 ? idx : locateSeg<false>(st, tLo);
                else if (cursor >= 0) b = cursor;
                else if (idx >= 0) b = searchFrom<false>(st.bot, ldTop(&st.top[idx]).botParse, st.numBot, tLo);
                else b = locateSeg<false>(st, tLo);
                cursor = -1;
                const int64_t ce = ldS(&st.child[b]);
                const int64_t b0 = botStart(st.bot, b), bNext = botStart(st.bot, b + 1);
                int64_t covEnd = bNext - 1;
                if (ce >= 0 && b0 + linkRun(ce) - 1 > covEnd) covEnd = b0 + linkRun(ce) - 1;
                if (covEnd < tHi) {
                    push.sLo = sLo; push.tLo = tLo; push.len = len;
                    subRange(push.sLo, push.tLo, push.len, sRev, tRev, covEnd + 1, tHi);
                    push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1) | (int64_t)(~7ull); // a bottom piece, segment unknown (-1)
                    push.aux = -1; push.p = p; push.type = 0;
                    doPush = true;
                    subRange(sLo, tLo, len, sRev, tRev, tLo, covEnd);
                }
                if (ce < 0) {
                    valid = false;
                } else {
                    const int64_t L = bNext - b0;
                    const int64_t ci = linkIdx(ce);
                    const bool fl = (ce & 1) != 0;
                    int64_t cs;
                    if (P.dupes) {
                        const TopRec rc = ldTop(&steps[p + 1].top[ci]);
                        cs = rc.start;
                        landedDown = rc.nextPara >= 0; // (then the run is 0 and the piece lies inside this one segment)
                    } else {
                        cs = topStart(steps[p + 1].top, ci);
                    }
                    const int64_t off = tLo - b0;
                    tLo = fl ? cs + L - off - len : cs + off;
                    tRev = tRev != fl;
                    kindTop = true; idx = (!fl || off + len <= L) ? ci : -1; ++p;
                }
            } else if (st.up) {
                if (!kindTop) { // bottom fragment: cut at the top-segment boundary (toParseUp, halTopSegmentIterator.cpp:55-81)
                    int64_t t = cursor;
                    if (t < 0) t = searchFrom<true>(st.top, ldBot(&st.bot[idx]).topParse, st.numTop, tLo);
                    const int64_t tEnd = topStart(st.top, t + 1) - 1;
                    if (tEnd < tHi) {
                        push.sLo = sLo; push.tLo = tLo; push.len = len;
                        subRange(push.sLo, push.tLo, push.len, sRev, tRev, tEnd + 1, tHi);
                        push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1) | (idx << 3);
                        push.aux = t + 1; push.p = p; push.type = 0;
                        doPush = true;
                        subRange(sLo, tLo, len, sRev, tRev, tLo, tEnd);
                    }
                    kindTop = true; idx = t; cursor = -1;
                }
                TopRec r; // toParent, halBottomSegmentIterator.cpp:40-49
                int64_t rNext;
                if (TILE && p == 0 && tile != nullptr && idx >= tileFirst && idx + 1 - tileFirst < tileN) {
                    r = tile[idx - tileFirst]; rNext = tile[idx + 1 - tileFirst].start;
                } else {
                    r = ldTop(&st.top[idx]); rNext = r.parentEnc < 0 ? 0 : topStart(st.top, idx + 1);
                }
                if (r.parentEnc < 0) {
                    valid = false;
                } else if (P.upCanonicalOnly && linkIdx(ldS(&st.child[linkIdx(r.parentEnc)])) != idx) {
                    valid = false; // ColumnIterator noDupes: only the canonical paralog goes up (halColumnIterator.cpp:559-560)
                } else {
                    const int64_t L = rNext - r.start;
                    const int64_t pi = linkIdx(r.parentEnc);
                    const bool fl = (r.parentEnc & 1) != 0;
                    const int64_t ps = botStart(steps[p + 1].bot, pi);
                    const int64_t off = tLo - r.start;
                    tLo = fl ? ps + L - off - len : ps + off;
                    tRev = tRev != fl;
                    kindTop = false; idx = pi; ++p;
                }
            } else {
                if (kindTop) { // top fragment: cut at the bottom-segment boundary (toParseDown, halBottomSegmentIterator.cpp:51-76)
                    int64_t b = cursor;
                    if (b < 0) b = searchFrom<false>(st.bot, ldTop(&st.top[idx]).botParse, st.numBot, tLo);
                    const int64_t bEnd = botStart(st.bot, b + 1) - 1;
                    if (bEnd < tHi) {
                        push.sLo = sLo; push.tLo = tLo; push.len = len;
                        subRange(push.sLo, push.tLo, push.len, sRev, tRev, bEnd + 1, tHi);
                        push.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1) | (1ll << 2) | (idx << 3);
                        push.aux = b + 1; push.p = p; push.type = 0;
                        doPush = true;
                        subRange(sLo, tLo, len, sRev, tRev, tLo, bEnd);
                    }
                    kindTop = false; idx = b; cursor = -1;
                }
                const int64_t ce = ldS(&st.child[idx]); // toChild, halTopSegmentIterator.cpp:36-45
                if (ce < 0) {
                    valid = false;
                } else {
                    const int64_t b0 = botStart(st.bot, idx);
                    const int64_t L = botStart(st.bot, idx + 1) - b0;
                    const int64_t ci = linkIdx(ce);
                    const bool fl = (ce & 1) != 0;
                    int64_t cs;
                    if (P.dupes && !(COAL && (st.flags & STEP_NODUPES))) { // one 32-byte read gives the landing start AND tells whether a paralogy ring hangs here
                        const TopRec rc = ldTop(&steps[p + 1].top[ci]);
                        cs = rc.start;
                        landedDown = rc.nextPara >= 0;
                    } else {
                        cs = topStart(steps[p + 1].top, ci);
                    }
                    const int64_t off = tLo - b0;
                    tLo = fl ? cs + L - off - len : cs + off;
                    tRev = tRev != fl;
                    kindTop = true; idx = ci; ++p;
                }
            }
        }
        if (landedDown) ringFirst = idx; // mapSelf: only a top with a next paralog starts a ring walk
        {
            const unsigned pm = __ballot_sync(HG_FULL, doPush);
            if (pm) {
                const int slot = poolCount + lanePrefix(pm, lane);
                if (doPush && slot < frameCap) frames[slot] = push;
                poolCount += __popc(pm);
                if (poolCount > frameCap) overflow = true;
            }
        }

        // 4. fragments that reached the target genome join the interval's result list
        {
            const bool doEmit = valid && p == np - 1 && ringFirst < 0;
            const unsigned em = __ballot_sync(HG_FULL, doEmit);
            if (WIG && em) {
                // WiggleLiftover::mapFragments (liftover/impl/halWiggleLiftover.cpp:134-158): the warp takes the arrived
                // fragments one at a time and writes max(value of the source base, current) to every target base -- 32
                // consecutive bases per step, so the value reads and the atomics are coalesced.  max is idempotent and
                // commutative: duplicates, the set order and the merge step of the BED path do not matter here.
                const int64_t vo = ldS(&P.wigValOff[item]);
                unsigned rem = em;
                while (rem) {
                    const int from = __ffs((int)rem) - 1;
                    rem &= rem - 1;
                    const int64_t fs = __shfl_sync(HG_FULL, sLo, from), ft = __shfl_sync(HG_FULL, tLo, from);
                    const int64_t fl = __shfl_sync(HG_FULL, len, from);
                    const int fr = __shfl_sync(HG_FULL, (int)tRev, from);
                    for (int64_t k = lane; k < fl; k += 32) {
                        const int64_t u = fr ? fl - 1 - k : k; // offset of target base ft + k along the (forward) source piece
                        const double v = vo >= 0 ? P.wigVals[vo + (fs + u - gs)] : P.wigVals[~vo];
                        wigRaise(&P.wigKeys[ft + k], v);
                    }
                }
                if (doEmit) valid = false;
            } else if (em) {
                const int slot = listCount + lanePrefix(em, lane);
                if (doEmit) {
                    if (slot < listCap) {
                        Frag f;
                        f.sLo = sLo; f.tLo = tLo; f.len = len;
                        f.meta = (int64_t)(sRev ? 1 : 0) | ((int64_t)(tRev ? 1 : 0) << 1);
                        listA[slot] = f;
                    }
                    valid = false;
                }
                listCount += __popc(em);
                if (listCount > listCap) overflow = true;
            }
        }
        __syncwarp();
        if (overflow) break;
    }

    if (overflow) {
        if (lane == 0) liftFail(P, item, ST_SCRATCH_OVERFLOW);
        return;
    }

    // ---- phase 2 ----
    int m = listCount;
    if (m == 0) {
        if (lane == 0) liftDone(P, item, 0, 0);
        return;
    }
    if (RAW) {
        // halgpu_liftover(HALGPU_RAW_FRAGMENTS): hand the mapped fragments out as they are (halgpu_frag has Frag's layout);
        // refinement and merging then happen over MANY intervals at once on the caller's side (halSynteny lifts whole
        // chromosomes: insertAndBreakOverlaps / extractSegment become global there)
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.poolCursor, (unsigned long long)m);
        base = __shfl_sync(HG_FULL, base, 0);
        if (base + (unsigned long long)m > P.poolCap) {
            if (lane == 0) liftFail(P, item, ST_POOL_FULL);
            return;
        }
        Frag *dst = reinterpret_cast<Frag *>(P.pool) + base;
        for (int i = lane; i < m; i += 32) dst[i] = listA[i];
        if (lane == 0) liftDone(P, item, base, (uint32_t)m);
        return;
    }
    // target sequence of every fragment (MappedSegment::getSequence)
    bool redoExact = false; // FUSE: this interval needs the piece-by-piece walk
    if (P.tgtNumSeq > 1) {
        for (int i = lane; i < m; i += 32) {
            const int sq = seqOf(P.tgtSeqStart, P.tgtNumSeq, listA[i].tLo);
            listA[i].meta |= (int64_t)sq << 3;
            if (FUSE && listA[i].tLo + listA[i].len > ldS(&P.tgtSeqStart[sq + 1])) redoExact = true; // pieces would not merge across sequences
        }
    }
    __syncwarp();
    if (P.columnMerge) {
        // ColumnLiftover::liftInterval (liftover/impl/halColumnLiftover.cpp:21-92): the target bases homologous to the
        // interval, per (sequence, strand), as maximal position runs -- forward-strand runs first, then reverse.  That set
        // is the union of the mapped fragments' target extents, so: order by (strand, sequence, start) and sweep.
        Frag *src = listA, *dst = listB;
        for (int i = lane; i < m; i += 32) {
            const Frag me = src[i];
            const int64_t kme = ((me.meta >> 1) & 1) << 40 | (me.meta >> 3);
            int r = 0;
            for (int j = 0; j < m; ++j) {
                const Frag o = src[j];
                const int64_t ko = ((o.meta >> 1) & 1) << 40 | (o.meta >> 3);
                const bool less = ko != kme ? ko < kme : (o.tLo != me.tLo ? o.tLo < me.tLo : (o.len != me.len ? o.len < me.len : j < i));
                r += less ? 1 : 0;
            }
            dst[r] = me;
        }
        __syncwarp();
        int nl = 0;
        if (lane == 0) { // sweep; line k is kept in src[k]: tLo = start, len = length, sLo = fragments merged
            int64_t key = -1, lo = 0, hi = -1, cnt = 0;
            for (int i = 0; i < m; ++i) {
                const Frag f = dst[i];
                const int64_t k = ((f.meta >> 1) & 1) << 40 | (f.meta >> 3);
                if (cnt > 0 && k == key && f.tLo <= hi + 1) {
                    const int64_t e = f.tLo + f.len - 1;
                    if (e > hi) hi = e;
                    ++cnt;
                } else {
                    if (cnt > 0) { Frag l; l.tLo = lo; l.len = hi - lo + 1; l.sLo = cnt; l.meta = key; src[nl++] = l; }
                    key = k; lo = f.tLo; hi = f.tLo + f.len - 1; cnt = 1;
                }
            }
            if (cnt > 0) { Frag l; l.tLo = lo; l.len = hi - lo + 1; l.sLo = cnt; l.meta = key; src[nl++] = l; }
        }
        nl = __shfl_sync(HG_FULL, nl, 0);
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.poolCursor, (unsigned long long)nl);
        base = __shfl_sync(HG_FULL, base, 0);
        if (base + (unsigned long long)nl > P.poolCap) {
            if (lane == 0) liftFail(P, item, ST_POOL_FULL);
            return;
        }
        for (int r = lane; r < nl; r += 32) {
            const Frag l = src[r];
            const int seq = (int)(l.meta & 0xffffffffffll);
            const int64_t seqStart = P.tgtNumSeq > 1 ? ldS(&P.tgtSeqStart[seq]) : 0;
            halgpu_lift_rec o;
            o.start = l.tLo - seqStart;
            o.end = l.tLo + l.len - seqStart;
            o.src_start = -1; // "not available from posMap" (halColumnLiftover.cpp:70)
            o.tgt_seq = seq;
            o.strand = bedStrand == '.' ? '.' : ((l.meta >> 40) & 1 ? '-' : '+');
            o.src_strand = o.strand;
            o.n_frag = (uint16_t)(l.sLo > 65535 ? 65535 : l.sLo);
            P.pool[base + r] = o;
        }
        if (lane == 0) liftDone(P, item, base, (uint32_t)nl);
        return;
    }
    // one pass over neighbouring pairs: order of the MappedSegmentSet, its "identical or disjoint" invariant
    // (insertAndBreakOverlaps) and whether any equal-target-start class has more than one member
    bool unsorted = false, clash = false, classes = false, split = false;
    for (int i = lane; i + 1 < m; i += 32) {
        const Frag a = listA[i], b = listA[i + 1];
        unsorted |= fragLess(b, a);
        const bool same = a.tLo == b.tLo && a.len == b.len;
        // (COAL) a fragment and its own image mapped up and back down meet again in the MRCA: the reference drops such exact
        // repeats with list::unique (halSegmentMapper.cpp:573-574); here they take the refinement path, which de-duplicates
        clash |= !(same || b.tLo > a.tLo + a.len - 1) || (COAL && fragSameCoords(a, b));
        classes |= a.tLo == b.tLo;
        split |= !canMergeRight(a, b);
    }
    // collinear interval (the common case): already in set order, no overlaps, every neighbour merges -> one line
    const bool single = !__any_sync(HG_FULL, unsorted || clash || classes || split);
    Frag *cur = listA, *oth = listB;
    if (__any_sync(HG_FULL, unsorted)) {
        warpRankSort(cur, oth, m, lane);
        Frag *t = cur; cur = oth; oth = t;
        clash = false; classes = false;
        for (int i = lane; i + 1 < m; i += 32) {
            const Frag a = cur[i], b = cur[i + 1];
            const bool same = a.tLo == b.tLo && a.len == b.len;
            clash |= !(same || b.tLo > a.tLo + a.len - 1) || (COAL && fragSameCoords(a, b));
            classes |= a.tLo == b.tLo;
        }
    }
    if (FUSE && __any_sync(HG_FULL, clash || classes || redoExact)) {
        // fused fragments that overlap in the target (paralogous source pieces hit the same target bases): the reference cuts
        // them against each other at its own piece boundaries -- walk this interval piece by piece instead
        if (lane == 0) liftFail(P, item, ST_REDO_EXACT);
        return;
    }
    bool refined = false;
    if (__any_sync(HG_FULL, clash)) {
        // cut every fragment at every other fragment's tLo and tHi+1 that falls strictly inside it
        int total = 0;
        for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            int pieces = 0;
            if (i < m) {
                const Frag f = cur[i];
                const int64_t hi = f.tLo + f.len - 1;
                int64_t at = f.tLo;
                pieces = 1;
                while (true) { // next breakpoint in (at, hi]
                    int64_t best = hi + 1;
                    for (int j = 0; j < m; ++j) {
                        const int64_t v0 = cur[j].tLo, v1 = v0 + cur[j].len;
                        if (v0 > at && v0 < best) best = v0;
                        if (v1 > at && v1 < best) best = v1;
                    }
                    if (best > hi) break;
                    ++pieces;
                    at = best;
                }
            }
            // exclusive scan of pieces over the 32 lanes
            int incl = pieces;
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(HG_FULL, incl, d);
                if (lane >= d) incl += v;
            }
            const int first = total + incl - pieces;
            total += __shfl_sync(HG_FULL, incl, 31);
            if (i < m && first + pieces <= listCap) {
                const Frag f = cur[i];
                const bool sR = f.meta & 1, tR = (f.meta >> 1) & 1;
                const int64_t hi = f.tLo + f.len - 1;
                int64_t at = f.tLo;
                int k = 0;
                while (true) {
                    int64_t best = hi + 1;
                    for (int j = 0; j < m; ++j) {
                        const int64_t v0 = cur[j].tLo, v1 = v0 + cur[j].len;
                        if (v0 > at && v0 < best) best = v0;
                        if (v1 > at && v1 < best) best = v1;
                    }
                    Frag g = f;
                    subRange(g.sLo, g.tLo, g.len, sR, tR, at, best - 1);
                    oth[first + k] = g;
                    ++k;
                    if (best > hi) break;
                    at = best;
                }
            }
        }
        __syncwarp();
        if (total > listCap) {
            if (lane == 0) liftFail(P, item, ST_SCRATCH_OVERFLOW);
            return;
        }
        m = total;
        // sort the pieces, then drop exact duplicates (set semantics: equal keys are one element)
        warpRankSort(oth, cur, m, lane);
        int kept = 0;
        for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            const bool keep = i < m && (i == 0 || !fragSameCoords(cur[i - 1], cur[i]));
            const unsigned km = __ballot_sync(HG_FULL, keep);
            if (keep) oth[kept + lanePrefix(km, lane)] = cur[i];
            kept += __popc(km);
        }
        __syncwarp();
        m = kept;
        Frag *t = cur; cur = oth; oth = t;
        refined = true;
    }

    // ---- merge into output lines (BlockMapper::extractSegment) ----
    // scratch in the free list: run heads/tails, then (general path only) cut points, flags, classes
    // layout (<= 32 bytes per fragment): runHead 4m | runTail 4m | qcut 8m | v1 4m | v2 4m | dead m | runOf 4m
    int32_t *runHead = reinterpret_cast<int32_t *>(oth);
    int32_t *runTail = runHead + m;
    int64_t *qcut = reinterpret_cast<int64_t *>(runTail + m);
    int32_t *v1 = reinterpret_cast<int32_t *>(qcut + m);
    int32_t *v2 = v1 + m;
    uint8_t *dead = reinterpret_cast<uint8_t *>(v2 + m);
    int32_t *runOf = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(oth) + (((size_t)25 * (size_t)m + 3) & ~(size_t)3));
    if (refined) { // the refinement changed the list: recompute
        classes = false;
        for (int i = lane; i + 1 < m; i += 32) classes |= cur[i].tLo == cur[i + 1].tLo;
    }
    int nLines = 0;
    if (single) {
        if (lane == 0) { runHead[0] = 0; runTail[0] = m - 1; }
        if (P.pslPool) for (int i = lane; i < m; i += 32) runOf[i] = 0;
        nLines = 1;
        __syncwarp();
    } else if (!__any_sync(HG_FULL, classes)) {
        // every equal-target-start class has one member: a run is a maximal chain of mergeable neighbours
        bool prevMerges = false; // does the last fragment of the previous 32 merge into this chunk's first?
        for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            const bool mr = i + 1 < m && canMergeRight(cur[i], cur[i + 1]);
            const unsigned mm = __ballot_sync(HG_FULL, mr);
            const bool mergesFromLeft = lane == 0 ? prevMerges : ((mm >> (lane - 1)) & 1u) != 0;
            const bool head = i < m && !mergesFromLeft;
            const bool tail = i < m && !mr;
            prevMerges = (mm >> 31) != 0;
            const unsigned hm = __ballot_sync(HG_FULL, head);
            if (head) runHead[nLines + lanePrefix(hm, lane)] = i;
            // a tail closes the run opened by the latest head at or before it
            const int headsUpToMe = __popc(hm & ((2u << lane) - 1u));
            if (tail) runTail[nLines + headsUpToMe - 1] = i;
            if (i < m) runOf[i] = nLines + headsUpToMe - 1;
            nLines += __popc(hm);
        }
        __syncwarp();
    } else {
        // general case, exact sequential restatement by one lane (rare: paralogous source pieces in one interval)
        if (lane == 0) {
            for (int i = 0; i < m; ++i) dead[i] = 0;
            int nq = 0;
            for (int x = 0; x < m; ++x) {
                if (dead[x]) continue;
                int n1 = 0, n2 = 0, tailIdx = x;
                runOf[x] = nLines;
                v1[n1++] = x;
                int nx = x + 1;
                while (nx < m && dead[nx]) ++nx;
                while (nx < m && cur[nx].tLo == cur[v1[n1 - 1]].tLo) {
                    v1[n1++] = nx;
                    ++nx;
                    while (nx < m && dead[nx]) ++nx;
                }
                while (nx < m) {
                    n2 = 0;
                    while (nx < m && (n2 == 0 || cur[v2[n2 - 1]].tLo == cur[nx].tLo) && n2 < n1) {
                        v2[n2++] = nx;
                        ++nx;
                        while (nx < m && dead[nx]) ++nx;
                    }
                    bool can = n1 == n2;
                    for (int i = 0; i < n1 && can; ++i) {
                        const Frag a = cur[v1[i]], b = cur[v2[i]];
                        bool ok = (b.meta >> 3) == (cur[x].meta >> 3) && canMergeRight(a, b);
                        if (ok) {
                            const int64_t cut = a.tLo + a.len - 1;
                            for (int q = 0; q < nq; ++q) ok &= qcut[q] != cut;
                        }
                        can = ok;
                    }
                    if (!can) break;
                    tailIdx = v2[0];
                    dead[v2[0]] = 1;
                    runOf[v2[0]] = nLines;
                    for (int i = 0; i < n2; ++i) v1[i] = v2[i];
                    n1 = n2;
                }
                if (n1 > 1) qcut[nq++] = cur[tailIdx].tLo + cur[tailIdx].len - 1;
                runHead[nLines] = x;
                runTail[nLines] = tailIdx;
                ++nLines;
            }
        }
        nLines = __shfl_sync(HG_FULL, nLines, 0);
        __syncwarp();
    }

    // ---- emit: stable by source start (Liftover::visitLine, halLiftover.cpp:90) ----
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.poolCursor, (unsigned long long)nLines);
    base = __shfl_sync(HG_FULL, base, 0);
    if (base + (unsigned long long)nLines > P.poolCap) {
        if (lane == 0) liftFail(P, item, ST_POOL_FULL);
        return;
    }
    // source start of every line first (qcut's slots are free after the merge scan: 8 bytes per line), then rank on those
    int64_t *lineSrc = qcut;
    for (int r = lane; r < nLines; r += 32) {
        const int64_t a = cur[runHead[r]].sLo, b = cur[runTail[r]].sLo;
        lineSrc[r] = a < b ? a : b;
    }
    __syncwarp();
    for (int r = lane; r < nLines; r += 32) {
        const Frag h = cur[runHead[r]], t = cur[runTail[r]];
        const int64_t src = lineSrc[r];
        int rank = 0;
        for (int j = 0; j < nLines; ++j) {
            const int64_t sj = lineSrc[j];
            rank += (sj < src || (sj == src && j < r)) ? 1 : 0;
        }
        const int seq = (int)(h.meta >> 3);
        const int64_t seqStart = P.tgtNumSeq > 1 ? ldS(&P.tgtSeqStart[seq]) : 0;
        halgpu_lift_rec o;
        o.start = (h.tLo < t.tLo ? h.tLo : t.tLo) - seqStart;
        const int64_t hHi = h.tLo + h.len, tHi = t.tLo + t.len;
        o.end = (hHi > tHi ? hHi : tHi) - seqStart;
        o.src_start = src;
        o.tgt_seq = seq;
        if (bedStrand == '.') {
            o.strand = '.'; o.src_strand = '.';
        } else {
            o.strand = ((h.meta >> 1) & 1) ? '-' : '+';
            o.src_strand = (h.meta & 1) ? '-' : '+';
        }
        const int nf = runTail[r] - runHead[r] + 1;
        o.n_frag = (uint16_t)(nf > 65535 ? 65535 : nf);
        P.pool[base + rank] = o;
        if (P.pslPool) v2[r] = rank; // v1/v2 are free after the merge scan; qcut/v1 may alias runOf's neighbours, v2 does not
    }
    if (P.pslPool) {
        // PSL base counts of every line (BlockLiftover::readPSLInfo, liftover/impl/halBlockLiftover.cpp:115-162):
        // walk each fragment's source and target strings in traversal order (reverse complement when reversed) and
        // classify every base pair: equal & upper-case -> match, equal & lower-case (masked) -> repMatch,
        // different with target n/N -> nCount, else mismatch.  Codes: bit 3 = upper case, low bits a,c,g,t,n = 0..4.
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            const Frag f = cur[i];
            const bool sR = f.meta & 1, tR = (f.meta >> 1) & 1;
            unsigned cnt[4] = {0u, 0u, 0u, 0u};
            for (int64_t j = 0; j < f.len; ++j) {
                const int64_t sp = sR ? f.sLo + f.len - 1 - j : f.sLo + j;
                const int64_t tp = tR ? f.tLo + f.len - 1 - j : f.tLo + j;
                unsigned sc = P.srcDna[sp >> 1], tc = P.tgtDna[tp >> 1];
                sc = (sp & 1) ? (sc & 0xFu) : (sc >> 4);
                tc = (tp & 1) ? (tc & 0xFu) : (tc >> 4);
                if (sR && (sc & 7u) < 4u) sc = (sc & 8u) | (3u - (sc & 7u));
                if (tR && (tc & 7u) < 4u) tc = (tc & 8u) | (3u - (tc & 7u));
                if (sc == tc) cnt[(sc & 8u) ? 0 : 2]++;
                else if ((tc & 7u) == 4u) cnt[3]++;
                else cnt[1]++;
            }
            unsigned *dst = P.pslPool + 4ull * (base + (unsigned long long)v2[runOf[i]]);
            for (int c = 0; c < 4; ++c)
                if (cnt[c]) atomicAdd(dst + c, cnt[c]);
        }
    }
    if (lane == 0) liftDone(P, item, base, (uint32_t)nLines);
}

// bytes of scratch one warp needs for the given capacities
__host__ __device__ inline uint64_t liftScratchBytes(int listCap, int frameCap) {
    return (uint64_t)listCap * 2u * sizeof(Frag) + (uint64_t)frameCap * sizeof(Frame);
}
// shared memory per warp of the optional seed tile (records + its mbarrier, 16-byte aligned)
__host__ __device__ inline uint64_t seedTileBytes() { return (uint64_t)HG_SEED_TILE * sizeof(TopRec) + 16u; }

#if !defined(HALGPU_SIMT_EMUL)
extern __shared__ __align__(16) uint8_t hg_dyn_smem[];
#endif

template <int MODE>
__global__ void __launch_bounds__(128, 8) liftoverKernel(const LiftParams P) {
#if defined(HALGPU_SIMT_EMUL)
    uint8_t *hg_dyn_smem = simt::dynamicSmem();
#endif
    const int lane = threadIdx.x & 31;
    const int warpInBlock = threadIdx.x >> 5;
    const int warpsPerBlock = blockDim.x >> 5;
    const uint64_t per = liftScratchBytes(P.listCap, P.frameCap);
    const int64_t gwarp = (int64_t)blockIdx.x * warpsPerBlock + warpInBlock;
    const int64_t nwarps = (int64_t)gridDim.x * warpsPerBlock;
    uint8_t *basePtr = P.gscratch ? P.gscratch + (uint64_t)gwarp * P.gscratchPerWarp : hg_dyn_smem + (uint64_t)warpInBlock * per;
    WarpScratch ws;
    ws.listA = reinterpret_cast<Frag *>(basePtr);
    ws.listB = ws.listA + P.listCap;
    ws.frames = reinterpret_cast<Frame *>(ws.listB + P.listCap);
    ws.tile = nullptr; ws.tileBar = 0; ws.tilePhase = 0;
#if !defined(HALGPU_SIMT_EMUL)
    if ((MODE & LIFT_TILE) != 0 && P.seedTile) { // behind the lists of all warps (the lists may live in global scratch; the tiles are always shared)
        uint8_t *t = hg_dyn_smem + (P.gscratch ? 0 : (uint64_t)warpsPerBlock * per) + (uint64_t)warpInBlock * seedTileBytes();
        ws.tile = reinterpret_cast<TopRec *>(t);
        ws.tileBar = smemAddr(t + (uint64_t)HG_SEED_TILE * sizeof(TopRec));
        if (lane == 0) tileBarInit(ws.tileBar);
        __syncwarp();
    }
#endif
    const int64_t n = P.nDev ? (int64_t)*P.nDev : P.n; // the complex list's length is only known on the device
    for (int64_t w = gwarp; w < n; w += nwarps) {
        const uint32_t item = P.work64 ? (uint32_t)P.work64[w] : (P.work ? __ldg(&P.work[w]) : (uint32_t)w);
        liftOneInterval<MODE>(P, ws, item, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fastLiftKernel -- one LANE per interval.
//
// BlockLiftover::liftInterval (liftover/impl/halBlockLiftover.cpp:46-113) cuts the interval at every segment boundary of
// every genome on the path (toRight seeds, toParseUp/Down pieces), maps the pieces one by one (halMapSegment) and then
// merges neighbours that are adjacent on both sides with equal strands (BlockMapper::extractSegment,
// liftover/impl/halBlockMapper.cpp:331-394).  When the WHOLE interval lies, in every genome of the path, inside one
// collinear run of the vertical links (the run field of a link, device_index.cuh) all those pieces are images of one affine
// map: they tile one target range, no paralogy ring adds anything (a ring at a landing top ends the run), so the set holds
// pairwise disjoint, pairwise mergeable neighbours and extractSegment returns exactly ONE line -- provided the target range
// stays inside one target sequence (:364-371).  That line is computed here with one record read per hop instead of one per
// piece.  Anything else (an unaligned piece, a rearrangement or a ring inside the interval, a sequence boundary, bad
// input) goes to the complex list and is walked piece by piece by liftoverKernel.
// ---------------------------------------------------------------------------------------------------------------
#define HG_FAST_MAX_PATH 24

__global__ void __launch_bounds__(256) fastLiftKernel(const FastParams P) {
    __shared__ PathStep sSteps[HG_FAST_MAX_PATH];
    for (int i = (int)threadIdx.x; i < P.P; i += (int)blockDim.x) sSteps[i] = P.steps[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // tiles of 32 work items, handed out P.tileGrab at a time by an atomic cursor (engine.cu: 4): the resident warps sweep the
    // sorted batch together
    const uint32_t nTiles = (uint32_t)((P.n + 31) >> 5);
    uint32_t tile = 0, grabEnd = 0;
    while (true) {
        if (tile >= grabEnd) {
            unsigned long long t0 = 0;
            if (lane == 0) t0 = atomicAdd(P.tileCursor, (unsigned long long)P.tileGrab);
            t0 = __shfl_sync(HG_FULL, t0, 0);
            if (t0 >= (unsigned long long)nTiles) break;
            tile = (uint32_t)t0;
            grabEnd = tile + (uint32_t)P.tileGrab;
        }
        if (tile >= nTiles) break;
        const int64_t w = (int64_t)tile * 32 + lane;
        ++tile;
        const bool have = w < P.n;
        uint32_t item = 0;
        int64_t gs = 0, ge = -1;
        if (have) {
            if (P.sortedKey) {
                const unsigned long long k = P.sortedKey[w];
                item = (uint32_t)k;
                gs = (int64_t)(k >> 32);
                ge = ldS(&P.ge[item]);
            } else if (P.sortedGs) {
                const unsigned long long v = P.sortedVal[w];
                item = (uint32_t)v;
                gs = (int64_t)P.sortedGs[w];
                const unsigned long long l32 = v >> 32;
                ge = l32 == 0xffffffffull ? ldS(&P.ge[item]) : gs + (int64_t)l32 - 1;
            } else {
                item = (uint32_t)w;
                gs = ldS(&P.gs[w]); ge = ldS(&P.ge[w]);
            }
        }
        bool ok = have && gs >= 0 && ge >= gs && ge < P.srcLen;
        int64_t tLo = gs;
        const int64_t len = ge - gs + 1;
        bool rev = false;
        if (ok) {
            const int np = P.P;
            for (int p = 0; p < np - 1; ++p) {
                const PathStep &st = sSteps[p];
                // One hop = ONE 32-byte read: the transition's FastRec of the bucket tLo falls into (start, link and translation
                // constant of the segment i0 holding the bucket's first base).  i0 starts at or before tLo; if [tLo, tLo + len)
                // lies inside i0's collinear run the run's affine map applies even when tLo is past i0's own end.  Otherwise
                // the exact segment is searched and tested before the interval is declared complex.
                int64_t e, x, s0;
                {
                    const FastRec *fr = &st.fast[tLo >> (st.up ? st.topShift : st.botShift)];
                    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(fr));
                    const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(fr) + 1);
                    s0 = a.x; e = a.y; x = b.x;
                    if (e < 0 || tLo - s0 + len > linkRun(e)) {
                        if (st.up) { // toParent (api/impl/halBottomSegmentIterator.cpp:40-49)
                            const int64_t i = searchFrom<true>(st.top, b.y, st.numTop, tLo);
                            const longlong2 h = __ldg(reinterpret_cast<const longlong2 *>(&st.top[i])); // start, parent link
                            x = ldS(&st.xlate[i]);
                            s0 = h.x; e = h.y;
                        } else { // toChild (api/impl/halTopSegmentIterator.cpp:36-45)
                            const int64_t i = searchFrom<false>(st.bot, b.y, st.numBot, tLo);
                            s0 = botStart(st.bot, i);
                            e = ldS(&st.child[i]);
                            x = ldS(&st.xlate[i]);
                        }
                    }
                }
                if (e < 0 || tLo - s0 + len > linkRun(e)) { ok = false; break; }
                if (linkRev(e)) {
                    tLo = x - tLo - (len - 1); // the image of the interval's last base is the lowest target position
                    rev = !rev;
                } else {
                    tLo += x;
                }
            }
        }
        int seq = 0;
        int64_t seqStart = 0;
        if (ok && P.tgtNumSeq > 1) {
            seq = seqOf(P.tgtSeqStart, P.tgtNumSeq, tLo);
            seqStart = ldS(&P.tgtSeqStart[seq]);
            if (tLo + len > ldS(&P.tgtSeqStart[seq + 1])) ok = false; // the pieces would not merge across sequences
        }
        const unsigned cm = __ballot_sync(HG_FULL, have && !ok);
        unsigned long long cbase = 0;
        if (lane == 0 && cm) cbase = atomicAdd(P.complexCount, (unsigned long long)__popc(cm));
        cbase = __shfl_sync(HG_FULL, cbase, 0);
        if (ok) {
            // the line goes straight to pool slot `item`, which is where outLoc[item] (initialised by the engine) already
            // points: when the whole batch ends here the pool IS the result in input order and nothing is gathered
            const uint8_t bs = P.strand ? P.strand[item] : (uint8_t)'+';
            const bool flip = bs == '-';
            const unsigned long long st8 = bs == '.' ? (unsigned long long)'.' : (unsigned long long)((flip != rev) ? '-' : '+');
            const unsigned long long ss8 = bs == '.' ? (unsigned long long)'.' : (unsigned long long)(flip ? '-' : '+');
            longlong2 a, b; // halgpu_lift_rec as two 16-byte stores
            a.x = tLo - seqStart; a.y = tLo + len - seqStart;
            b.x = gs;
            b.y = (long long)((unsigned long long)(uint32_t)seq | (st8 << 32) | (ss8 << 40) | (1ull << 48));
            longlong2 *dst = reinterpret_cast<longlong2 *>(&P.pool[item]);
            dst[0] = a; dst[1] = b;
        } else if (have) {
            P.complexList[cbase + (unsigned long long)lanePrefix(cm, lane)] = item;
        }
    }
}

} // namespace halgpu
