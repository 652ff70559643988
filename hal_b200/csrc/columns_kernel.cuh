// columns_kernel.cuh -- per-reference-base alignment column walk, one thread per column.
//
// Replaces, for the default ColumnIterator flags (unique=false, maxInsertLength=0: every reference base is an
// independent query, api/impl/halColumnIterator.cpp:785-787), the reference call chain
//   ColumnIterator::toRight -> recursiveUpdate      api/impl/halColumnIterator.cpp:65-144, 246-355
//     updateParent / updateChild / updateNextTopDup  :557-605, 607-640, 642-681
//     updateParseUp / updateParseDown                :683-709, 711-744
//     colMapInsert (noAncestors / targets filters)   :766-819
// and the per-column reduction of halAlignmentDepth (alignmentDepth/halAlignmentDepth.cpp:262-280).
//
// The reference re-derives every row by pointer-chasing linked iterators once per base (one full tree walk
// per column).  Here every thread runs the same walk as an explicit-stack DFS in discovery order; the 32
// threads of a warp hold 32 consecutive reference bases, which almost always sit in the same segments, so
// each record fetch is one broadcast transaction per warp and the segment arrays stream through L2 once.
#pragma once
#include "liftover_kernel.cuh"

namespace halgpu {

struct GenomeTab { // one per genome, device resident
    const TopRec *top;
    const BotCore *bot;
    const int64_t *child;       // nc columns of numBot entries
    const int32_t *childGenome; // nc entries
    const uint32_t *topBucket, *botBucket;
    int64_t numTop, numBot;
    int32_t nc, parent, slot, topShift, botShift;
    uint8_t inScope, isTarget, pad[2];
};

enum : uint32_t { COL_COUNT_DUPES = 1u, COL_NO_ANCESTORS = 2u, COL_NO_DUPES = 4u, COL_ONLY_ORTHOLOGS = 8u };

struct DepthParams {
    const GenomeTab *genomes;
    int32_t numGenomes, ref;
    int64_t first, step, n; // column i is reference position first + i*step (forward genome coordinates)
    uint32_t flags;
    int32_t *depth;         // n entries
    uint32_t *error;        // set to 1 if a walk overflowed its stack
};

enum : uint8_t { W_UP = 0, W_RING, W_DOWN, W_PARSEUP, W_CHILD, W_RINGNEXT };

struct WalkItem { // 32 B
    int64_t a, b, pos;
    int32_t g;
    uint8_t type, rev;
    uint16_t k;
};

#define HG_WALK_STACK 96

// Visitor: void emit(int g, int64_t pos, bool rev) for rows that pass the colMapInsert filters.
template <class Emit>
__device__ __forceinline__ bool walkColumn(const GenomeTab *G, int ref, int64_t p, uint32_t flags, WalkItem *stack, Emit &&emit) {
    const bool noDupes = (flags & COL_NO_DUPES) != 0, noAnc = (flags & COL_NO_ANCESTORS) != 0,
               onlyOrtho = (flags & COL_ONLY_ORTHOLOGS) != 0;
    int sp = 0;
    bool ok = true;
    auto push = [&](uint8_t type, int g, int64_t a, int64_t b, int64_t pos, bool rev, int k) {
        if (sp >= HG_WALK_STACK) { ok = false; return; }
        WalkItem w;
        w.a = a; w.b = b; w.pos = pos; w.g = g; w.type = type; w.rev = rev ? 1 : 0; w.k = (uint16_t)k;
        stack[sp++] = w;
    };
    auto report = [&](int g, int64_t pos, bool rev) {
        const GenomeTab &T = G[g];
        if (noAnc && T.nc > 0) return;
        if (!T.isTarget) return;
        emit(g, pos, rev);
    };
    const GenomeTab R = G[ref];
    report(ref, p, false);
    if (R.numTop > 0) { // recursiveUpdate, reference with top segments (:254-303)
        const int64_t t = searchFrom<true>(R.top, (int64_t)__ldg(&R.topBucket[p >> R.topShift]), R.numTop, p);
        push(W_DOWN, ref, t, 0, p, false, 0);
        if (!onlyOrtho) push(W_RING, ref, t, 0, p, false, 0);
        push(W_UP, ref, t, 0, p, false, 0);
    } else { // root reference (:306-354)
        const int64_t b = searchFrom<false>(R.bot, (int64_t)__ldg(&R.botBucket[p >> R.botShift]), R.numBot, p);
        for (int k = R.nc - 1; k >= 0; --k) push(W_CHILD, ref, b, 0, p, false, k);
    }
    while (sp > 0 && ok) {
        const WalkItem w = stack[--sp];
        const GenomeTab T = G[w.g];
        const bool rev = w.rev != 0;
        switch (w.type) {
        case W_UP: { // updateParent
            const TopRec r = ldTop(&T.top[w.a]);
            if (r.parentEnc < 0 || T.parent < 0 || !G[T.parent].inScope) break;
            const GenomeTab P = G[T.parent];
            const int64_t pi = r.parentEnc >> 1;
            if (noDupes && (ldS(&P.child[(int64_t)T.slot * P.numBot + pi]) >> 1) != w.a) break; // isCanonicalParalog
            const int64_t L = topStart(T.top, w.a + 1) - r.start, f = w.pos - r.start;
            const bool fl = (r.parentEnc & 1) != 0;
            const int64_t ps = botStart(P.bot, pi);
            const int64_t pp = fl ? ps + L - 1 - f : ps + f;
            const bool pr = rev != fl;
            report(T.parent, pp, pr);
            for (int k = P.nc - 1; k >= 0; --k)
                if (k != T.slot) push(W_CHILD, T.parent, pi, 0, pp, pr, k);
            push(W_PARSEUP, T.parent, pi, 0, pp, pr, 0);
            break;
        }
        case W_PARSEUP: { // updateParseUp
            const int64_t tp = ldBot(&T.bot[w.a]).topParse;
            if (tp < 0) break;
            const int64_t t = searchFrom<true>(T.top, tp, T.numTop, w.pos);
            if (!onlyOrtho) push(W_RING, w.g, t, 0, w.pos, rev, 0);
            push(W_UP, w.g, t, 0, w.pos, rev, 0);
            break;
        }
        case W_CHILD: { // updateChild
            const int64_t ce = ldS(&T.child[(int64_t)w.k * T.numBot + w.a]);
            if (ce < 0) break;
            const int c = T.childGenome[w.k];
            if (!G[c].inScope) break;
            const GenomeTab C = G[c];
            const int64_t b0 = botStart(T.bot, w.a), L = botStart(T.bot, w.a + 1) - b0, f = w.pos - b0;
            const int64_t ci = ce >> 1;
            const bool fl = (ce & 1) != 0;
            const int64_t cs = topStart(C.top, ci);
            const int64_t cp = fl ? cs + L - 1 - f : cs + f;
            const bool cr = rev != fl;
            report(c, cp, cr);
            push(W_DOWN, c, ci, 0, cp, cr, 0);
            push(W_RING, c, ci, 0, cp, cr, 0);
            break;
        }
        case W_RING: // updateNextTopDup entry: a = starting member
        case W_RINGNEXT: { // a = current member, b = first member
            if (noDupes || T.parent < 0 || !G[T.parent].inScope) break;
            const int64_t cur = w.a, first = w.type == W_RING ? w.a : w.b;
            const TopRec rc = ldTop(&T.top[cur]);
            const int64_t nx = rc.nextPara;
            if (nx < 0) break; // W_RING: no paralogs; W_RINGNEXT never gets here with nx < 0
            const TopRec rn = ldTop(&T.top[nx]);
            const int64_t L = topStart(T.top, cur + 1) - rc.start, f = w.pos - rc.start;
            const bool fl = ((rn.parentEnc ^ rc.parentEnc) & 1) != 0;
            const int64_t np = fl ? rn.start + L - 1 - f : rn.start + f;
            const bool nr = rev != fl;
            report(w.g, np, nr);
            if (rn.nextPara >= 0 && rn.nextPara != first) push(W_RINGNEXT, w.g, nx, first, np, nr, 0);
            push(W_DOWN, w.g, nx, 0, np, nr, 0);
            break;
        }
        case W_DOWN: { // updateParseDown
            const int64_t bp = ldTop(&T.top[w.a]).botParse;
            if (bp < 0 || T.nc == 0) break;
            const int64_t b = searchFrom<false>(T.bot, bp, T.numBot, w.pos);
            for (int k = T.nc - 1; k >= 0; --k) push(W_CHILD, w.g, b, 0, w.pos, rev, k);
            break;
        }
        default: break;
        }
    }
    return ok;
}

__global__ void __launch_bounds__(128) depthKernel(const DepthParams P) {
    WalkItem stack[HG_WALK_STACK];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        uint64_t seen[4] = {0, 0, 0, 0}; // distinct genomes (<= 256, checked on the host)
        int rows = 0;
        const bool ok = walkColumn(P.genomes, P.ref, P.first + i * P.step, P.flags, stack, [&](int g, int64_t, bool) {
            seen[g >> 6] |= 1ull << (g & 63);
            ++rows;
        });
        if (!ok) *P.error = 1u;
        int d;
        if (P.flags & COL_COUNT_DUPES) {
            d = rows - 1;
        } else {
            d = -1;
            for (int w = 0; w < 4; ++w) {
                uint64_t x = seen[w];
                while (x) { x &= x - 1; ++d; }
            }
        }
        P.depth[i] = d;
    }
}

} // namespace halgpu
