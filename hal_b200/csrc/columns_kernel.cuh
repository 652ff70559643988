// columns_kernel.cuh -- alignment column walk, one walk per PIECE of a reference segment.
//
// Replaces, for the default ColumnIterator flags (unique=false, maxInsertLength=0: every reference base is an
// independent query, api/impl/halColumnIterator.cpp:785-787), the reference call chain
//   ColumnIterator::toRight -> recursiveUpdate      api/impl/halColumnIterator.cpp:65-144, 246-355
//     updateParent / updateChild / updateNextTopDup  :557-605, 607-640, 642-681
//     updateParseUp / updateParseDown                :683-709, 711-744
//     colMapInsert (noAncestors / targets filters)   :766-819
// and the per-column reduction of halAlignmentDepth (alignmentDepth/halAlignmentDepth.cpp:262-280).
//
// The reference re-derives every row by pointer-chasing linked iterators once per base (one full tree walk per column).
// The walk of column p only looks at WHICH segment holds the current position in each genome it visits (the vertical
// hops between homologous segments keep the offset), so all columns p .. p + room, where room is the smallest distance
// from a visited position to the end of its segment in walking direction, run the identical walk with every position
// shifted by the column offset: same rows, same discovery order, collinear.  walkColumn returns that room; the kernels
// walk once per piece [p, p + room] and fill / emit the whole piece: one walk per ~segment instead of one per base.
#pragma once
#include "liftover_kernel.cuh"

namespace halgpu {

enum : uint32_t { COL_COUNT_DUPES = 1u, COL_NO_ANCESTORS = 2u, COL_NO_DUPES = 4u, COL_ONLY_ORTHOLOGS = 8u, COL_UNIQUE = 16u };

struct DepthParams {
    const GenomeTab *genomes;
    int32_t numGenomes, ref;
    int64_t first, step, n; // column i is reference position first + i*step (forward genome coordinates)
    int64_t seg0, nSegs;    // the reference segments (top array if the genome has one, else bottom) that hold the columns
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
#define HG_WALK_SMEM 12 // stack slots kept in shared memory per thread; deeper pushes spill to the local array

// Per-thread LIFO of pending walk branches.  The first HG_WALK_SMEM slots live in shared memory, laid out
// [slot][thread] so a warp's accesses are conflict-free; a local-memory array catches the (rare) deeper stacks.
// (A purely local stack cost 15.8 GB of DRAM writes per 50 M columns: profiles/r01_depthKernel_ncu.txt.)
struct WalkStack {
    int64_t (*sPos)[128];
    uint32_t (*sA)[128], (*sB)[128], (*sMeta)[128];
    WalkItem *spill;
    int tid;
    __device__ __forceinline__ void put(int sp, const WalkItem &w) {
        if (sp < HG_WALK_SMEM) {
            sPos[sp][tid] = w.pos; sA[sp][tid] = (uint32_t)w.a; sB[sp][tid] = (uint32_t)w.b;
            sMeta[sp][tid] = (uint32_t)w.g | ((uint32_t)w.type << 16) | ((uint32_t)w.rev << 19) | ((uint32_t)w.k << 20);
        } else {
            spill[sp - HG_WALK_SMEM] = w;
        }
    }
    __device__ __forceinline__ WalkItem get(int sp) const {
        if (sp < HG_WALK_SMEM) {
            WalkItem w;
            const uint32_t m = sMeta[sp][tid];
            w.pos = sPos[sp][tid]; w.a = sA[sp][tid]; w.b = sB[sp][tid];
            w.g = (int32_t)(m & 0xffffu); w.type = (uint8_t)((m >> 16) & 7u); w.rev = (uint8_t)((m >> 19) & 1u); w.k = (uint16_t)(m >> 20);
            return w;
        }
        return spill[sp - HG_WALK_SMEM];
    }
};
#define HG_WALK_STACK_DECL                                                                                       \
    __shared__ int64_t hgPos[HG_WALK_SMEM][128];                                                                   \
    __shared__ uint32_t hgA[HG_WALK_SMEM][128], hgB[HG_WALK_SMEM][128], hgMeta[HG_WALK_SMEM][128];                 \
    WalkItem hgSpill[HG_WALK_STACK - HG_WALK_SMEM];                                                                \
    WalkStack stack;                                                                                               \
    stack.sPos = hgPos; stack.sA = hgA; stack.sB = hgB; stack.sMeta = hgMeta; stack.spill = hgSpill; stack.tid = (int)threadIdx.x;

struct NoRefHook {
    __device__ __forceinline__ void operator()(int64_t, bool) const {}
};

// ColumnIterator(unique = true) as used by MafExport (maf/impl/halMafExport.cpp:46-81): class of the column of reference
// position p in a sweep that started at `window`, from the reference-genome bases its walk meets (before any row filter,
// like the visit cache and _leftmostRefPos of api/impl/halColumnIterator.cpp:771-818):
//   2  some reference-genome base lies in [window, p): an earlier column of the sweep put p into the visit cache, so
//      nextFreeIndex (:749-762) skips p without walking it
//   1  walked, but its left-most reference-genome base lies left of the window: isCanonicalOnRef (:208-212) is false and
//      MafExport does not write it (its sequences still became ColumnMap keys)
//   0  written
// Along a piece (columns p + j, rows at pos + j or pos - j) the two tests are monotone in j, so the class is constant up
// to the first j at which one of them flips: `flip` collects that bound and the piece is cut there.
struct UniqueClass {
    int64_t window, p;
    int cls;
    int64_t flip; // smallest j > 0 at which the class of column p + j may differ from cls (INT64_MAX: never)
    __device__ __forceinline__ void bound(int64_t j) { if (j > 0 && j < flip) flip = j; }
    __device__ __forceinline__ void operator()(int64_t pos, bool rev) {
        if (pos >= window && pos < p) cls = 2;
        else if (pos < window && cls == 0) cls = 1;
        if (!rev) {
            bound(window - pos);                  // pos + j reaches the window (both tests)
        } else {
            bound(pos - window + 1);              // pos - j drops below the window (both tests)
            if (pos >= p) bound((pos - p) / 2 + 1); // pos - j < p + j from here on
        }
    }
};

// Visitor: void emit(int g, int64_t pos, bool rev) for rows that pass the colMapInsert filters; refHook(pos) sees every
// base of the reference genome the walk meets, filtered or not.
// room (out): every column p .. p + room runs this same walk shifted by its offset (see the header comment)
template <class Emit, class RefHook>
__device__ __forceinline__ bool walkColumn(const GenomeTab *G, int ref, int64_t p, uint32_t flags, WalkStack &stack, Emit &&emit,
                                           RefHook &&refHook, int64_t &room) {
    room = INT64_MAX;
    // the position pos (walking direction rev) was located in the segment [s0, s1) of some array
    auto located = [&](int64_t pos, bool rev, int64_t s0, int64_t s1) {
        const int64_t r = rev ? pos - s0 : s1 - 1 - pos;
        if (r < room) room = r;
    };
    const bool noDupes = (flags & COL_NO_DUPES) != 0, noAnc = (flags & COL_NO_ANCESTORS) != 0,
               onlyOrtho = (flags & COL_ONLY_ORTHOLOGS) != 0;
    int sp = 0;
    bool ok = true;
    auto push = [&](uint8_t type, int g, int64_t a, int64_t b, int64_t pos, bool rev, int k) {
        if (sp >= HG_WALK_STACK) { ok = false; return; }
        WalkItem w;
        w.a = a; w.b = b; w.pos = pos; w.g = g; w.type = type; w.rev = rev ? 1 : 0; w.k = (uint16_t)k;
        stack.put(sp++, w);
    };
    auto report = [&](int g, int64_t pos, bool rev) {
        const GenomeTab &T = G[g];
        if (g == ref) refHook(pos, rev);
        if (noAnc && T.nc > 0) return;
        if (!T.isTarget) return;
        emit(g, pos, rev);
    };
    const GenomeTab R = G[ref];
    report(ref, p, false);
    if (R.numTop > 0) { // recursiveUpdate, reference with top segments (:254-303)
        const int64_t t = searchFrom<true>(R.top, (int64_t)__ldg(&R.topBucket[p >> R.topShift]), R.numTop, p);
        located(p, false, topStart(R.top, t), topStart(R.top, t + 1));
        push(W_DOWN, ref, t, 0, p, false, 0);
        if (!onlyOrtho) push(W_RING, ref, t, 0, p, false, 0);
        push(W_UP, ref, t, 0, p, false, 0);
    } else { // root reference (:306-354)
        const int64_t b = searchFrom<false>(R.bot, (int64_t)__ldg(&R.botBucket[p >> R.botShift]), R.numBot, p);
        located(p, false, botStart(R.bot, b), botStart(R.bot, b + 1));
        for (int k = R.nc - 1; k >= 0; --k) push(W_CHILD, ref, b, 0, p, false, k);
    }
    while (sp > 0 && ok) {
        const WalkItem w = stack.get(--sp);
        const GenomeTab T = G[w.g];
        const bool rev = w.rev != 0;
        switch (w.type) {
        case W_UP: { // updateParent
            const TopRec r = ldTop(&T.top[w.a]);
            if (r.parentEnc < 0 || T.parent < 0 || !G[T.parent].inScope) break;
            const GenomeTab P = G[T.parent];
            const int64_t pi = linkIdx(r.parentEnc);
            if (noDupes && linkIdx(ldS(&P.child[(int64_t)T.slot * P.numBot + pi])) != w.a) break; // isCanonicalParalog
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
            located(w.pos, rev, topStart(T.top, t), topStart(T.top, t + 1));
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
            const int64_t ci = linkIdx(ce);
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
            located(w.pos, rev, botStart(T.bot, b), botStart(T.bot, b + 1));
            for (int k = T.nc - 1; k >= 0; --k) push(W_CHILD, w.g, b, 0, w.pos, rev, k);
            break;
        }
        default: break;
        }
    }
    return ok;
}

// start of reference segment i (top array if the genome has one, else bottom)
__device__ __forceinline__ int64_t refSegStart(const GenomeTab &R, int64_t i) {
    return R.numTop > 0 ? topStart(R.top, i) : botStart(R.bot, i);
}

// One thread per reference segment; the thread walks once per piece of its segment and the warp then fills the
// pieces of its 32 lanes with coalesced stores.
__global__ void __launch_bounds__(128) depthKernel(const DepthParams P) {
    HG_WALK_STACK_DECL
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const GenomeTab R = P.genomes[P.ref];
    const int64_t lastPos = P.first + (P.n - 1) * P.step;
    const int64_t nRound = (P.nSegs + 31) & ~(int64_t)31; // whole warps stay in the loop together
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nRound; k += stride) {
        int64_t col = 0, colEnd = -1; // this thread's columns [col, colEnd] (indices into depth[])
        if (k < P.nSegs) {
            const int64_t s = refSegStart(R, P.seg0 + k), e = refSegStart(R, P.seg0 + k + 1) - 1;
            const int64_t lo = s > P.first ? s : P.first, hi = e < lastPos ? e : lastPos;
            if (lo <= hi) {
                col = (lo - P.first + P.step - 1) / P.step;
                colEnd = (hi - P.first) / P.step;
            }
        }
        while (__any_sync(HG_FULL, col <= colEnd)) {
            const bool active = col <= colEnd;
            int d = 0;
            int64_t pieceEnd = col;
            if (active) {
                uint64_t seen[4] = {0, 0, 0, 0}; // distinct genomes (<= 256, checked on the host)
                int rows = 0;
                int64_t room;
                const int64_t p = P.first + col * P.step;
                const bool ok = walkColumn(P.genomes, P.ref, p, P.flags, stack, [&](int g, int64_t, bool) {
                    seen[g >> 6] |= 1ull << (g & 63);
                    ++rows;
                }, NoRefHook(), room);
                if (!ok) *P.error = 1u;
                if (P.flags & COL_COUNT_DUPES) {
                    d = rows - 1;
                } else {
                    d = -1;
                    for (int w = 0; w < 4; ++w) {
                        uint64_t x = seen[w];
                        while (x) { x &= x - 1; ++d; }
                    }
                }
                const int64_t far = room / P.step; // columns col .. col + far share the walk
                pieceEnd = far < colEnd - col ? col + far : colEnd;
            }
            unsigned am = __ballot_sync(HG_FULL, active);
            while (am) {
                const int j = __ffs((int)am) - 1;
                am &= am - 1;
                const int64_t c0 = __shfl_sync(HG_FULL, col, j), c1 = __shfl_sync(HG_FULL, pieceEnd, j);
                const int dj = __shfl_sync(HG_FULL, d, j);
                for (int64_t c = c0 + lane; c <= c1; c += 32) P.depth[c] = dj;
            }
            if (active) col = pieceEnd + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Column runs: the ColumnIterator sweep (toRight per base) compressed to maximal runs of consecutive reference
// columns whose rows are the same sequences/strands advancing collinearly -- what MafBlock::canAppendColumn
// (maf/impl/halMafBlock.cpp:401-450) needs to know.  Rows of a column are kept in ColumnMap order: by
// (genome name, sequence index), discovery order within a sequence (api/inc/halColumnIterator.h:45-54).
// Pass 1 counts the pieces and rows of every reference segment, pass 2 (after two scans) walks again and stores
// each piece's first column and rows; pieceMergeKernel then joins neighbouring pieces whose rows continue each
// other exactly (same count, same genomes / sequences / strands, positions advanced by the piece length), so a
// run is as long as the alignment is collinear, not as long as the shortest segment.
// ---------------------------------------------------------------------------------------------------------
#define HG_MAX_ROWS 128

struct ColPieceParams {
    const GenomeTab *genomes;
    int32_t ref;
    uint32_t flags;
    int64_t first, n;          // columns first .. first + n - 1
    int64_t seg0, nSegs;       // reference segments holding them
    int64_t window;            // COL_UNIQUE: reference position the sweep started at (<= first)
    // pass 1 (pieceOff == NULL): per segment counts
    uint32_t *segPieces, *segRows; // nSegs + 1 entries each (last = 0) for the scans
    // pass 2: per segment offsets in, pieces out
    const uint64_t *pieceOff, *rowOff;
    int64_t *pieceCol;         // per piece: first column (relative to `first`)
    uint64_t *pieceRowOff;     // per piece: first row
    uint32_t *pieceRows;       // per piece: number of rows
    uint8_t *pieceClass;       // per piece: UniqueClass (0 without COL_UNIQUE)
    ColRowRec *rows;
    uint32_t *error;
};

// walk column p and leave its rows sorted in ColumnMap order; returns the row count or -1 on overflow
__device__ __forceinline__ int sortedColumn(const GenomeTab *G, int ref, int64_t p, uint32_t flags, WalkStack &stack, ColRowRec *rows,
                                            uint64_t *keys, UniqueClass &uc, int64_t &room) {
    int n = 0;
    bool over = false;
    const bool ok = walkColumn(G, ref, p, flags, stack, [&](int g, int64_t pos, bool rev) {
        if (n >= HG_MAX_ROWS) { over = true; return; }
        const GenomeTab &T = G[g];
        const int sq = T.numSeq > 1 ? seqOf(T.seqStart, T.numSeq, pos) : 0;
        ColRowRec r;
        r.pos = pos; r.seq = sq; r.genome = (int16_t)g; r.rev = rev ? 1 : 0; r.pad = 0;
        const uint64_t key = ((uint64_t)(uint32_t)T.nameRank << 32) | (uint32_t)sq;
        int j = n; // stable insertion: after every row with key <= this key
        while (j > 0 && keys[j - 1] > key) { rows[j] = rows[j - 1]; keys[j] = keys[j - 1]; --j; }
        rows[j] = r; keys[j] = key;
        ++n;
    }, uc, room);
    return (ok && !over) ? n : -1;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) colPieceKernel(const ColPieceParams P) {
    HG_WALK_STACK_DECL
    ColRowRec rows[EMIT ? HG_MAX_ROWS : 1];
    uint64_t keys[EMIT ? HG_MAX_ROWS : 1];
    const GenomeTab *G = P.genomes;
    const GenomeTab R = G[P.ref];
    const bool unique = (P.flags & COL_UNIQUE) != 0;
    const int64_t lastPos = P.first + P.n - 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < P.nSegs; k += stride) {
        const int64_t s = refSegStart(R, P.seg0 + k), e = refSegStart(R, P.seg0 + k + 1) - 1;
        int64_t p = s > P.first ? s : P.first;
        const int64_t hi = e < lastPos ? e : lastPos;
        uint32_t nPieces = 0, nRows = 0;
        uint64_t pieceAt = EMIT ? P.pieceOff[k] : 0, rowAt = EMIT ? P.rowOff[k] : 0;
        while (p <= hi) {
            UniqueClass uc;
            uc.window = unique ? P.window : INT64_MIN; uc.p = unique ? p : INT64_MIN; uc.cls = 0; uc.flip = INT64_MAX;
            int64_t room;
            int n;
            if (EMIT) {
                n = sortedColumn(G, P.ref, p, P.flags, stack, rows, keys, uc, room);
            } else {
                n = 0;
                const bool ok = walkColumn(G, P.ref, p, P.flags, stack, [&](int, int64_t, bool) { ++n; }, uc, room);
                if (!ok || n > HG_MAX_ROWS) n = -1;
            }
            if (n < 0) { *P.error = 1u; n = 0; }
            if (unique && uc.flip - 1 < room) room = uc.flip - 1;
            if (EMIT) {
                P.pieceCol[pieceAt] = p - P.first;
                P.pieceRowOff[pieceAt] = rowAt;
                P.pieceRows[pieceAt] = (uint32_t)n;
                P.pieceClass[pieceAt] = (uint8_t)uc.cls;
                for (int r = 0; r < n; ++r) P.rows[rowAt + r] = rows[r];
                ++pieceAt; rowAt += (uint64_t)n;
            }
            ++nPieces; nRows += (uint32_t)n;
            p = room >= hi - p ? hi + 1 : p + room + 1;
        }
        if (!EMIT) { P.segPieces[k] = nPieces; P.segRows[k] = nRows; }
    }
    if (!EMIT && blockIdx.x == 0 && threadIdx.x == 0) { P.segPieces[P.nSegs] = 0; P.segRows[P.nSegs] = 0; }
}

struct PieceMergeParams {
    const int64_t *pieceCol;
    const uint64_t *pieceRowOff;
    const uint32_t *pieceRows;
    const uint8_t *pieceClass;
    const ColRowRec *rows;
    uint32_t *isStart, *startRows; // nPieces + 1 entries each (last = 0) for the scans
    int64_t nPieces;
};
// piece q opens a new run unless its rows continue piece q - 1's exactly
__global__ void pieceMergeKernel(const PieceMergeParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= P.nPieces; q += stride) {
        uint32_t start = 0, nr = 0;
        if (q < P.nPieces) {
            nr = P.pieceRows[q];
            start = 1;
            if (q > 0 && P.pieceRows[q - 1] == nr && P.pieceClass[q - 1] == P.pieceClass[q]) {
                const int64_t len = P.pieceCol[q] - P.pieceCol[q - 1];
                const ColRowRec *a = P.rows + P.pieceRowOff[q - 1], *b = P.rows + P.pieceRowOff[q];
                bool same = true;
                for (uint32_t r = 0; r < nr && same; ++r) {
                    same = a[r].genome == b[r].genome && a[r].seq == b[r].seq && a[r].rev == b[r].rev &&
                           b[r].pos == (a[r].rev ? a[r].pos - len : a[r].pos + len);
                }
                if (same) start = 0;
            }
        }
        P.isStart[q] = start;
        P.startRows[q] = start ? nr : 0u;
    }
}

struct RunScatterParams {
    const uint32_t *isStart;
    const uint64_t *runIndex, *runRowOffset; // exclusive scans of isStart / startRows
    const int64_t *pieceCol;
    const uint64_t *pieceRowOff;
    const uint32_t *pieceRows;
    const uint8_t *pieceClass;
    const ColRowRec *rows;
    int64_t *runCol;     // nRuns + 1 (sentinel = n columns)
    uint64_t *runRowOff; // nRuns + 1 (sentinel = total rows)
    uint8_t *runClass;   // nRuns (may be NULL)
    ColRowRec *runRows;
    int64_t nPieces, nCols;
};
__global__ void runScatterKernel(const RunScatterParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= P.nPieces; q += stride) {
        if (q == P.nPieces) {
            P.runCol[P.runIndex[q]] = P.nCols;
            P.runRowOff[P.runIndex[q]] = P.runRowOffset[q];
        } else if (P.isStart[q]) {
            const uint64_t r = P.runIndex[q], to = P.runRowOffset[q], from = P.pieceRowOff[q];
            P.runCol[r] = P.pieceCol[q];
            P.runRowOff[r] = to;
            if (P.runClass) P.runClass[r] = P.pieceClass[q];
            const uint32_t nr = P.pieceRows[q];
            for (uint32_t k = 0; k < nr; ++k) P.runRows[to + k] = P.rows[from + k];
        }
    }
}

} // namespace halgpu
