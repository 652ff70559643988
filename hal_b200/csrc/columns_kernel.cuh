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

enum : uint32_t { COL_COUNT_DUPES = 1u, COL_NO_ANCESTORS = 2u, COL_NO_DUPES = 4u, COL_ONLY_ORTHOLOGS = 8u, COL_UNIQUE = 16u };

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
    __device__ __forceinline__ void operator()(int64_t) const {}
};

// ColumnIterator(unique = true) as used by MafExport (maf/impl/halMafExport.cpp:46-81): class of the column of reference
// position p in a sweep that started at `window`, from the reference-genome bases its walk meets (before any row filter,
// like the visit cache and _leftmostRefPos of api/impl/halColumnIterator.cpp:771-818):
//   2  some reference-genome base lies in [window, p): an earlier column of the sweep put p into the visit cache, so
//      nextFreeIndex (:749-762) skips p without walking it
//   1  walked, but its left-most reference-genome base lies left of the window: isCanonicalOnRef (:208-212) is false and
//      MafExport does not write it (its sequences still became ColumnMap keys)
//   0  written
struct UniqueClass {
    int64_t window, p;
    int cls;
    __device__ __forceinline__ void operator()(int64_t pos) {
        if (pos >= window && pos < p) cls = 2;
        else if (pos < window && cls == 0) cls = 1;
    }
};

// Visitor: void emit(int g, int64_t pos, bool rev) for rows that pass the colMapInsert filters; refHook(pos) sees every
// base of the reference genome the walk meets, filtered or not.
template <class Emit, class RefHook>
__device__ __forceinline__ bool walkColumn(const GenomeTab *G, int ref, int64_t p, uint32_t flags, WalkStack &stack, Emit &&emit,
                                           RefHook &&refHook) {
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
        if (g == ref) refHook(pos);
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
            for (int k = T.nc - 1; k >= 0; --k) push(W_CHILD, w.g, b, 0, w.pos, rev, k);
            break;
        }
        default: break;
        }
    }
    return ok;
}

__global__ void __launch_bounds__(128) depthKernel(const DepthParams P) {
    HG_WALK_STACK_DECL
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        uint64_t seen[4] = {0, 0, 0, 0}; // distinct genomes (<= 256, checked on the host)
        int rows = 0;
        const bool ok = walkColumn(P.genomes, P.ref, P.first + i * P.step, P.flags, stack, [&](int g, int64_t, bool) {
            seen[g >> 6] |= 1ull << (g & 63);
            ++rows;
        }, NoRefHook());
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

// ---------------------------------------------------------------------------------------------------------
// Column runs: the ColumnIterator sweep (toRight per base) compressed to maximal runs of consecutive reference
// columns whose rows are the same sequences/strands advancing collinearly -- what MafBlock::canAppendColumn
// (maf/impl/halMafBlock.cpp:401-450) needs to know.  Rows of a column are kept in ColumnMap order: by
// (genome name, sequence index), discovery order within a sequence (api/inc/halColumnIterator.h:45-54).
// ---------------------------------------------------------------------------------------------------------
#define HG_MAX_ROWS 128

struct ColSigParams {
    const GenomeTab *genomes;
    int32_t ref;
    uint32_t flags;
    int64_t first, n;
    int64_t window;        // COL_UNIQUE: reference position the sweep started at (<= first)
    uint64_t *sigA, *sigB; // n entries each: 128-bit signature of the column's normalised rows
    uint32_t *nrows;       // n entries
    uint32_t *error;
};

struct ColEmitParams {
    const GenomeTab *genomes;
    int32_t ref;
    uint32_t flags;
    int64_t first, n;         // n = number of runs
    const int64_t *runCol;    // per run: column index (relative to first)
    const uint64_t *runRowOff; // per run: first row
    ColRowRec *rows;
    uint32_t *error;
    int64_t window;           // COL_UNIQUE
    uint8_t *runClass;        // COL_UNIQUE: per run, UniqueClass of its columns
};

__device__ __forceinline__ uint64_t hgMix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// walk column p and leave its rows sorted in ColumnMap order; returns the row count or -1 on overflow
__device__ __forceinline__ int sortedColumn(const GenomeTab *G, int ref, int64_t p, uint32_t flags, WalkStack &stack, ColRowRec *rows,
                                            uint64_t *keys, UniqueClass &uc) {
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
    }, uc);
    return (ok && !over) ? n : -1;
}

// Signature of a column = 128-bit hash of its rows IN DISCOVERY ORDER, each normalised by the column offset, so
// that it is constant along a collinear run.  (Equal discovery order implies equal ColumnMap order; the converse
// can only split a run, which the host state machine handles.)  Nothing but the walk stack is stored.
__global__ void __launch_bounds__(128) colSigKernel(const ColSigParams P) {
    HG_WALK_STACK_DECL
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        uint64_t a = 0x9e3779b97f4a7c15ull, b = 0xd1b54a32d192ed03ull;
        int n = 0;
        const GenomeTab *G = P.genomes;
        UniqueClass uc;
        uc.window = (P.flags & COL_UNIQUE) ? P.window : INT64_MIN; uc.p = (P.flags & COL_UNIQUE) ? P.first + i : INT64_MIN; uc.cls = 0;
        const bool ok = walkColumn(G, P.ref, P.first + i, P.flags, stack, [&](int g, int64_t pos, bool rev) {
            const GenomeTab &T = G[g];
            const int sq = T.numSeq > 1 ? seqOf(T.seqStart, T.numSeq, pos) : 0;
            const uint64_t norm = (uint64_t)(rev ? pos + i : pos - i);
            const uint64_t id = ((((uint64_t)(uint32_t)g << 32) | (uint32_t)sq) << 1) | (rev ? 1u : 0u);
            a = hgMix(a ^ norm) + hgMix(id + 0x632be59bd9b4e019ull * (uint64_t)(n + 1));
            b = hgMix(b + id) ^ hgMix(norm * 0x9fb21c651e98df25ull + (uint64_t)n);
            ++n;
        }, uc);
        if (!ok || n > HG_MAX_ROWS) { *P.error = 1u; P.nrows[i] = 0; P.sigA[i] = 0; P.sigB[i] = 0; continue; }
        // the class is part of the signature: a run never mixes written and skipped columns
        P.sigA[i] = a ^ (0x2545f4914f6cdd1dull * (uint64_t)uc.cls); P.sigB[i] = b; P.nrows[i] = (uint32_t)n;
    }
}

struct RunFlagParams {
    const uint64_t *sigA, *sigB;
    const uint32_t *nrows;
    uint32_t *isStart, *startRows; // n + 1 entries each (last = 0) for the scans
    int64_t n;
};
__global__ void runFlagKernel(const RunFlagParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= P.n; i += stride) {
        uint32_t s = 0;
        if (i < P.n) s = (i == 0 || P.sigA[i] != P.sigA[i - 1] || P.sigB[i] != P.sigB[i - 1] || P.nrows[i] != P.nrows[i - 1]) ? 1u : 0u;
        P.isStart[i] = s;
        P.startRows[i] = s ? P.nrows[i] : 0u;
    }
}

struct RunScatterParams {
    const uint32_t *isStart;
    const uint64_t *runIndex, *rowOffset;
    int64_t *runCol;     // nRuns + 1 (sentinel = n)
    uint64_t *runRowOff; // nRuns + 1 (sentinel = total rows)
    int64_t n;
};
__global__ void runScatterKernel(const RunScatterParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= P.n; i += stride) {
        if (i == P.n || P.isStart[i]) {
            P.runCol[P.runIndex[i]] = i;
            P.runRowOff[P.runIndex[i]] = P.rowOffset[i];
        }
    }
}

// one thread per run: re-walk the run's first column and store its rows
__global__ void __launch_bounds__(128) colEmitKernel(const ColEmitParams P) {
    HG_WALK_STACK_DECL
    ColRowRec rows[HG_MAX_ROWS];
    uint64_t keys[HG_MAX_ROWS];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n; r += stride) {
        const int64_t i = P.runCol[r];
        UniqueClass uc;
        uc.window = (P.flags & COL_UNIQUE) ? P.window : INT64_MIN; uc.p = (P.flags & COL_UNIQUE) ? P.first + i : INT64_MIN; uc.cls = 0;
        const int n = sortedColumn(P.genomes, P.ref, P.first + i, P.flags, stack, rows, keys, uc);
        if (n < 0) { *P.error = 1u; continue; }
        if (P.runClass) P.runClass[r] = (uint8_t)uc.cls;
        const uint64_t off = P.runRowOff[r];
        for (int k = 0; k < n; ++k) P.rows[off + k] = rows[k];
    }
}

} // namespace halgpu
