/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
 *
 * Array-level restatement of halLiftover's per-interval walk.  Each step
 * cites the reference code it restates:
 *   seed           liftover/impl/halBlockLiftover.cpp:46-72, api/impl/halSegmentIterator.cpp:208-299
 *   plan           liftover/impl/halBlockLiftover.cpp:23-44, api/impl/halCommon.cpp:123-152
 *   vertical hop   api/impl/halBottomSegmentIterator.cpp:40-49, halTopSegmentIterator.cpp:36-45
 *   parse split    api/impl/halTopSegmentIterator.cpp:55-81, halBottomSegmentIterator.cpp:51-76,
 *                  api/impl/halSegmentMapper.cpp:40-77,145-183
 *   paralogy ring  api/impl/halSegmentMapper.cpp:263-288, halTopSegmentIterator.cpp:99-107
 *   refine         api/impl/halSegmentMapper.cpp:332-520 (insertAndBreakOverlaps)
 *   order          api/impl/halMappedSegment.cpp:36-43,167-206,254-281
 *   merge          liftover/impl/halBlockMapper.cpp:331-394, api/impl/halMappedSegment.cpp:109-161
 *   output line    liftover/impl/halBlockLiftover.cpp:79-112, liftover/impl/halLiftover.cpp:90
 * Parity pinned against oracle/_ref (the reference compiled from /root/reference): see
 * tests/test_oracle.py and tests/golden/.
 */
#include "liftover.h"
#include <algorithm>
#include <set>

namespace oracle {

Plan makePlan(const HalView &v, int src, int tgt, int coal) {
    Plan p;
    p.src = src;
    p.tgt = tgt;
    std::vector<int> a, b;
    for (int g = src; g >= 0; g = v.genomes[g].parent) a.push_back(g);
    for (int g = tgt; g >= 0; g = v.genomes[g].parent) b.push_back(g);
    size_t ia = a.size(), ib = b.size();
    while (ia > 0 && ib > 0 && a[ia - 1] == b[ib - 1]) { ia--; ib--; }
    p.mrca = a[ia]; /* last common element walking down from the root */
    p.up.assign(a.begin(), a.begin() + ia + 1);
    p.down.assign(b.begin(), b.begin() + ib + 1);
    std::reverse(p.down.begin(), p.down.end());
    if (coal >= 0 && coal != p.mrca) {
        int g = p.mrca;
        while (g >= 0 && g != coal) { p.para.push_back(g); g = v.genomes[g].parent; }
        if (g < 0) throw std::runtime_error("Hit root genome when attempting to map paralogies");
    }
    return p;
}

namespace {

struct Ctx {
    const HalView &v;
    Stats *st;
    void visitTop(const GenomeView &) { if (st) { st->visitsTop++; st->visitBytes += 40 + 8; } }
    void visitBot(const GenomeView &g) { if (st) { st->visitsBot++; st->visitBytes += g.bstride + 8; } }
};

/* sub-range rule (App. A.1): the part of f whose target extent is [a,b] */
inline Frag sub(const Frag &f, int64_t a, int64_t b) {
    Frag r = f;
    int64_t u = f.tRev ? f.tHi - b : a - f.tLo;
    int64_t m = b - a + 1;
    if (f.sRev) { r.sHi = f.sHi - u; r.sLo = r.sHi - m + 1; }
    else        { r.sLo = f.sLo + u; r.sHi = r.sLo + m - 1; }
    r.tLo = a;
    r.tHi = b;
    return r;
}

/* vertical hop (App. A.4) from a segment starting at Sg to its homolog starting at P */
inline void hop(Frag &f, int64_t Sg, int64_t L, int64_t P, bool flip) {
    int64_t n = f.tHi - f.tLo + 1;
    int64_t off = f.tLo - Sg;
    f.tLo = flip ? P + L - off - n : P + off;
    f.tHi = f.tLo + n - 1;
    f.tRev ^= flip;
}

inline bool lessSourceThenTarget(const Frag &a, const Frag &b) {
    if (a.sLo != b.sLo) return a.sLo < b.sLo;
    if (a.sHi != b.sHi) return a.sHi < b.sHi;
    if (a.tLo != b.tLo) return a.tLo < b.tLo;
    return a.tHi < b.tHi;
}
inline bool sameCoords(const Frag &a, const Frag &b) {
    return a.sLo == b.sLo && a.sHi == b.sHi && a.tLo == b.tLo && a.tHi == b.tHi;
}
void sortUnique(std::vector<Frag> &l) { /* list::sort (stable) + list::unique, halSegmentMapper.cpp:122-123,257-258 */
    std::stable_sort(l.begin(), l.end(), lessSourceThenTarget);
    l.erase(std::unique(l.begin(), l.end(), sameCoords), l.end());
}

/* one upward level g -> parent (mapUp, halSegmentMapper.cpp:25-80) */
void levelUp(Ctx &c, const GenomeView &g, const GenomeView &par, const std::vector<Frag> &in, std::vector<Frag> &out) {
    for (const Frag &f0 : in) {
        auto hopTop = [&](Frag f) {
            int64_t t = f.idx;
            int64_t pi = g.tParent(t);
            if (pi < 0) return;
            int64_t Sg = g.tStart(t), L = g.tStart(t + 1) - Sg;
            hop(f, Sg, L, par.bStart(pi), g.tRev(t));
            f.top = false;
            f.idx = pi;
            c.visitBot(par);
            out.push_back(f);
        };
        if (f0.top) { hopTop(f0); continue; }
        /* bottom fragment inside g: split by g's top boundaries (toParseUp + toRight(cutoff)) */
        int64_t t = g.bTopParse(f0.idx);
        while (g.tStart(t + 1) <= f0.tLo) t++;
        for (; t < g.numTop && g.tStart(t) <= f0.tHi; t++) {
            c.visitTop(g);
            int64_t a = std::max(f0.tLo, g.tStart(t)), b = std::min(f0.tHi, g.tStart(t + 1) - 1);
            Frag p = sub(f0, a, b);
            p.top = true;
            p.idx = t;
            hopTop(p);
        }
    }
}

/* one downward level g -> child slot k (mapDown + mapSelf, halSegmentMapper.cpp:128-186,263-288) */
void levelDown(Ctx &c, const GenomeView &g, const GenomeView &ch, int k, bool dupes, const std::vector<Frag> &in,
               std::vector<Frag> &out) {
    std::vector<Frag> landed;
    for (const Frag &f0 : in) {
        auto hopBot = [&](Frag f) {
            int64_t b = f.idx;
            int64_t ci = g.bChild(b, k);
            if (ci < 0) return;
            int64_t Sg = g.bStart(b), L = g.bStart(b + 1) - Sg;
            hop(f, Sg, L, ch.tStart(ci), g.bChildRev(b, k));
            f.top = true;
            f.idx = ci;
            c.visitTop(ch);
            landed.push_back(f);
        };
        if (!f0.top) { hopBot(f0); continue; }
        int64_t b = g.tBotParse(f0.idx);
        while (g.bStart(b + 1) <= f0.tLo) b++;
        for (; b < g.numBot && g.bStart(b) <= f0.tHi; b++) {
            c.visitBot(g);
            int64_t a = std::max(f0.tLo, g.bStart(b)), e = std::min(f0.tHi, g.bStart(b + 1) - 1);
            Frag p = sub(f0, a, e);
            p.top = false;
            p.idx = b;
            hopBot(p);
        }
    }
    if (!dupes) { out.insert(out.end(), landed.begin(), landed.end()); return; }
    for (const Frag &f0 : landed) {
        /* do { emit(cur); if (hasNext) toNext; } while (cur.hasNext && cur != first)  -- :265-288 */
        Frag cur = f0;
        int64_t first = f0.idx;
        do {
            out.push_back(cur);
            int64_t nx = ch.tNextPara(cur.idx);
            if (nx >= 0) {
                int64_t Sg = ch.tStart(cur.idx), L = ch.tStart(cur.idx + 1) - Sg;
                bool flip = ch.tRev(nx) != ch.tRev(cur.idx);
                hop(cur, Sg, L, ch.tStart(nx), flip);
                cur.idx = nx;
                c.visitTop(ch);
            }
        } while (ch.tNextPara(cur.idx) >= 0 && cur.idx != first);
    }
}

/* mapSelf (halSegmentMapper.cpp:263-330) inside genome g: a top fragment yields every member of its paralogy ring
 * (itself first); a bottom fragment is first split over g's top segments (nothing at all in the root genome) */
void selfMap(Ctx &c, const GenomeView &g, const std::vector<Frag> &in, std::vector<Frag> &out) {
    auto ring = [&](const Frag &f0) {
        Frag cur = f0;
        int64_t first = f0.idx;
        do {
            out.push_back(cur);
            int64_t nx = g.tNextPara(cur.idx);
            if (nx >= 0) {
                int64_t Sg = g.tStart(cur.idx), L = g.tStart(cur.idx + 1) - Sg;
                bool flip = g.tRev(nx) != g.tRev(cur.idx);
                hop(cur, Sg, L, g.tStart(nx), flip);
                cur.idx = nx;
                c.visitTop(g);
            }
        } while (g.tNextPara(cur.idx) >= 0 && cur.idx != first);
    };
    for (const Frag &f0 : in) {
        if (f0.top) { ring(f0); continue; }
        if (g.parent < 0) continue;
        int64_t t = g.bTopParse(f0.idx);
        while (g.tStart(t + 1) <= f0.tLo) t++;
        for (; t < g.numTop && g.tStart(t) <= f0.tHi; t++) {
            c.visitTop(g);
            int64_t a = std::max(f0.tLo, g.tStart(t)), b = std::min(f0.tHi, g.tStart(t + 1) - 1);
            Frag p = sub(f0, a, b);
            p.top = true;
            p.idx = t;
            ring(p);
        }
    }
}

/* mapRecursiveParalogies (halSegmentMapper.cpp:525-576): level k of plan.para holds `in` (the interval's fragments mapped
 * up to that genome).  Their paralogs there are mapped back down to the MRCA without dupes; the fragments themselves go
 * one genome further up unless that is the limit.  Everything ends up in the MRCA. */
void paralogies(Ctx &c, const HalView &v, const Plan &plan, size_t k, const std::vector<Frag> &in, std::vector<Frag> &results) {
    results.clear();
    if (in.empty() || k >= plan.para.size()) { results = in; return; }
    const GenomeView &g = v.genomes[plan.para[k]];
    std::vector<Frag> paralogs;
    selfMap(c, g, in, paralogs);
    if (k + 1 < plan.para.size()) {
        std::vector<Frag> next;
        levelUp(c, g, v.genomes[plan.para[k + 1]], in, next);
        paralogies(c, v, plan, k + 1, next, results);
    }
    /* mapRecursiveDown(paralogs, ..., srcGenome = MRCA, doDupes = false): one level at a time, sort + unique after each */
    std::vector<Frag> cur = paralogs, nxt;
    for (size_t l = k; l > 0 && !cur.empty(); --l) {
        nxt.clear();
        const GenomeView &ch = v.genomes[plan.para[l - 1]];
        levelDown(c, v.genomes[plan.para[l]], ch, ch.slot, false, cur, nxt);
        cur.swap(nxt);
    }
    if (k > 0) sortUnique(cur);
    results.insert(results.begin(), cur.begin(), cur.end()); /* results.splice(results.begin(), paralogsMappedToSrc) */
    sortUnique(results);
}

inline bool lessTargetThenSource(const Frag &a, const Frag &b) {
    if (a.tLo != b.tLo) return a.tLo < b.tLo;
    if (a.tHi != b.tHi) return a.tHi < b.tHi;
    if (a.sLo != b.sLo) return a.sLo < b.sLo;
    return a.sHi < b.sHi;
}

} // namespace

void liftInterval(const HalView &v, const Plan &plan, bool dupes, int64_t gs, int64_t ge, char strand,
                  std::vector<OutLine> &out, Stats *stats, std::vector<Frag> *fragsOut, std::vector<size_t> *runSizesOut) {
    Ctx c{v, stats};
    const GenomeView &S = v.genomes[plan.src];
    const GenomeView &T = v.genomes[plan.tgt];
    const bool srcTop = S.numTop > 0; /* halBlockLiftover.cpp:24-30 */
    const int64_t N = srcTop ? S.numTop : S.numBot;
    auto sstart = [&](int64_t i) { return srcTop ? S.tStart(i) : S.bStart(i); };
    const bool flip = strand == '-';

    /* first source segment: index of the segment containing gs (unique, any exact search is equivalent
     * to the interpolation search of halSegmentIterator.cpp:240-299) */
    int64_t lo = 0, hi = N - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) / 2;
        if (stats) stats->searchProbes++;
        if (sstart(mid) <= gs) lo = mid; else hi = mid - 1;
    }

    std::vector<Frag> all; /* results of every seed, in insertion order */
    std::vector<Frag> cur, nxt;
    int64_t order = 0;
    for (int64_t i = lo; i < N && sstart(i) <= ge; i++) {
        int64_t a = std::max(gs, sstart(i)), b = std::min(ge, sstart(i + 1) - 1);
        if (stats) stats->seeds++;
        if (srcTop) c.visitTop(S); else c.visitBot(S);
        cur.clear();
        cur.push_back(Frag{a, b, a, b, i, 0, flip, flip, srcTop});
        /* up phase (mapRecursiveUp): one sort+unique after the deepest level, :122-123 */
        for (size_t l = 0; l + 1 < plan.up.size(); l++) {
            nxt.clear();
            levelUp(c, v.genomes[plan.up[l]], v.genomes[plan.up[l + 1]], cur, nxt);
            cur.swap(nxt);
        }
        if (plan.up.size() > 1) sortUnique(cur);
        /* paralogs that coalesce below the coalescence limit (mapSource, halSegmentMapper.cpp:617-623) */
        if (!plan.para.empty() && dupes) {
            paralogies(c, v, plan, 0, cur, nxt);
            cur.swap(nxt);
        }
        /* down phase (mapRecursiveDown) */
        for (size_t l = 0; l + 1 < plan.down.size(); l++) {
            nxt.clear();
            const GenomeView &ch = v.genomes[plan.down[l + 1]];
            levelDown(c, v.genomes[plan.down[l]], ch, ch.slot, dupes, cur, nxt);
            cur.swap(nxt);
        }
        if (plan.down.size() > 1) sortUnique(cur);
        for (Frag &f : cur) { f.order = order++; all.push_back(f); }
    }
    if (stats) stats->rawFrags += all.size();

    /* A.7: common refinement of the target extents (== incremental insertAndBreakOverlaps) */
    std::vector<int64_t> bps;
    bps.reserve(all.size() * 2);
    for (const Frag &f : all) { bps.push_back(f.tLo); bps.push_back(f.tHi + 1); }
    std::sort(bps.begin(), bps.end());
    bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
    std::vector<Frag> ref;
    for (const Frag &f : all) {
        auto it = std::upper_bound(bps.begin(), bps.end(), f.tLo);
        int64_t a = f.tLo;
        for (; it != bps.end() && *it <= f.tHi; ++it) {
            ref.push_back(sub(f, a, *it - 1));
            a = *it;
        }
        ref.push_back(sub(f, a, f.tHi));
    }
    /* set semantics: key (tLo,tHi,sLo,sHi), strand-blind, first inserted wins */
    std::stable_sort(ref.begin(), ref.end(), [](const Frag &x, const Frag &y) {
        if (lessTargetThenSource(x, y)) return true;
        if (lessTargetThenSource(y, x)) return false;
        return x.order < y.order;
    });
    ref.erase(std::unique(ref.begin(), ref.end(), sameCoords), ref.end());
    if (stats) stats->refinedFrags += ref.size();

    /* A.8: greedy merge into output lines */
    const size_t n = ref.size();
    std::vector<char> dead(n, 0);
    std::set<int64_t> qcut;
    std::vector<OutLine> lines;
    std::vector<size_t> v1, v2, run;
    auto nextAlive = [&](size_t j) { while (j < n && dead[j]) j++; return j; };
    for (size_t x = 0; x < n; x++) {
        if (dead[x]) continue;
        const int xseq = T.seqOf(ref[x].tLo);
        run.assign(1, x);
        v1.assign(1, x);
        size_t nx = nextAlive(x + 1);
        while (nx < n && ref[nx].tLo == ref[v1.back()].tLo) { v1.push_back(nx); nx = nextAlive(nx + 1); }
        while (nx < n) {
            v2.clear();
            while (nx < n && (v2.empty() || ref[v2.back()].tLo == ref[nx].tLo) && v2.size() < v1.size()) {
                v2.push_back(nx);
                nx = nextAlive(nx + 1);
            }
            bool can = v1.size() == v2.size();
            for (size_t i = 0; i < v1.size() && can; i++) {
                const Frag &p = ref[v1[i]], &q = ref[v2[i]];
                bool ok = T.seqOf(q.tLo) == xseq && p.tRev == q.tRev && p.sRev == q.sRev && q.tLo - p.tHi == 1;
                if (ok) ok = (p.sRev == p.tRev) ? (q.sLo - p.sHi == 1) : (p.sLo - q.sHi == 1);
                if (ok) ok = qcut.find(p.tHi) == qcut.end();
                can = ok;
            }
            if (!can) break;
            run.push_back(v2[0]);
            dead[v2[0]] = 2; /* erased after the scan; invisible to later starts, still skipped by nextAlive */
            v1.swap(v2);
        }
        if (v1.size() > 1) qcut.insert(ref[run.back()].tHi);
        const Frag &f0 = ref[run.front()], &f1 = ref[run.back()];
        const SeqView &sq = T.seqs[xseq];
        OutLine o;
        o.tgtSeq = xseq;
        o.start = std::min(f0.tLo, f1.tLo) - sq.start;
        o.end = std::max(f0.tHi, f1.tHi) + 1 - sq.start;
        o.strand = strand == '.' ? '.' : (ref[x].tRev ? '-' : '+');
        o.srcStart = std::min(f0.sLo, f1.sLo);
        o.srcStrand = strand == '.' ? '.' : (f0.sRev ? '-' : '+');
        o.nFrag = (int32_t)run.size();
        lines.push_back(o);
        if (fragsOut) for (size_t r : run) fragsOut->push_back(ref[r]);
        if (runSizesOut) runSizesOut->push_back(run.size());
    }
    std::stable_sort(lines.begin(), lines.end(), [](const OutLine &a, const OutLine &b) { return a.srcStart < b.srcStart; });
    if (stats) stats->outLines += lines.size();
    out.insert(out.end(), lines.begin(), lines.end());
}

} // namespace oracle
