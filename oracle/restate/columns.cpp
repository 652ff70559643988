/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
 *
 * Restates ColumnIterator::recursiveUpdate and its update* helpers for one reference base:
 *   recursiveUpdate   api/impl/halColumnIterator.cpp:246-355
 *   updateParent      :557-605     updateChild :607-640     updateNextTopDup :642-681
 *   updateParseUp     :683-709     updateParseDown :711-744  colMapInsert (filters) :766-819
 * 1-base hop (halTopSegmentIterator.cpp:36-45, halBottomSegmentIterator.cpp:40-49):
 *   f = pos - Sg;  pos' = flip ? P + L - 1 - f : P + f;  rev' = rev ^ flip
 */
#include "columns.h"
#include <algorithm>
#include <set>

namespace oracle {

namespace {
struct Walk {
    const HalView &v;
    const ColumnOpts &o;
    std::vector<ColRow> &rows;
    uint64_t *visits;

    bool scopeOk(int g) const { return o.inScope.empty() || o.inScope[g]; }
    void emit(int g, int64_t pos, bool rev) {
        const GenomeView &G = v.genomes[g];
        if (o.noAncestors && G.nc > 0) return;
        if (!o.isTarget.empty() && !o.isTarget[g]) return;
        rows.push_back(ColRow{g, pos, rev});
    }
    static int64_t findTop(const GenomeView &G, int64_t hint, int64_t pos) {
        int64_t t = hint;
        while (G.tStart(t + 1) <= pos) t++;
        return t;
    }
    static int64_t findBot(const GenomeView &G, int64_t hint, int64_t pos) {
        int64_t b = hint;
        while (G.bStart(b + 1) <= pos) b++;
        return b;
    }
    static int64_t searchTop(const GenomeView &G, int64_t pos) {
        int64_t lo = 0, hi = G.numTop - 1;
        while (lo < hi) { int64_t m = (lo + hi + 1) / 2; if (G.tStart(m) <= pos) lo = m; else hi = m - 1; }
        return lo;
    }
    static int64_t searchBot(const GenomeView &G, int64_t pos) {
        int64_t lo = 0, hi = G.numBot - 1;
        while (lo < hi) { int64_t m = (lo + hi + 1) / 2; if (G.bStart(m) <= pos) lo = m; else hi = m - 1; }
        return lo;
    }
    void tick() { if (visits) ++*visits; }

    void up(int g, int64_t t, int64_t pos, bool rev) { /* updateParent */
        const GenomeView &G = v.genomes[g];
        int64_t pi = G.tParent(t);
        if (pi < 0 || G.parent < 0 || !scopeOk(G.parent)) return;
        const GenomeView &P = v.genomes[G.parent];
        if (o.noDupes && P.bChild(pi, G.slot) != t) return; /* isCanonicalParalog, mmapTopSegment.cpp:30-40 */
        int64_t Sg = G.tStart(t), L = G.tStart(t + 1) - Sg, f = pos - Sg;
        bool flip = G.tRev(t);
        int64_t pp = flip ? P.bStart(pi) + L - 1 - f : P.bStart(pi) + f;
        bool pr = rev != flip;
        tick();
        emit(G.parent, pp, pr);
        parseUp(G.parent, pi, pp, pr);
        for (int k = 0; k < P.nc; k++)
            if (k != G.slot) child(G.parent, pi, pp, pr, k);
    }
    void parseUp(int g, int64_t b, int64_t pos, bool rev) { /* updateParseUp */
        const GenomeView &G = v.genomes[g];
        if (G.bTopParse(b) < 0) return; /* hasParseUp: root has none */
        int64_t t = findTop(G, G.bTopParse(b), pos);
        tick();
        up(g, t, pos, rev);
        if (!o.onlyOrthologs) ring(g, t, pos, rev);
    }
    void child(int g, int64_t b, int64_t pos, bool rev, int k) { /* updateChild */
        const GenomeView &G = v.genomes[g];
        int64_t ci = G.bChild(b, k);
        int c = G.children[k];
        if (ci < 0 || !scopeOk(c)) return;
        const GenomeView &C = v.genomes[c];
        int64_t Sg = G.bStart(b), L = G.bStart(b + 1) - Sg, f = pos - Sg;
        bool flip = G.bChildRev(b, k);
        int64_t cp = flip ? C.tStart(ci) + L - 1 - f : C.tStart(ci) + f;
        bool cr = rev != flip;
        tick();
        emit(c, cp, cr);
        ring(c, ci, cp, cr);
        down(c, ci, cp, cr);
    }
    void ring(int g, int64_t t, int64_t pos, bool rev) { /* updateNextTopDup */
        const GenomeView &G = v.genomes[g];
        if (o.noDupes || G.tNextPara(t) < 0 || G.parent < 0 || !scopeOk(G.parent)) return;
        int64_t first = t, cur = t;
        do {
            int64_t nx = G.tNextPara(cur);
            int64_t Sg = G.tStart(cur), L = G.tStart(cur + 1) - Sg, f = pos - Sg;
            bool flip = G.tRev(nx) != G.tRev(cur);
            pos = flip ? G.tStart(nx) + L - 1 - f : G.tStart(nx) + f;
            rev = rev != flip;
            cur = nx;
            tick();
            emit(g, pos, rev);
            down(g, cur, pos, rev);
        } while (G.tNextPara(cur) >= 0 && G.tNextPara(cur) != first);
    }
    void down(int g, int64_t t, int64_t pos, bool rev) { /* updateParseDown */
        const GenomeView &G = v.genomes[g];
        if (G.tBotParse(t) < 0) return; /* hasParseDown: leaves have none */
        int64_t b = findBot(G, G.tBotParse(t), pos);
        tick();
        for (int k = 0; k < G.nc; k++) child(g, b, pos, rev, k);
    }
};
} // namespace

void column(const HalView &v, int ref, int64_t p, const ColumnOpts &o, std::vector<ColRow> &rows, uint64_t *visits) {
    rows.clear();
    Walk w{v, o, rows, visits};
    const GenomeView &R = v.genomes[ref];
    w.emit(ref, p, false);
    if (R.numTop > 0) {
        int64_t t = Walk::searchTop(R, p);
        w.tick();
        w.up(ref, t, p, false);
        if (!o.onlyOrthologs) w.ring(ref, t, p, false);
        w.down(ref, t, p, false);
    } else {
        int64_t b = Walk::searchBot(R, p);
        w.tick();
        for (int k = 0; k < R.nc; k++) w.child(ref, b, p, false, k);
    }
}

int64_t depthOf(const HalView &, const std::vector<ColRow> &rows, bool countDupes) {
    if (countDupes) return (int64_t)rows.size() - 1;
    std::set<int32_t> g;
    for (const ColRow &r : rows) g.insert(r.genome);
    return (int64_t)g.size() - 1;
}

ColumnOpts makeColumnOpts(const HalView &v, int ref, const std::vector<int> &targets, bool noDupes, bool noAncestors,
                          bool onlyOrthologs) {
    ColumnOpts o;
    o.noDupes = noDupes; o.noAncestors = noAncestors; o.onlyOrthologs = onlyOrthologs;
    if (!targets.empty()) { /* halColumnIterator.cpp:47-51: scope = spanning tree of targets + reference */
        const size_t n = v.genomes.size();
        o.isTarget.assign(n, 0);
        o.inScope.assign(n, 0);
        std::vector<int> all(targets);
        all.push_back(ref);
        for (int t : all) o.isTarget[t] = 1;
        /* spanning tree: union of paths from each member to the MRCA of all members */
        auto depth = [&](int g) { int d = 0; while (v.genomes[g].parent >= 0) { g = v.genomes[g].parent; d++; } return d; };
        int m = all[0];
        for (int t : all) {
            int a = m, b = t;
            while (a != b) { if (depth(a) >= depth(b)) a = v.genomes[a].parent; else b = v.genomes[b].parent; }
            m = a;
        }
        for (int t : all) for (int g = t;; g = v.genomes[g].parent) { o.inScope[g] = 1; if (g == m) break; }
    }
    return o;
}

} // namespace oracle
