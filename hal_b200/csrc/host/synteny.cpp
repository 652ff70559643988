#include "synteny.hpp"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <map>
#include <queue>
#include <ostream>
#include <set>
#include <sstream>
#include <stdexcept>
#include <unordered_set>

namespace halgpu {

namespace {

struct F { // one mapped fragment / refined piece (forward genome coordinates; source is never reversed here)
    int64_t sLo, tLo, len;
    int32_t seq; // target sequence index
    bool tRev;
    int64_t tHi() const { return tLo + len - 1; }
};

// the part of f whose target extent is [a, b] (MappedSegment::slice; same rule as subRange() in liftover_kernel.cuh)
inline F sub(const F &f, int64_t a, int64_t b) {
    F r = f;
    const int64_t u = f.tRev ? (f.tLo + f.len - 1 - b) : (a - f.tLo);
    r.sLo = f.sLo + u;
    r.tLo = a;
    r.len = b - a + 1;
    return r;
}

// MappedSegmentLess (api/impl/halMappedSegment.cpp:36-43): target (start, end) then source
inline bool fragLess(const F &a, const F &b) {
    if (a.tLo != b.tLo) return a.tLo < b.tLo;
    if (a.len != b.len) return a.len < b.len;
    if (a.sLo != b.sLo) return a.sLo < b.sLo;
    return (int)a.tRev < (int)b.tRev;
}

// MappedSegment::canMergeRightWith (api/impl/halMappedSegment.cpp:109-161) + the same-sequence test of extractSegment
inline bool canMergeRight(const F &p, const F &q) {
    if (p.tRev != q.tRev || p.seq != q.seq) return false;
    if (q.tLo - p.tHi() != 1) return false;
    if (!p.tRev) return q.sLo - (p.sLo + p.len - 1) == 1;
    return p.sLo - (q.sLo + q.len - 1) == 1;
}

struct Line { // one BedLine of BlockLiftover::liftInterval (liftover/impl/halBlockLiftover.cpp:82-105)
    int32_t seq;
    int64_t start, end; // sequence relative, end exclusive
    char strand;
    int64_t srcStart;   // genome coordinate
};

double seconds(std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }

} // namespace

std::vector<PslBlock> GpuHal2Psl::convert2psl(int srcGenome, int tgtGenome, const std::string &srcChrom) {
    std::vector<PslBlock> pslBlocks;
    const halgpu_seq *sseq = nullptr, *tseq = nullptr;
    size_t ns = 0, nt = 0;
    if (halgpu_sequence_table(_ctx, srcGenome, &sseq, &ns) != 0 || halgpu_sequence_table(_ctx, tgtGenome, &tseq, &nt) != 0) {
        throw std::runtime_error("genome index out of range");
    }
    // source segments: the top array if the genome has one, else the bottom array (halBlockLiftover.cpp:24-30)
    size_t stride = 40;
    const uint8_t *segs = static_cast<const uint8_t *>(halgpu_genome_top_segments(_ctx, srcGenome));
    const bool srcTop = halgpu_genome_num_top(_ctx, srcGenome) > 0;
    if (!srcTop) segs = static_cast<const uint8_t *>(halgpu_genome_bottom_segments(_ctx, srcGenome, &stride));
    auto segStart = [&](int64_t i) {
        int64_t v;
        std::memcpy(&v, segs + stride * (size_t)i, 8);
        return v;
    };
    for (size_t si = 0; si < ns; ++si) { // SequenceIterator order
        const halgpu_seq &Q = sseq[si];
        if (srcChrom != "\"\"" && srcChrom != Q.name) continue;
        if (Q.length == 0) continue;
        // ---- intervals: windows of whole source segments covering the chromosome ----
        const int64_t numSegs = srcTop ? Q.num_top : Q.num_bottom;
        if (numSegs <= 0 || segs == nullptr) continue;
        // first segment of the sequence: segments of a sequence are contiguous in the genome's array; find it by position
        int64_t lo = 0, hi = (srcTop ? halgpu_genome_num_top(_ctx, srcGenome) : halgpu_genome_num_bottom(_ctx, srcGenome));
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (segStart(mid) <= Q.start) lo = mid; else hi = mid;
        }
        const int64_t firstSeg = lo;
        std::vector<int64_t> gs, ge;
        const int64_t per = (int64_t)std::max<size_t>(1, segmentsPerInterval);
        for (int64_t a = 0; a < numSegs; a += per) {
            const int64_t b = std::min(numSegs, a + per);
            gs.push_back(segStart(firstSeg + a));
            ge.push_back(segStart(firstSeg + b) - 1);
        }
        gs.front() = Q.start; // (== the first segment's start)
        ge.back() = Q.start + Q.length - 1;
        intervals += gs.size();
        // ---- GPU: halMapSegment for every source segment ----
        auto t0 = std::chrono::steady_clock::now();
        halgpu_lift_result *res = nullptr;
        char *err = nullptr;
        if (halgpu_liftover(_ctx, srcGenome, tgtGenome, -1, HALGPU_RAW_FRAGMENTS, gs.size(), gs.data(), ge.data(), nullptr, &res, &err) != 0) {
            std::string m = err ? err : "halgpu_liftover failed";
            halgpu_free_string(err);
            throw std::runtime_error(m);
        }
        gpuSeconds += seconds(t0);
        t0 = std::chrono::steady_clock::now();
        const halgpu_frag *raw = reinterpret_cast<const halgpu_frag *>(res->recs);
        const size_t nRaw = res->n_rec;
        fragments += nRaw;
        std::vector<int64_t> tstarts(nt);
        for (size_t i = 0; i < nt; ++i) tstarts[i] = tseq[i].start;
        auto seqOf = [&](int64_t pos) { return (int32_t)(std::upper_bound(tstarts.begin(), tstarts.end(), pos) - tstarts.begin() - 1); };
        // ---- insertAndBreakOverlaps over the whole chromosome == common refinement of all target extents ----
        std::vector<int64_t> bps;
        bps.reserve(nRaw * 2);
        for (size_t i = 0; i < nRaw; ++i) { bps.push_back(raw[i].tgt_start); bps.push_back(raw[i].tgt_start + raw[i].length); }
        std::sort(bps.begin(), bps.end());
        bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
        std::vector<F> cur;
        cur.reserve(nRaw + nRaw / 8);
        for (size_t i = 0; i < nRaw; ++i) {
            F f;
            f.sLo = raw[i].src_start; f.tLo = raw[i].tgt_start; f.len = raw[i].length; f.tRev = (raw[i].flags & 2u) != 0;
            f.seq = nt > 1 ? seqOf(f.tLo) : 0;
            const int64_t hiT = f.tHi();
            auto it = std::upper_bound(bps.begin(), bps.end(), f.tLo);
            int64_t a = f.tLo;
            for (; it != bps.end() && *it <= hiT; ++it) {
                cur.push_back(sub(f, a, *it - 1));
                a = *it;
            }
            cur.push_back(sub(f, a, hiT));
        }
        halgpu_free_result(res);
        std::sort(cur.begin(), cur.end(), fragLess);
        cur.erase(std::unique(cur.begin(), cur.end(), [](const F &x, const F &y) { return x.tLo == y.tLo && x.len == y.len && x.sLo == y.sLo; }), cur.end());
        refined += cur.size();
        // ---- BlockMapper::extractSegment over the sorted set (liftover/impl/halBlockMapper.cpp:331-394), the exact sequential
        //      form with equal-target-start classes and cut points (same as the general path of the kernel's phase 2) ----
        const size_t m = cur.size();
        std::vector<uint8_t> dead(m, 0);
        std::unordered_set<int64_t> qcut;
        std::vector<size_t> v1, v2;
        std::vector<Line> mapped;
        for (size_t x = 0; x < m; ++x) {
            if (dead[x]) continue;
            v1.clear();
            size_t tailIdx = x;
            v1.push_back(x);
            size_t nx = x + 1;
            while (nx < m && dead[nx]) ++nx;
            while (nx < m && cur[nx].tLo == cur[v1.back()].tLo) {
                v1.push_back(nx);
                ++nx;
                while (nx < m && dead[nx]) ++nx;
            }
            while (nx < m) {
                v2.clear();
                while (nx < m && (v2.empty() || cur[v2.back()].tLo == cur[nx].tLo) && v2.size() < v1.size()) {
                    v2.push_back(nx);
                    ++nx;
                    while (nx < m && dead[nx]) ++nx;
                }
                bool can = v1.size() == v2.size();
                for (size_t i = 0; i < v1.size() && can; ++i) {
                    const F &a = cur[v1[i]], &b = cur[v2[i]];
                    can = b.seq == cur[x].seq && canMergeRight(a, b) && qcut.count(a.tHi()) == 0;
                }
                if (!can) break;
                tailIdx = v2[0];
                dead[v2[0]] = 1;
                v1 = v2;
            }
            if (v1.size() > 1) qcut.insert(cur[tailIdx].tHi());
            const F &h = cur[x], &t = cur[tailIdx];
            Line l;
            l.seq = h.seq;
            l.start = std::min(h.tLo, t.tLo) - tseq[h.seq].start;
            l.end = std::max(h.tHi(), t.tHi()) + 1 - tseq[h.seq].start;
            l.strand = h.tRev ? '-' : '+';
            l.srcStart = std::min(h.sLo, t.sLo);
            mapped.push_back(l);
        }
        lines += mapped.size();
        // ---- Liftover::assignBlocksToIntervals (liftover/impl/halLiftover.cpp:108-167) with _outPSL: it only decides the
        //      ORDER in which the blocks reach dag_merge (whose std::sort is not stable), every block is kept ----
        std::stable_sort(mapped.begin(), mapped.end(), [](const Line &a, const Line &b) { return a.srcStart < b.srcStart; });
        struct Group {
            int32_t seq;
            char strand;
            int64_t srcStart;
            std::vector<size_t> blocks; // indices into mapped
        };
        std::vector<Group> groups;
        int64_t prevSrcBlockEnd = -1;
        for (size_t bi = 0; bi < mapped.size(); ++bi) {
            const Line &blk = mapped[bi];
            const int64_t srcBlockEnd = blk.srcStart + (blk.end - blk.start);
            const bool dupe = blk.srcStart < prevSrcBlockEnd || (bi + 1 < mapped.size() && mapped[bi + 1].srcStart < srcBlockEnd);
            bool fresh = groups.empty() || dupe;
            if (!fresh) { // Liftover::compatible (:169-195), input strand '+'
                const Group &g = groups.back();
                const Line &tb = mapped[g.blocks.back()];
                int64_t delta;
                if (g.strand != '+') delta = tb.start - blk.end; else delta = blk.start - tb.end;
                fresh = g.strand != blk.strand || g.srcStart == blk.srcStart || delta < 0 || g.seq != blk.seq;
            }
            if (fresh) groups.push_back(Group{blk.seq, blk.strand, blk.srcStart, {}});
            groups.back().blocks.push_back(bi);
            prevSrcBlockEnd = srcBlockEnd;
        }
        for (Group &g : groups) {
            if (g.blocks.size() > 1) { // flipBlocks (:197-234), PSL flavour
                const Line &b0 = mapped[g.blocks[0]], &b1 = mapped[g.blocks[1]];
                const int64_t delta = b1.start - b0.end;
                if ((g.strand == '-' && delta >= 0) || (g.strand != '-' && delta < 0)) std::reverse(g.blocks.begin(), g.blocks.end());
            }
            for (size_t bi : g.blocks) { // Hal2Psl::makeUpPsl (synteny/impl/hal2psl.cpp:55-91)
                const Line &l = mapped[bi];
                PslBlock b;
                b.size = (uint64_t)(l.end - l.start);
                b.qName = Q.name;
                b.qSize = (uint64_t)Q.length;
                b.qStart = (uint64_t)(l.srcStart - Q.start);
                b.qEnd = b.qStart + b.size;
                b.tSize = (uint64_t)tseq[g.seq].length;
                b.tStart = (uint64_t)l.start;
                b.tEnd = b.tStart + b.size;
                if (g.strand == '-') {
                    const uint64_t posStart = b.tStart;
                    b.tStart = b.tSize - posStart - b.size;
                    b.tEnd = b.tSize - posStart;
                }
                b.strand = std::string("+") + g.strand;
                b.tName = tseq[g.seq].name;
                pslBlocks.push_back(std::move(b));
            }
        }
        hostSeconds += seconds(t0);
    }
    return pslBlocks;
}

// ---------------------------------------------------------------------------------------------------------------------
// dag_merge (synteny/impl/psl_merger.cpp), restated with vectors in place of the std::map / std::set of ints (keys are the
// dense vertex indices 0..n-1, iterated in the same ascending order)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// one vertex of the merge DAG: the block's coordinates with its target-name / strand strings interned to integers
struct Vertex {
    uint64_t qStart, qEnd, tStart, tEnd, size;
    int tName, strand;
};
inline bool areSyntenic(const Vertex &a, const Vertex &b) { // assumes a.start < b.start
    return a.qEnd <= b.qStart && a.tEnd <= b.tStart && a.tName == b.tName && a.strand == b.strand;
}
inline bool isNotOverlappingOrderedPair(const Vertex &a, const Vertex &b, uint64_t threshold) {
    return areSyntenic(a, b) && b.qStart - a.qEnd < threshold && b.tStart - a.tEnd < threshold;
}
// get_next (psl_merger.cpp): the candidates after `pos` up to the first one that could itself follow the first candidate.
// The vertices are sorted by qStart, so once a block starts `threshold` or more past this block's query end no later block can
// pass the distance test: the scan stops there (the reference scans to the end of the group, with the same result).
std::vector<int> getNext(int pos, const std::vector<Vertex> &group, uint64_t maxAnchorDistance) {
    std::vector<int> f;
    const Vertex &a = group[pos];
    for (int i = pos + 1; i < (int)group.size(); ++i) {
        if (group[i].qStart >= a.qEnd && group[i].qStart - a.qEnd >= maxAnchorDistance) break;
        if (isNotOverlappingOrderedPair(a, group[i], maxAnchorDistance)) {
            if (f.empty()) {
                f.push_back(i);
            } else if (isNotOverlappingOrderedPair(group[f[0]], group[i], maxAnchorDistance)) {
                return f;
            } else {
                f.push_back(i);
            }
        }
    }
    return f;
}

} // namespace

// dag_merge (psl_merger.cpp): repeat { weigh the DAG of the still-visible blocks: weight(v) = size(v) + the largest weight among
// its visible predecessors (the first such predecessor in vertex order is remembered); take the heaviest vertex (the last one
// among equals), trace its chain back, emit it, hide its vertices } until every block is hidden.  The reference re-weighs the
// whole DAG for every chain; only the vertices downstream of the chain just hidden can change, so here those are re-weighed in
// ascending vertex order (a min-heap of dirty vertices) with the same recurrence -- same weights, same predecessors, same chains.
std::vector<std::vector<PslBlock>> dagMerge(const std::vector<PslBlock> &blocks, uint64_t minBlockBreath, uint64_t maxAnchorDistance) {
    std::map<std::string, std::vector<PslBlock>> blocksByQName;
    for (const PslBlock &b : blocks) blocksByQName[b.qName].push_back(b);
    std::vector<std::vector<PslBlock>> paths;
    for (auto &kv : blocksByQName) {
        std::vector<PslBlock> &group = kv.second;
        std::sort(group.begin(), group.end(), [](const PslBlock &a, const PslBlock &b) { // qStartLess; std::sort like the reference: same permutation of ties
            if (a.qStart < b.qStart) return true;
            if (a.qStart == b.qStart) return a.tStart < b.tStart;
            return false;
        });
        const int n = (int)group.size();
        std::vector<Vertex> vx(n);
        {
            std::map<std::string, int> ids;
            auto intern = [&](const std::string &x) { return ids.emplace(x, (int)ids.size()).first->second; };
            for (int i = 0; i < n; ++i) {
                const PslBlock &b = group[i];
                vx[i] = Vertex{b.qStart, b.qEnd, b.tStart, b.tEnd, b.size, intern("t" + b.tName), intern("s" + b.strand)};
            }
        }
        std::vector<std::vector<int>> dag(n), pred(n);
        for (int i = 0; i < n; ++i) {
            dag[i] = getNext(i, vx, maxAnchorDistance);
            for (int j : dag[i]) pred[j].push_back(i); // ascending i
        }
        std::vector<char> hidden(n, 0), dirty(n, 0);
        std::vector<int> prev(n, -1);
        std::vector<uint64_t> weight(n);
        auto reweigh = [&](int j) { // weigh_dag's relaxations into j, in the order the reference applies them
            bool seen = false;
            uint64_t w = vx[j].size;
            int p = -1;
            for (int i : pred[j]) {
                if (hidden[i]) continue;
                const uint64_t alt = weight[i] + vx[j].size;
                if (!seen || w < alt) { seen = true; w = alt; p = i; }
            }
            const bool changed = w != weight[j];
            weight[j] = w;
            prev[j] = p;
            return changed;
        };
        for (int j = 0; j < n; ++j) { weight[j] = 0; reweigh(j); }
        std::priority_queue<int, std::vector<int>, std::greater<int>> work;
        int numHidden = 0;
        while (numHidden != n) {
            // the heaviest visible vertex (the LAST one among equals: >= in get_maxed_vertex)
            int start = -1;
            uint64_t best = 0;
            for (int i = 0; i < n; ++i) {
                if (hidden[i]) continue;
                if (start < 0 || weight[i] >= best) { best = weight[i]; start = i; }
            }
            if (start < 0) break;
            std::vector<int> path = {start};
            for (int p = prev[start]; p != -1; p = prev[p]) path.push_back(p);
            for (int v : path) {
                if (!hidden[v]) { hidden[v] = 1; ++numHidden; }
            }
            for (int v : path) {
                for (int j : dag[v]) if (!hidden[j] && !dirty[j]) { dirty[j] = 1; work.push(j); }
            }
            while (!work.empty()) {
                const int j = work.top();
                work.pop();
                dirty[j] = 0;
                if (hidden[j]) continue;
                if (reweigh(j)) {
                    for (int k : dag[j]) if (!hidden[k] && !dirty[k]) { dirty[k] = 1; work.push(k); }
                }
            }
            std::vector<PslBlock> blockPath;
            for (auto it = path.rbegin(); it != path.rend(); ++it) blockPath.push_back(group[*it]);
            if (blockPath.empty()) break;
            const uint64_t qLen = blockPath.back().qEnd - blockPath[0].qStart;
            const uint64_t tLen = blockPath.back().tEnd - blockPath[0].tStart;
            if (qLen >= minBlockBreath && tLen >= minBlockBreath) paths.push_back(std::move(blockPath));
        }
    }
    return paths;
}

void writePsl(const std::vector<std::vector<PslBlock>> &mergedBlocks, std::ostream &os) {
    for (const std::vector<PslBlock> &blocks : mergedBlocks) { // psl_io::construct_psl (synteny/impl/psl_io.cpp:52-83) + operator<<(Psl)
        int match = 0;
        for (const PslBlock &b : blocks) match = (int)((uint64_t)(int64_t)match + b.qEnd - b.qStart);
        int qNum = 0, qBase = 0, tNum = 0, tBase = 0;
        for (size_t i = 0; i + 1 < blocks.size(); ++i) {
            const uint64_t dq = blocks[i + 1].qStart - blocks[i].qEnd, dt = blocks[i + 1].tStart - blocks[i].tEnd;
            if (dq > 0) { ++qNum; qBase += (int)dq; }
            if (dt > 0) { ++tNum; tBase += (int)dt; }
        }
        const PslBlock &f = blocks.front(), &l = blocks.back();
        uint64_t tStart = 0, tEnd = 0;
        if (f.strand == "++") { tStart = f.tStart; tEnd = l.tEnd; }
        else if (f.strand == "+-") { tEnd = f.tSize - f.tStart; tStart = f.tSize - l.tEnd; }
        std::string s;
        s += std::to_string(match); s += "\t0\t0\t0\t";
        s += std::to_string(qNum) + "\t" + std::to_string(qBase) + "\t" + std::to_string(tNum) + "\t" + std::to_string(tBase) + "\t";
        s += f.strand + "\t" + f.qName + "\t" + std::to_string(f.qSize) + "\t" + std::to_string(f.qStart) + "\t" + std::to_string(l.qEnd) + "\t";
        s += f.tName + "\t" + std::to_string(f.tSize) + "\t" + std::to_string(tStart) + "\t" + std::to_string(tEnd) + "\t";
        s += std::to_string((int)blocks.size()) + "\t";
        for (const PslBlock &b : blocks) { s += std::to_string(b.size); s += ','; }
        s += '\t';
        for (const PslBlock &b : blocks) { s += std::to_string(b.qStart); s += ','; }
        s += '\t';
        for (const PslBlock &b : blocks) { s += std::to_string(b.tStart); s += ','; }
        os << s << std::endl;
    }
}

std::vector<PslBlock> readPslBlocks(const std::string &pslPath) { // psl_io::get_blocks_set + Psl::parse (synteny/inc/psl.h:96-118)
    std::ifstream in(pslPath);
    std::vector<PslBlock> blocks;
    auto split = [](const std::string &s, char delim) {
        std::stringstream ss(s);
        std::string item;
        std::vector<std::string> elems;
        while (std::getline(ss, item, delim)) elems.push_back(item);
        return elems;
    };
    for (std::string line; std::getline(in, line);) {
        const std::vector<std::string> row = split(line, '\t');
        if (!(row.size() != 1 || (!line.empty() && line[0] == '#'))) continue;
        for (int k : {0, 1, 2, 3, 4, 5, 6, 7}) (void)std::stoi(row.at(k));
        const std::string strand = row.at(8), qName = row.at(9), tName = row.at(13);
        const uint64_t qSize = (uint64_t)std::stoi(row.at(10)), tSize = (uint64_t)std::stoi(row.at(14));
        (void)std::stoi(row.at(11)); (void)std::stoi(row.at(12)); (void)std::stoi(row.at(15)); (void)std::stoi(row.at(16));
        const int blockCount = std::stoi(row.at(17));
        std::vector<int> sizes, qStarts, tStarts;
        for (const std::string &x : split(row.at(18), ',')) sizes.push_back(std::stoi(x));
        for (const std::string &x : split(row.at(19), ',')) qStarts.push_back(std::stoi(x));
        for (const std::string &x : split(row.at(20), ',')) tStarts.push_back(std::stoi(x));
        for (int i = 0; i < blockCount; ++i) {
            PslBlock b;
            b.qStart = (uint64_t)qStarts.at(i); b.tStart = (uint64_t)tStarts.at(i); b.size = (uint64_t)sizes.at(i);
            b.qEnd = b.qStart + b.size; b.tEnd = b.tStart + b.size;
            b.strand = strand; b.qName = qName; b.tName = tName; b.qSize = (uint64_t)(int)qSize; b.tSize = (uint64_t)(int)tSize;
            blocks.push_back(std::move(b));
        }
    }
    return blocks;
}

} // namespace halgpu
