// libhalBlockVizGpu -- the reference's blockViz C API (blockViz/inc/halBlockViz.h, blockViz/impl/halBlockViz.cpp) over the
// B200 context (include/halgpu.h).  See include/halgpu_blockviz.h for what is implemented.
//
// halGetBlocksInTargetRange = BlockMapper::init / map (liftover/impl/halBlockMapper.cpp:37-103) + readBlocks
// (halBlockViz.cpp:759-827).  BlockMapper::map is halMapSegment for every reference segment of the range -- that part runs
// on the GPU: the range is cut at segment boundaries into windows, one interval each of a halgpu_liftover call with
// HALGPU_RAW_FRAGMENTS (reference -> query genome, the coalescence limit and duplication mode passed through).  The
// MappedSegmentSet of the whole range (common refinement of the fragments' query extents), chainReferenceParalogies
// (:1072-1175), the extractSegment sweep with its cut sets, readBlock and processTargetDupes (:939-1070) are sequential
// heuristics over that set and run here on the host, restated from the lines cited.
#include "../../../include/halgpu_blockviz.h"
#include "../../../include/halgpu.h"
#include "maf_export.hpp"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

std::mutex gLock; // the reference serialises every call on one mutex too (halBlockViz.cpp:29-38)
// LodManager (lod/impl/halLodManager.cpp): the alignments of one handle by minimum query length; a plain HAL file is the
// single entry 0.  Contexts (one staged copy on the GPU each) are opened on first use.
struct LodEntry {
    std::string path;
    halgpu_ctx *ctx = nullptr;
};
struct Handle {
    std::string path;
    std::map<uint64_t, LodEntry> lod;
    uint64_t maxLodLowerBound = (uint64_t)std::numeric_limits<int64_t>::max();
};
std::map<int, Handle> gHandles;
const char *const kMaxLodToken = "max";

void handleError(const std::string &msg, char **errStr) {
    if (errStr == nullptr) {
        fprintf(stderr, "%s\n", msg.c_str());
        abort(); // (the reference throws a hal_exception through the C boundary)
    }
    *errStr = static_cast<char *>(malloc(msg.size() + 1));
    memcpy(*errStr, msg.c_str(), msg.size() + 1);
}
char *copyCString(const std::string &s) {
    char *o = static_cast<char *>(malloc(s.size() + 1));
    memcpy(o, s.c_str(), s.size() + 1);
    return o;
}
Handle &handleOf(int handle) { // checkHandle
    auto it = gHandles.find(handle);
    if (it == gHandles.end()) throw std::runtime_error("Handle " + std::to_string(handle) + "not found in alignment map");
    return it->second;
}
// LodManager::getAlignment (halLodManager.cpp:112-133): the entry with the largest minimum length <= queryLength, or level 0
// when DNA is needed
halgpu_ctx *alignmentFor(int handle, uint64_t queryLength, bool needDNA) {
    Handle &h = handleOf(handle);
    auto it = h.lod.begin();
    if (!needDNA) {
        it = h.lod.upper_bound(queryLength);
        --it;
    }
    if (it->first == h.maxLodLowerBound) {
        throw std::runtime_error("Query length " + std::to_string(queryLength) + " above maximum LOD size of " + std::to_string(h.maxLodLowerBound - 1));
    }
    if (it->second.ctx == nullptr) {
        char *err = nullptr;
        int device = 0;
        if (const char *d = getenv("HALGPU_DEVICE")) device = atoi(d);
        if (halgpu_open(it->second.path.c_str(), device, &it->second.ctx, &err) != 0) {
            std::string m = err ? err : "cannot open";
            halgpu_free_string(err);
            throw std::runtime_error(m);
        }
        if (halgpu_num_genomes(it->second.ctx) == 0) throw std::runtime_error("No genomes found in base alignment specified in " + it->second.path);
    }
    return it->second.ctx;
}
bool isLod0(int handle, uint64_t queryLength) { // LodManager::isLod0
    Handle &h = handleOf(handle);
    auto it = h.lod.upper_bound(queryLength);
    --it;
    return it == h.lod.begin();
}
halgpu_ctx *ctxOf(int handle) { return alignmentFor(handle, std::numeric_limits<uint64_t>::max(), false); } // "the lowest level of detail"
bool isHalFile(const char *path) { // detectHalAlignmentFormat: the HAL-MMAP header string
    FILE *f = fopen(path, "rb");
    if (f == nullptr) return false;
    char hdr[9] = {0};
    const size_t n = fread(hdr, 1, 8, f);
    fclose(f);
    return n == 8 && memcmp(hdr, "HAL-MMAP", 8) == 0;
}
// LodManager::loadLODFile (halLodManager.cpp:44-104): lines "<minQueryLength> <path|max>"
void loadLodFile(Handle &h, const std::string &lodPath) {
    FILE *f = fopen(lodPath.c_str(), "r");
    if (f == nullptr) throw std::runtime_error("Error opening " + lodPath);
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    size_t lineNum = 1, at = 0;
    bool foundMax = false;
    while (at <= text.size()) {
        size_t nl = text.find('\n', at);
        if (nl == std::string::npos) nl = text.size();
        const std::string line = text.substr(at, nl - at);
        at = nl + 1;
        if (line.empty()) continue;
        std::istringstream ss(line);
        uint64_t minLen = 0;
        std::string path;
        ss >> minLen >> path;
        if (foundMax) {
            throw std::runtime_error("Error on line " + std::to_string(lineNum) + " of " + lodPath + ": Limit token (" + kMaxLodToken + ") can only appear on final line.");
        }
        LodEntry e;
        if (path == kMaxLodToken) {
            foundMax = true;
            h.maxLodLowerBound = minLen;
            e.path = kMaxLodToken;
        } else if (!path.empty() && (path[0] == '/' || path.find(":/") != std::string::npos)) { // LodManager::resolvePath
            e.path = path;
        } else {
            const size_t sl = lodPath.find_last_of('/');
            e.path = sl == std::string::npos ? path : lodPath.substr(0, sl + 1) + path;
        }
        h.lod.insert(std::make_pair(minLen, e));
        ++lineNum;
    }
}
void checkLodMap(const Handle &h, const std::string &path) { // LodManager::checkMap
    if (h.lod.empty()) throw std::runtime_error("No entries were found in " + path);
    if (h.lod.begin()->first != 0) {
        throw std::runtime_error("No alignment with range value 0 found in " + path + ". A record of the form \"0 pathToOriginalHALFile\" must be present");
    }
    if (h.maxLodLowerBound == 0) throw std::runtime_error("Maximum LOD query length must be > 0");
}
int openLodOrHal(const char *inputPath, bool isLod, char **errStr) {
    for (auto &kv : gHandles) if (kv.second.path == inputPath) return kv.first; // findOrAllocHandle: one handle per path
    Handle h;
    h.path = inputPath;
    try {
        if (isLod) {
            loadLodFile(h, inputPath);
        } else {
            LodEntry e;
            e.path = inputPath;
            h.lod.insert(std::make_pair((uint64_t)0, e));
        }
        checkLodMap(h, inputPath);
    } catch (std::exception &e) {
        handleError("openLodOrHal error: " + std::string(inputPath) + ": " + e.what(), errStr);
        return -1;
    }
    int id = 0;
    while (gHandles.count(id)) ++id;
    gHandles[id] = h;
    return id;
}

// branch lengths of the newick string (Alignment::getBranchLength): name -> length of the branch to its parent
std::map<std::string, double> branchLengths(const std::string &nw) {
    std::map<std::string, double> out;
    size_t i = 0;
    while (i < nw.size()) {
        const char c = nw[i];
        if (c == '(' || c == ',' || c == ')' || c == ';' || c == ' ') { ++i; continue; }
        size_t j = i;
        while (j < nw.size() && nw[j] != ':' && nw[j] != ',' && nw[j] != ')' && nw[j] != '(' && nw[j] != ';') ++j;
        const std::string name = nw.substr(i, j - i);
        double len = 0;
        if (j < nw.size() && nw[j] == ':') {
            size_t k = j + 1;
            while (k < nw.size() && nw[k] != ',' && nw[k] != ')' && nw[k] != ';') ++k;
            len = atof(nw.substr(j + 1, k - j - 1).c_str());
            j = k;
        }
        out[name] = len;
        i = j;
    }
    return out;
}

hal_species_t *speciesOf(halgpu_ctx *ctx, int g, const std::map<std::string, double> &bl) {
    hal_species_t *cur = static_cast<hal_species_t *>(calloc(1, sizeof(hal_species_t)));
    const halgpu_seq *seqs = nullptr;
    size_t ns = 0;
    halgpu_sequence_table(ctx, g, &seqs, &ns);
    cur->name = copyCString(halgpu_genome_name(ctx, g));
    cur->length = (hal_int_t)halgpu_genome_length(ctx, g);
    cur->numChroms = (hal_int_t)ns;
    const int par = halgpu_genome_parent(ctx, g);
    if (par < 0) {
        cur->parentName = nullptr;
        cur->parentBranchLength = 0;
    } else {
        cur->parentName = copyCString(halgpu_genome_name(ctx, par));
        auto it = bl.find(cur->name);
        cur->parentBranchLength = it == bl.end() ? 0 : it->second;
    }
    return cur;
}

const char NIB[16] = {'a', 'c', 'g', 't', 'n', '?', '?', '?', 'A', 'C', 'G', 'T', 'N', '?', '?', '?'};
std::string dnaOf(halgpu_ctx *ctx, int g, int64_t pos, int64_t n) { // Sequence::getSubString
    const uint8_t *d = halgpu_genome_dna(ctx, g);
    std::string s((size_t)n, ' ');
    for (int64_t i = 0; i < n; ++i) {
        const int64_t p = pos + i;
        const uint8_t b = d[p >> 1];
        s[(size_t)i] = NIB[(p & 1) ? (b & 0xF) : (b >> 4)];
    }
    return s;
}
void reverseComplement(std::string &s) { // hal::reverseComplement(std::string&) (api/impl/halCommon.cpp)
    std::reverse(s.begin(), s.end());
    for (char &c : s) {
        switch (c) {
        case 'A': c = 'T'; break; case 'a': c = 't'; break;
        case 'C': c = 'G'; break; case 'c': c = 'g'; break;
        case 'G': c = 'C'; break; case 'g': c = 'c'; break;
        case 'T': c = 'A'; break; case 't': c = 'a'; break;
        default: break;
        }
    }
}

// one element of the MappedSegmentSet: reference ("source") piece <-> query ("mapped target") piece, forward coordinates
struct Seg {
    int64_t sLo, qLo, len;
    bool sRev, qRev;
    int32_t qSeq;
    bool alive, inPara;
    int64_t sHi() const { return sLo + len - 1; }
    int64_t qHi() const { return qLo + len - 1; }
    // oriented positions, as SegmentIterator::getStartPosition / getEndPosition report them
    int64_t qStartPos() const { return qRev ? qHi() : qLo; }
    int64_t qEndPos() const { return qRev ? qLo : qHi(); }
    int64_t sStartPos() const { return sRev ? sHi() : sLo; }
    int64_t sEndPos() const { return sRev ? sLo : sHi(); }
};

inline Seg sub(const Seg &f, int64_t a, int64_t b) { // the part of f whose query extent is [a, b] (MappedSegment::slice)
    Seg r = f;
    const int64_t u = f.qRev ? (f.qLo + f.len - 1 - b) : (a - f.qLo);
    const int64_t m = b - a + 1;
    r.sLo = f.sRev ? (f.sLo + f.len - u - m) : (f.sLo + u);
    r.qLo = a;
    r.len = m;
    return r;
}
inline bool segLess(const Seg &a, const Seg &b) { // MappedSegmentLess: mapped target first, then source
    if (a.qLo != b.qLo) return a.qLo < b.qLo;
    if (a.len != b.len) return a.len < b.len;
    if (a.sLo != b.sLo) return a.sLo < b.sLo;
    return ((int)a.sRev | ((int)a.qRev << 1)) < ((int)b.sRev | ((int)b.qRev << 1));
}

// MappedSegment::canMergeRightWith (api/impl/halMappedSegment.cpp:109-161) with both cut sets
bool canMergeRightWith(const Seg &a, const Seg &n, const std::set<int64_t> &cutSet, const std::set<int64_t> &sourceCutSet) {
    if (a.qRev != n.qRev || a.sRev != n.sRev) return false;
    int64_t qdelta, rdelta, cut, sourceCut;
    if (!a.qRev && !a.sRev) {
        qdelta = n.qStartPos() - a.qEndPos(); rdelta = n.sStartPos() - a.sEndPos(); cut = a.qEndPos(); sourceCut = a.sEndPos();
    } else if (a.qRev && a.sRev) {
        qdelta = n.qEndPos() - a.qStartPos(); rdelta = n.sEndPos() - a.sStartPos(); cut = a.qStartPos(); sourceCut = a.sStartPos();
    } else if (!a.qRev && a.sRev) {
        qdelta = n.qStartPos() - a.qEndPos(); rdelta = a.sEndPos() - n.sStartPos(); cut = a.qEndPos(); sourceCut = n.sStartPos();
    } else {
        qdelta = n.qEndPos() - a.qStartPos(); rdelta = a.sStartPos() - n.sEndPos(); cut = a.qStartPos(); sourceCut = n.sEndPos();
    }
    if (!(qdelta == 1 && rdelta == 1)) return false;
    if (sourceCutSet.count(sourceCut)) return false;
    if (cutSet.count(cut)) return false;
    return true;
}

size_t nextAlive(const std::vector<Seg> &v, size_t i) {
    while (i < v.size() && !v[i].alive) ++i;
    return i;
}

// chainReferenceParalogies (halBlockViz.cpp:1072-1175): keep the query genome single copy by greedy chaining
void chainReferenceParalogies(std::vector<Seg> &segs, double minChainPct = 0.025) {
    std::vector<std::vector<size_t>> chains;
    std::vector<int64_t> chainSizes;
    std::deque<int64_t> chainStack;
    std::vector<size_t> filtered;
    const size_t n = segs.size();
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        int64_t copies = 1;
        while (j < n && (segs[j].qStartPos() == segs[i].qStartPos() || segs[j].qEndPos() == segs[i].qStartPos())) { ++j; ++copies; }
        int64_t bestScore = -(int64_t)std::numeric_limits<int32_t>::max(), bestStackIdx = -1;
        size_t best = n, leftmost = n;
        int64_t leftSrcPos = std::numeric_limits<int64_t>::max();
        for (size_t k = i; k < j; ++k) {
            for (int64_t csi = (int64_t)chainStack.size() - 1; csi >= 0; --csi) {
                const Seg &back = segs[chains[(size_t)chainStack[(size_t)csi]].back()];
                int64_t srcDelta = segs[k].sStartPos() - back.sEndPos();
                if (segs[k].qRev) srcDelta = -srcDelta;
                const int64_t tgtDelta = segs[k].qStartPos() - back.qEndPos();
                if (srcDelta >= 0 && tgtDelta >= 0) {
                    const int64_t score = chainSizes[(size_t)chainStack[(size_t)csi]] * 2 - tgtDelta - srcDelta;
                    if (score > bestScore) { bestStackIdx = csi; bestScore = score; best = k; }
                }
            }
            const int64_t mpos = std::min(segs[k].qStartPos(), segs[k].qEndPos());
            if (mpos < leftSrcPos) { leftSrcPos = mpos; leftmost = k; }
        }
        if (bestStackIdx < 0) {
            best = leftmost;
            chains.push_back({best});
            chainSizes.push_back(segs[best].len);
            chainStack.push_back((int64_t)chains.size() - 1);
        } else {
            chains[(size_t)chainStack[(size_t)bestStackIdx]].push_back(best);
            chainSizes[(size_t)chainStack[(size_t)bestStackIdx]] += segs[best].len;
            while ((int64_t)chainStack.size() - 1 > bestStackIdx) chainStack.pop_back();
        }
        if (copies > 1) {
            for (size_t k = i; k < j; ++k) {
                segs[k].inPara = true;
                if (k != best) filtered.push_back(k);
            }
        }
        i = j;
    }
    for (size_t k : filtered) segs[k].alive = false;
    int64_t total = 0;
    for (int64_t s : chainSizes) total += s;
    for (size_t c = 0; c < chains.size(); ++c) {
        if ((double)chainSizes[c] / (double)total < minChainPct) {
            for (size_t k : chains[c]) segs[k].alive = false;
        }
    }
}

// processTargetDupes (halBlockViz.cpp:939-1070) over the paralogy set (set order; elements erased from the map still count)
hal_target_dupe_list_t *processTargetDupes(const std::vector<Seg> &para, const std::string &chromName, int64_t chromOffset) {
    std::vector<std::pair<std::set<int64_t>, int64_t>> lists;
    for (size_t i = 0; i < para.size();) {
        size_t j = i + 1;
        while (j < para.size() && (para[j].qStartPos() == para[i].qStartPos() || para[j].qEndPos() == para[i].qStartPos())) ++j;
        std::set<int64_t> starts;
        for (size_t k = i; k < j; ++k) starts.insert(para[k].sStartPos());
        lists.push_back(std::make_pair(starts, para[i].len));
        i = j;
    }
    std::sort(lists.begin(), lists.end(), [](const std::pair<std::set<int64_t>, int64_t> &a, const std::pair<std::set<int64_t>, int64_t> &b) {
        return *a.first.begin() < *b.first.begin();
    });
    for (size_t i = 0; i < lists.size(); ++i) {
        if (lists[i].second <= 0) continue;
        for (size_t j = i + 1; j < lists.size(); ++j) {
            bool merged = false;
            if (lists[j].first.size() == lists[i].first.size()) {
                auto k1 = lists[i].first.begin();
                auto k2 = lists[j].first.begin();
                int64_t minExtension = std::numeric_limits<int64_t>::max();
                for (; k1 != lists[i].first.end(); ++k1, ++k2) {
                    int64_t leftOverlap = -1;
                    if (*k2 >= *k1) {
                        leftOverlap = (*k1 + lists[i].second) - *k2;
                        if (leftOverlap > 0) leftOverlap = std::min(leftOverlap, lists[j].second);
                    }
                    const int64_t rightExtension = leftOverlap < 0 ? -1 : leftOverlap - lists[j].second;
                    minExtension = std::min(minExtension, rightExtension);
                }
                if (minExtension == 0) {
                    lists[j].second = 0;
                } else if (minExtension > 0) {
                    lists[i].second += minExtension;
                    lists[j].second -= minExtension;
                }
                merged = minExtension >= 0;
            }
            if (!merged) break;
        }
    }
    hal_target_dupe_list_t *head = nullptr, *tail = nullptr;
    int64_t curId = 0, prev = -1;
    for (size_t i = 0; i < lists.size(); ++i) {
        if (lists[i].second == 0) continue;
        hal_target_dupe_list_t *d = static_cast<hal_target_dupe_list_t *>(calloc(1, sizeof(hal_target_dupe_list_t)));
        if (prev >= 0) {
            const int64_t prevEnd = *lists[(size_t)prev].first.begin() + lists[(size_t)prev].second;
            if (*lists[i].first.begin() > prevEnd) ++curId;
        }
        d->id = (hal_int_t)curId;
        d->qChrom = copyCString(chromName);
        hal_target_range_t *rt = nullptr;
        for (int64_t s : lists[i].first) {
            hal_target_range_t *r = static_cast<hal_target_range_t *>(calloc(1, sizeof(hal_target_range_t)));
            r->tStart = (hal_int_t)(s - chromOffset);
            r->size = (hal_int_t)lists[i].second;
            if (rt == nullptr) d->tRange = r; else rt->next = r;
            rt = r;
        }
        if (head == nullptr) head = d; else tail->next = d;
        tail = d;
        prev = (int64_t)i;
    }
    return head;
}


// ---------------------------------------------------------------------------------------------------------------------
// BlockMapper::mapAdjacencies (liftover/impl/halBlockMapper.cpp:121-245), _maxAdjScan == 1: for every element of the set
// (in set order, while the set grows) the query segment to its right and the one to its left -- or what is left of its own
// segment -- are cut by the neighbouring set element (cutByNext, :272-329), mapped BACK to the reference genome, and the
// results that land on the reference sequence without overlapping their neighbours in the set join it ("off-screen" blocks).
// The back-mapping of every candidate piece is done beforehand in ONE raw-fragment GPU call (uncut pieces; cutting a piece
// restricts its fragments), the sequential part below only slices and inserts.
// ---------------------------------------------------------------------------------------------------------------------
struct SegSetLess {
    bool operator()(const Seg &a, const Seg &b) const { return segLess(a, b); }
};
typedef std::set<Seg, SegSetLess> SegSet;

struct QueryArray { // the query genome's segment array the mapped segments live in
    const uint8_t *base;
    size_t stride;
    int64_t n;
    int64_t start(int64_t i) const { int64_t v; memcpy(&v, base + stride * (size_t)i, 8); return v; }
    int64_t indexOf(int64_t pos) const {
        int64_t lo = 0, hi = n;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (start(mid) <= pos) lo = mid; else hi = mid; }
        return lo;
    }
};

struct Piece { // an adjacent query piece in forward coordinates, oriented like the element it belongs to
    int64_t lo, hi, idx;
    bool valid;
};

// restriction of raw back-mapped fragments (source = query piece, forward; target = reference) to source range [lo, hi]
void restrictFrags(const halgpu_frag *f, size_t n, int64_t lo, int64_t hi, bool reversed, std::vector<Seg> &out) {
    for (size_t i = 0; i < n; ++i) {
        const int64_t s0 = f[i].src_start, s1 = s0 + f[i].length - 1;
        const int64_t a = std::max(s0, lo), b = std::min(s1, hi);
        if (a > b) continue;
        const bool tRev = (f[i].flags & 2u) != 0;
        // source offset u = a - s0 (source forward in the precomputed call); target piece accordingly
        const int64_t m = b - a + 1, u = a - s0;
        Seg g;
        g.sLo = a; g.len = m;
        g.qLo = tRev ? f[i].tgt_start + f[i].length - u - m : f[i].tgt_start + u; // ("q" fields hold the mapped side = reference here)
        g.sRev = reversed;           // a reversed iterator maps with both strands flipped
        g.qRev = tRev != reversed;
        g.qSeq = 0; g.alive = true; g.inPara = false;
        out.push_back(g);
    }
}

void mapAdjacencies(halgpu_ctx *ctx, int tGenome, int qGenome, bool doDupes, bool qTop, const halgpu_seq *qseq, size_t nqs, const halgpu_seq &refSeq,
                    std::vector<Seg> &segsVec) {
    QueryArray Q;
    size_t stride = 40;
    if (qTop) {
        Q.base = static_cast<const uint8_t *>(halgpu_genome_top_segments(ctx, qGenome));
        Q.n = halgpu_genome_num_top(ctx, qGenome);
    } else {
        Q.base = static_cast<const uint8_t *>(halgpu_genome_bottom_segments(ctx, qGenome, &stride));
        Q.n = halgpu_genome_num_bottom(ctx, qGenome);
    }
    Q.stride = stride;
    if (Q.base == nullptr || Q.n <= 0 || segsVec.empty()) return;
    // [minIndex, maxIndex) of every query sequence: segments of a sequence are contiguous, sequences in order
    std::vector<int64_t> firstIdx(nqs + 1, 0);
    for (size_t i = 0; i < nqs; ++i) firstIdx[i + 1] = firstIdx[i] + (qTop ? qseq[i].num_top : qseq[i].num_bottom);
    auto rightPiece = [&](const Seg &e) { // queryIt->toRight() from the element's slice
        Piece p;
        const int64_t idx = Q.indexOf(e.qLo), S = Q.start(idx), E = Q.start(idx + 1) - 1;
        const int64_t mn = firstIdx[(size_t)e.qSeq], mx = firstIdx[(size_t)e.qSeq + 1];
        if (!e.qRev) {
            if (e.qHi() == E) { p.idx = idx + 1; p.valid = p.idx >= mn && p.idx < mx; if (p.valid) { p.lo = Q.start(p.idx); p.hi = Q.start(p.idx + 1) - 1; } }
            else { p.idx = idx; p.lo = e.qHi() + 1; p.hi = E; p.valid = true; }
        } else {
            if (e.qLo == S) { p.idx = idx - 1; p.valid = p.idx >= mn && p.idx < mx; if (p.valid) { p.lo = Q.start(p.idx); p.hi = Q.start(p.idx + 1) - 1; } }
            else { p.idx = idx; p.lo = S; p.hi = e.qLo - 1; p.valid = true; }
        }
        return p;
    };
    auto leftPiece = [&](const Seg &e) { // queryIt->toLeft()
        Piece p;
        const int64_t idx = Q.indexOf(e.qLo), S = Q.start(idx), E = Q.start(idx + 1) - 1;
        const int64_t mn = firstIdx[(size_t)e.qSeq], mx = firstIdx[(size_t)e.qSeq + 1];
        if (!e.qRev) {
            if (e.qLo == S) { p.idx = idx - 1; p.valid = p.idx >= mn && p.idx < mx; if (p.valid) { p.lo = Q.start(p.idx); p.hi = Q.start(p.idx + 1) - 1; } }
            else { p.idx = idx; p.lo = S; p.hi = e.qLo - 1; p.valid = true; }
        } else {
            if (e.qHi() == E) { p.idx = idx + 1; p.valid = p.idx >= mn && p.idx < mx; if (p.valid) { p.lo = Q.start(p.idx); p.hi = Q.start(p.idx + 1) - 1; } }
            else { p.idx = idx; p.lo = e.qHi() + 1; p.hi = E; p.valid = true; }
        }
        return p;
    };
    // ---- one GPU call: every candidate piece mapped back to the reference genome ----
    std::vector<int64_t> gs, ge;
    std::map<std::pair<int64_t, int64_t>, size_t> candidate; // (lo, hi) -> interval
    auto addCandidate = [&](const Piece &p) {
        if (!p.valid) return;
        auto key = std::make_pair(p.lo, p.hi);
        if (candidate.count(key)) return;
        candidate[key] = gs.size();
        gs.push_back(p.lo);
        ge.push_back(p.hi);
    };
    for (const Seg &e : segsVec) { addCandidate(rightPiece(e)); addCandidate(leftPiece(e)); }
    halgpu_lift_result *res = nullptr;
    if (!gs.empty()) {
        char *err = nullptr;
        const uint32_t flags = HALGPU_RAW_FRAGMENTS | (doDupes ? 0u : (uint32_t)HALGPU_NO_DUPES) | (qTop ? 0u : (uint32_t)HALGPU_SEED_BOTTOM);
        if (halgpu_liftover(ctx, qGenome, tGenome, -1, flags, gs.size(), gs.data(), ge.data(), nullptr, &res, &err) != 0) {
            std::string m = err ? err : "halgpu_liftover failed";
            halgpu_free_string(err);
            throw std::runtime_error(m);
        }
    }
    const halgpu_frag *raw = res ? reinterpret_cast<const halgpu_frag *>(res->recs) : nullptr;

    SegSet segSet(segsVec.begin(), segsVec.end());
    SegSet adjSet;
    auto overlapsPos = [](const Seg &s, int64_t pos) { return pos >= s.qLo && pos <= s.qHi(); };
    for (SegSet::const_iterator it = segSet.begin(); it != segSet.end(); ++it) {
        if (adjSet.count(*it)) continue;
        const Seg e = *it;
        std::vector<Seg> back; // halMapSegment results of the right and the left piece
        for (int side = 0; side < 2; ++side) {
            Piece p = side == 0 ? rightPiece(e) : leftPiece(e);
            if (!p.valid) continue;
            const auto key = std::make_pair(p.lo, p.hi);
            // neighbour in the CURRENT set: right scan -> next element (previous for a reversed iterator); left scan the other way
            const bool wantNext = (side == 0) != e.qRev;
            SegSet::const_iterator nb = it;
            bool haveNb = true;
            if (wantNext) { ++nb; haveNb = nb != segSet.end(); }
            else if (nb == segSet.begin()) haveNb = false;
            else --nb;
            bool wasCut = false;
            if (haveNb && Q.indexOf(nb->qLo) == p.idx) { // cutByNext(queryIt, neighbour's target, right)
                const bool right = side == 0 ? !e.qRev : e.qRev;
                if (right) {
                    if (p.lo >= nb->qLo) wasCut = true;
                    else if (p.hi >= nb->qLo) p.hi = nb->qLo - 1;
                } else {
                    if (p.hi <= nb->qHi()) wasCut = true;
                    else if (p.lo <= nb->qHi()) p.lo = nb->qHi() + 1;
                }
            }
            if (wasCut) continue;
            const size_t iv = candidate[key];
            restrictFrags(raw + res->offsets[iv], (size_t)(res->offsets[iv + 1] - res->offsets[iv]), p.lo, p.hi, e.qRev, back);
        }
        if (back.empty()) continue;
        // backResults: one MappedSegmentSet (insertAndBreakOverlaps over the mapped = reference side)
        std::vector<int64_t> bps;
        for (const Seg &b : back) { bps.push_back(b.qLo); bps.push_back(b.qLo + b.len); }
        std::sort(bps.begin(), bps.end());
        bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
        std::vector<Seg> refined;
        for (const Seg &b : back) {
            auto c = std::upper_bound(bps.begin(), bps.end(), b.qLo);
            int64_t a = b.qLo;
            for (; c != bps.end() && *c <= b.qHi(); ++c) { refined.push_back(sub(b, a, *c - 1)); a = *c; }
            refined.push_back(sub(b, a, b.qHi()));
        }
        std::sort(refined.begin(), refined.end(), segLess);
        refined.erase(std::unique(refined.begin(), refined.end(), [](const Seg &x, const Seg &y) { return x.qLo == y.qLo && x.len == y.len && x.sLo == y.sLo; }),
                      refined.end());
        // flip (source <-> target), make the reference side forward, keep what lies on the reference sequence and does not
        // overlap its neighbours in the set (:196-216)
        SegSet outSet;
        for (const Seg &b : refined) {
            if (b.qLo < refSeq.start || b.qLo >= refSeq.start + refSeq.length) continue; // mseg->getSequence() == _refSequence
            Seg m;
            m.sLo = b.qLo; m.qLo = b.sLo; m.len = b.len;
            m.sRev = b.qRev; m.qRev = b.sRev;
            if (m.sRev) { m.sRev = false; m.qRev = !m.qRev; } // fullReverse
            m.qSeq = e.qSeq; m.alive = true; m.inPara = false;
            SegSet::const_iterator j = segSet.lower_bound(m);
            if (j != segSet.begin()) --j;
            bool overlaps = false;
            for (size_t count = 0; count < 3 && j != segSet.end() && !overlaps; ++count, ++j) {
                overlaps = overlapsPos(m, j->qStartPos()) || overlapsPos(m, j->qEndPos()) || overlapsPos(*j, m.qStartPos()) || overlapsPos(*j, m.qEndPos());
            }
            if (!overlaps) outSet.insert(m);
        }
        // one copy per query interval: the one whose reference piece is nearest to this element's (:219-243)
        for (SegSet::const_iterator i = outSet.begin(); i != outSet.end();) {
            SegSet::const_iterator j = i;
            ++j;
            while (j != outSet.end() && (j->qStartPos() == i->qStartPos() || j->qEndPos() == i->qStartPos())) ++j;
            SegSet::const_iterator best = i;
            int64_t bestDelta = std::numeric_limits<int64_t>::max();
            for (SegSet::const_iterator k = i; k != j; ++k) {
                const int64_t d1 = k->sStartPos() - e.sStartPos(), d2 = k->sEndPos() - e.sStartPos();
                const int64_t delta = std::min(d1 < 0 ? -d1 : d1, d2 < 0 ? -d2 : d2);
                if (delta < bestDelta) { bestDelta = delta; best = k; }
            }
            segSet.insert(*best);
            adjSet.insert(*best);
            i = j;
        }
    }
    if (res) halgpu_free_result(res);
    segsVec.assign(segSet.begin(), segSet.end());
}

int findSeq(const halgpu_seq *seqs, size_t n, const char *name) {
    for (size_t i = 0; i < n; ++i) if (strcmp(seqs[i].name, name) == 0) return (int)i;
    return -1;
}

hal_block_results_t *readBlocks(halgpu_ctx *ctx, halgpu_ctx *seqCtx, int tGenome, int tSeqIdx, int64_t absStart, int64_t absEnd, bool tReversed, int qGenome,
                                bool getSeq, bool doDupes, bool doTargetDupes, bool doAdjes, const char *limitName) {
    const halgpu_seq *tseq = nullptr, *qseq = nullptr;
    size_t nts = 0, nqs = 0;
    halgpu_sequence_table(ctx, tGenome, &tseq, &nts);
    halgpu_sequence_table(ctx, qGenome, &qseq, &nqs);
    const std::string qGenomeName = halgpu_genome_name(ctx, qGenome);
    // coalescence limit (halBlockViz.cpp:766-789): self-alignment tracks walk back to the root by default
    int coal = -1;
    if (qGenome == tGenome && limitName == nullptr) {
        int g = tGenome;
        while (halgpu_genome_parent(ctx, g) >= 0) g = halgpu_genome_parent(ctx, g);
        coal = g;
    } else if (limitName != nullptr) {
        coal = halgpu_genome_id(ctx, limitName);
        if (coal < 0) throw std::runtime_error("Could not find coalescence limit " + std::string(limitName) + " in alignment");
    }
    // BlockMapper::map (halBlockMapper.cpp:76-83): bottom segments when the reference genome is the MRCA, else top segments
    const int mrca = halgpu_mrca(ctx, tGenome, qGenome);
    const bool seedBottom = mrca == tGenome && tGenome != qGenome;
    size_t stride = 40;
    const uint8_t *segArr;
    int64_t numSegs;
    if (seedBottom) {
        segArr = static_cast<const uint8_t *>(halgpu_genome_bottom_segments(ctx, tGenome, &stride));
        numSegs = halgpu_genome_num_bottom(ctx, tGenome);
    } else {
        segArr = static_cast<const uint8_t *>(halgpu_genome_top_segments(ctx, tGenome));
        numSegs = halgpu_genome_num_top(ctx, tGenome);
    }
    hal_block_results_t *results = static_cast<hal_block_results_t *>(calloc(1, sizeof(hal_block_results_t)));
    if (numSegs <= 0 || segArr == nullptr) return results; // (a leaf reference that is the MRCA cannot happen; a root has no tops)
    auto segStart = [&](int64_t i) {
        int64_t v;
        memcpy(&v, segArr + stride * (size_t)i, 8);
        return v;
    };
    int64_t lo = 0, hi = numSegs;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (segStart(mid) <= absStart) lo = mid; else hi = mid;
    }
    std::vector<int64_t> gs, ge;
    std::vector<uint8_t> st;
    const int64_t per = 48;
    for (int64_t a = lo; a < numSegs && segStart(a) <= absEnd; a += per) {
        const int64_t b = std::min(numSegs, a + per);
        gs.push_back(std::max(absStart, segStart(a)));
        ge.push_back(std::min(absEnd, segStart(b) - 1));
        st.push_back(tReversed ? '-' : '+');
    }
    halgpu_lift_result *res = nullptr;
    char *err = nullptr;
    const uint32_t flags = HALGPU_RAW_FRAGMENTS | (doDupes ? 0u : (uint32_t)HALGPU_NO_DUPES) | (seedBottom ? (uint32_t)HALGPU_SEED_BOTTOM : 0u);
    if (halgpu_liftover(ctx, tGenome, qGenome, coal, flags, gs.size(), gs.data(), ge.data(), st.data(), &res, &err) != 0) {
        std::string m = err ? err : "halgpu_liftover failed";
        halgpu_free_string(err);
        free(results);
        throw std::runtime_error(m);
    }
    const halgpu_frag *raw = reinterpret_cast<const halgpu_frag *>(res->recs);
    const size_t nRaw = res->n_rec;
    std::vector<int64_t> qstarts(nqs);
    for (size_t i = 0; i < nqs; ++i) qstarts[i] = qseq[i].start;
    // the MappedSegmentSet: common refinement of the query extents, set order, exact repeats dropped
    std::vector<int64_t> bps;
    for (size_t i = 0; i < nRaw; ++i) { bps.push_back(raw[i].tgt_start); bps.push_back(raw[i].tgt_start + raw[i].length); }
    std::sort(bps.begin(), bps.end());
    bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
    std::vector<Seg> segs;
    for (size_t i = 0; i < nRaw; ++i) {
        Seg f;
        f.sLo = raw[i].src_start; f.qLo = raw[i].tgt_start; f.len = raw[i].length;
        f.sRev = (raw[i].flags & 1u) != 0; f.qRev = (raw[i].flags & 2u) != 0;
        f.qSeq = (int32_t)(std::upper_bound(qstarts.begin(), qstarts.end(), f.qLo) - qstarts.begin() - 1);
        f.alive = true; f.inPara = false;
        const int64_t hiQ = f.qHi();
        auto it = std::upper_bound(bps.begin(), bps.end(), f.qLo);
        int64_t a = f.qLo;
        for (; it != bps.end() && *it <= hiQ; ++it) { segs.push_back(sub(f, a, *it - 1)); a = *it; }
        segs.push_back(sub(f, a, hiQ));
    }
    halgpu_free_result(res);
    std::sort(segs.begin(), segs.end(), segLess);
    segs.erase(std::unique(segs.begin(), segs.end(), [](const Seg &x, const Seg &y) { return x.qLo == y.qLo && x.len == y.len && x.sLo == y.sLo; }), segs.end());

    if (doAdjes) {
        // kind of the mapped segments in the query genome: arrived from the parent (top) unless the query genome is the MRCA and
        // was reached from below (bottom) -- but with a coalescence limit above the MRCA everything in the MRCA went through
        // mapSelf, which turns bottom segments into their top pieces; a self-alignment stays on the top array
        const bool coalActive = doDupes && coal >= 0 && coal != mrca;
        const bool qTop = !(mrca == qGenome && qGenome != tGenome) || coalActive;
        mapAdjacencies(ctx, tGenome, qGenome, doDupes, qTop, qseq, nqs, tseq[tSeqIdx], segs);
    }
    if (doDupes && qGenome != tGenome) chainReferenceParalogies(segs);
    std::vector<Seg> paraSet;
    for (const Seg &s : segs) if (s.inPara) paraSet.push_back(s);

    std::set<int64_t> queryCutSet, targetCutSet;
    targetCutSet.insert(absStart);
    targetCutSet.insert(absEnd);
    hal_block_t *prev = nullptr;
    std::vector<size_t> v1, v2, fragments;
    for (size_t x = nextAlive(segs, 0); x < segs.size(); x = nextAlive(segs, x + 1)) {
        // BlockMapper::extractSegment (liftover/impl/halBlockMapper.cpp:331-394)
        fragments.assign(1, x);
        v1.assign(1, x);
        size_t nx = nextAlive(segs, x + 1);
        while (nx < segs.size() && segs[v1.back()].qLo == segs[nx].qLo) { v1.push_back(nx); nx = nextAlive(segs, nx + 1); }
        while (nx < segs.size()) {
            v2.clear();
            while (nx < segs.size() && (v2.empty() || segs[v2.back()].qLo == segs[nx].qLo) && v2.size() < v1.size()) {
                v2.push_back(nx);
                nx = nextAlive(segs, nx + 1);
            }
            bool can = v1.size() == v2.size();
            for (size_t i = 0; i < v1.size() && can; ++i) {
                can = segs[v2[i]].qSeq == segs[x].qSeq && canMergeRightWith(segs[v1[i]], segs[v2[i]], queryCutSet, targetCutSet) &&
                      segs[v1[i]].inPara == segs[v2[i]].inPara;
            }
            if (!can) break;
            fragments.push_back(v2[0]);
            segs[v2[0]].alive = false; // (erased from the start set once the sweep of this element is over; nothing looks back)
            v1 = v2;
        }
        if (v1.size() > 1) queryCutSet.insert(std::max(segs[fragments.back()].qStartPos(), segs[fragments.back()].qEndPos()));
        // readBlock (halBlockViz.cpp:829-901)
        const Seg &fq = segs[fragments.front()], &lq = segs[fragments.back()];
        hal_block_t *cur = static_cast<hal_block_t *>(calloc(1, sizeof(hal_block_t)));
        if (results->mappedBlocks == nullptr) results->mappedBlocks = cur; else prev->next = cur;
        std::string qn = qseq[fq.qSeq].name;
        const size_t prefix = qn.find(qGenomeName + '.') != 0 ? 0 : qGenomeName.size() + 1;
        cur->qChrom = copyCString(qn.substr(prefix));
        cur->tStart = (hal_int_t)(std::min(fq.sLo, lq.sLo) - tseq[tSeqIdx].start);
        cur->qStart = (hal_int_t)(std::min(fq.qLo, lq.qLo) - qseq[fq.qSeq].start);
        const int64_t tEnd = std::max(fq.sHi(), lq.sHi()) - tseq[tSeqIdx].start;
        cur->size = (hal_int_t)(1 + tEnd - cur->tStart);
        cur->strand = fq.qRev ? '-' : '+';
        if (getSeq) { // DNA comes from the level-0 alignment, looked up by genome and sequence NAME (halBlockViz.cpp:867-899)
            auto locate = [&](const char *genomeName, const char *seqName, int &g, int64_t &start) {
                g = halgpu_genome_id(seqCtx, genomeName);
                if (g < 0) throw std::runtime_error("Unable to open genome " + std::string(genomeName) + " for DNA sequence extraction");
                const halgpu_seq *ss = nullptr;
                size_t nn = 0;
                halgpu_sequence_table(seqCtx, g, &ss, &nn);
                const int si = findSeq(ss, nn, seqName);
                if (si < 0) throw std::runtime_error("Unable to open sequence " + std::string(seqName) + " for DNA sequence extraction");
                start = ss[si].start;
            };
            int qg, tg;
            int64_t qs0, ts0;
            locate(qGenomeName.c_str(), qseq[fq.qSeq].name, qg, qs0);
            locate(halgpu_genome_name(ctx, tGenome), tseq[tSeqIdx].name, tg, ts0);
            std::string qd = dnaOf(seqCtx, qg, qs0 + cur->qStart, cur->size);
            const std::string td = dnaOf(seqCtx, tg, ts0 + cur->tStart, cur->size);
            if (cur->strand == '-') reverseComplement(qd);
            cur->qSequence = copyCString(qd);
            cur->tSequence = copyCString(td);
        }
        prev = cur;
    }
    if (!paraSet.empty() && doTargetDupes) {
        // the source sequence of the paralogy set's first element (the range lies within one reference sequence)
        results->targetDupeBlocks = processTargetDupes(paraSet, tseq[tSeqIdx].name, tseq[tSeqIdx].start);
    }
    return results;
}

} // namespace

extern "C" {

int halOpen(char *halFilePath, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    const int h = openLodOrHal(halFilePath, false, errStr);
    if (h < 0) return h;
    try { // the reference opens lazily too, but a missing / non-HAL file is better reported here
        alignmentFor(h, 0, true);
    } catch (std::exception &e) {
        gHandles.erase(h);
        handleError("openLodOrHal error: " + std::string(halFilePath) + ": " + e.what(), errStr);
        return -1;
    }
    return h;
}
int halOpenHalOrLod(char *lodFilePath, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    return openLodOrHal(lodFilePath, !isHalFile(lodFilePath), errStr);
}
int halOpenLOD(char *lodFilePath, char **errStr) { return halOpenHalOrLod(lodFilePath, errStr); } // deprecated spelling

int halClose(int handle, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    auto it = gHandles.find(handle);
    if (it == gHandles.end()) {
        handleError("halClose error closing handle " + std::to_string(handle) + ": not found", errStr);
        return -1;
    }
    for (auto &kv : it->second.lod) halgpu_close(kv.second.ctx);
    gHandles.erase(it);
    return 0;
}
int halCloseGenome(int, const char *, char **) { return 0; } // (MMapAlignment::closeGenome is a no-op as well)

void halFreeBlocks(struct hal_block_t *head) {
    while (head != nullptr) {
        hal_block_t *next = head->next;
        free(head->qChrom); free(head->qSequence); free(head->tSequence); free(head);
        head = next;
    }
}
void halFreeTargetDupeLists(struct hal_target_dupe_list_t *dupes) {
    while (dupes != nullptr) {
        hal_target_dupe_list_t *next = dupes->next;
        while (dupes->tRange != nullptr) {
            hal_target_range_t *rn = dupes->tRange->next;
            free(dupes->tRange);
            dupes->tRange = rn;
        }
        free(dupes->qChrom); free(dupes);
        dupes = next;
    }
}
void halFreeBlockResults(struct hal_block_results_t *results) {
    if (results != nullptr) {
        halFreeBlocks(results->mappedBlocks);
        halFreeTargetDupeLists(results->targetDupeBlocks);
        free(results);
    }
}
void halFreeSpeciesList(struct hal_species_t *s) {
    while (s != nullptr) {
        hal_species_t *next = s->next;
        free(s->name); free(s->parentName); free(s);
        s = next;
    }
}
void halFreeChromList(struct hal_chromosome_t *c) {
    while (c != nullptr) {
        hal_chromosome_t *next = c->next;
        free(c->name); free(c);
        c = next;
    }
}
void halFreeMetadataList(struct hal_metadata_t *m) {
    while (m != nullptr) {
        hal_metadata_t *next = m->next;
        free(m->key); free(m->value); free(m);
        m = next;
    }
}

struct hal_block_results_t *halGetBlocksInTargetRange(int halHandle, char *qSpecies, char *tSpecies, char *tChrom, hal_int_t tStart,
                                                      hal_int_t tEnd, hal_int_t tReversed, hal_seqmode_type_t seqMode,
                                                      hal_dup_type_t dupMode, int mapBackAdjacencies, const char *coalescenceLimitName,
                                                      char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try {
        const hal_int_t rangeLength = tEnd - tStart;
        if (rangeLength < 0) {
            handleError("halGetBlocksInTargetRange invalid query range [" + std::to_string(tStart) + "," + std::to_string(tEnd) + ")", errStr);
            return nullptr;
        }
        if (tReversed != 0 && mapBackAdjacencies != 0) {
            handleError("halGetBlocksInTargetRange tReversed can only be set when mapBackAdjacencies is 0", errStr);
            return nullptr;
        }
        if (tReversed != 0 && dupMode == HAL_QUERY_AND_TARGET_DUPS) {
            handleError("tReversed cannot be set in conjunction with dupMode=HAL_QUERY_AND_TARGET_DUPS", errStr);
            return nullptr;
        }
        bool getSeq; // halBlockViz.cpp:268-279
        switch (seqMode) {
        case HAL_NO_SEQUENCE: getSeq = false; break;
        case HAL_FORCE_LOD0_SEQUENCE: getSeq = true; break;
        case HAL_LOD0_SEQUENCE:
        default: getSeq = isLod0(halHandle, (uint64_t)rangeLength);
        }
        halgpu_ctx *ctx = alignmentFor(halHandle, (uint64_t)rangeLength, getSeq);
        int q = -1, t = -1, ts = -1;
        const halgpu_seq *tseq = nullptr;
        size_t nts = 0;
        auto checkGenomes = [&](halgpu_ctx *c, const std::string &chrom) {
            q = halgpu_genome_id(c, qSpecies);
            if (q < 0) throw std::runtime_error("Query species " + std::string(qSpecies) + " not found in alignment with handle " + std::to_string(halHandle));
            t = halgpu_genome_id(c, tSpecies);
            if (t < 0) throw std::runtime_error("Reference species " + std::string(tSpecies) + " not found in alignment with handle " + std::to_string(halHandle));
            halgpu_sequence_table(c, t, &tseq, &nts);
            ts = findSeq(tseq, nts, chrom.c_str());
            if (ts < 0) throw std::runtime_error("Unable to locate sequence " + chrom + " in genome " + tSpecies);
        };
        checkGenomes(ctx, tChrom);
        const int64_t myEnd = tEnd > 0 ? tEnd : tseq[ts].length;
        const int64_t absStart = tseq[ts].start + tStart, absEnd = tseq[ts].start + myEnd - 1;
        if (absStart > absEnd) {
            handleError("halGetBlocksInTargetRange invalid range", errStr);
            return nullptr;
        }
        // (MMapSequence::getEndPosition() is start + length, api/mmap_impl/mmapSequence.h:50-52: one base past the end is let through by
        //  the reference and then fails inside the iterator; reject it here)
        if (absEnd > tseq[ts].start + tseq[ts].length - 1) {
            handleError("halGetBlocksInTargetRange target end position outside of target sequence", errStr);
            return nullptr;
        }
        if (tEnd == 0) { // the query length is known now: a proper level-of-detail choice (:299-306)
            ctx = alignmentFor(halHandle, (uint64_t)(absEnd - absStart), false);
            checkGenomes(ctx, tChrom);
        }
        halgpu_ctx *seqCtx = getSeq ? alignmentFor(halHandle, (uint64_t)(absEnd - absStart), true) : nullptr;
        return readBlocks(ctx, seqCtx, t, ts, absStart, absEnd, tReversed != 0, q, getSeq, dupMode != HAL_NO_DUPS, dupMode == HAL_QUERY_AND_TARGET_DUPS,
                          mapBackAdjacencies != 0, coalescenceLimitName);
    } catch (std::exception &e) {
        handleError("halGetBlocksInTargetRange error reading blocks: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

struct hal_block_results_t *halGetBlocksInTargetRange_filterByChrom(int halHandle, char *qSpecies, char *tSpecies, char *tChrom,
                                                                    hal_int_t tStart, hal_int_t tEnd, hal_int_t tReversed,
                                                                    hal_seqmode_type_t seqMode, hal_dup_type_t dupMode,
                                                                    int mapBackAdjacencies, char *qChrom,
                                                                    const char *coalescenceLimitName, char **errStr) {
    hal_block_results_t *r = halGetBlocksInTargetRange(halHandle, qSpecies, tSpecies, tChrom, tStart, tEnd, tReversed, seqMode, dupMode,
                                                       mapBackAdjacencies, coalescenceLimitName, errStr);
    if (r == nullptr) return nullptr;
    hal_block_t **bp = &r->mappedBlocks; // keep only the blocks / dupe lists on qChrom (halBlockViz.cpp:349-404)
    while (*bp != nullptr) {
        if (strcmp((*bp)->qChrom, qChrom) != 0) {
            hal_block_t *dead = *bp;
            *bp = dead->next;
            dead->next = nullptr;
            halFreeBlocks(dead);
        } else {
            bp = &(*bp)->next;
        }
    }
    hal_target_dupe_list_t **dp = &r->targetDupeBlocks;
    while (*dp != nullptr) {
        if (strcmp((*dp)->qChrom, qChrom) != 0) {
            hal_target_dupe_list_t *dead = *dp;
            *dp = dead->next;
            dead->next = nullptr;
            halFreeTargetDupeLists(dead);
        } else {
            dp = &(*dp)->next;
        }
    }
    return r;
}

hal_int_t halGetMaf(FILE *outFile, int halHandle, struct hal_species_t *qSpeciesNames, char *tSpecies, char *tChrom, hal_int_t tStart,
                    hal_int_t tEnd, int maxRefGap, int maxBlockLength, int doDupes, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    hal_int_t numBytes = 0;
    try { // halBlockViz.cpp:407-473
        if (tEnd - tStart < 0) {
            handleError("halGetMaf invalid query range [" + std::to_string(tStart) + "," + std::to_string(tEnd) + ")", errStr);
            return -1;
        }
        if (maxRefGap != 0) {
            handleError("halGetMaf: maxRefGap > 0 (gapped column iterators) is not implemented in the GPU build", errStr);
            return -1;
        }
        halgpu_ctx *ctx = alignmentFor(halHandle, 0, true);
        const int t = halgpu_genome_id(ctx, tSpecies);
        std::set<int> qSet;
        const halgpu_seq *tseq = nullptr;
        size_t nts = 0;
        for (hal_species_t *q = qSpeciesNames; q != nullptr; q = q->next) { // checkGenomes per query species
            const int qi = halgpu_genome_id(ctx, q->name);
            if (qi < 0) throw std::runtime_error("Query species " + std::string(q->name) + " not found in alignment with handle " + std::to_string(halHandle));
            if (t < 0) throw std::runtime_error("Reference species " + std::string(tSpecies) + " not found in alignment with handle " + std::to_string(halHandle));
            halgpu_sequence_table(ctx, t, &tseq, &nts);
            if (findSeq(tseq, nts, tChrom) < 0) throw std::runtime_error("Unable to locate sequence " + std::string(tChrom) + " in genome " + tSpecies);
            qSet.insert(qi);
        }
        if (t < 0) throw std::runtime_error("Reference species " + std::string(tSpecies) + " not found in alignment with handle " + std::to_string(halHandle));
        halgpu_sequence_table(ctx, t, &tseq, &nts);
        const int ts = findSeq(tseq, nts, tChrom);
        if (ts < 0) throw std::runtime_error("Unable to locate sequence " + std::string(tChrom) + " in genome " + tSpecies);
        const int64_t myEnd = tEnd > 0 ? tEnd : tseq[ts].length;
        const int64_t absStart = tseq[ts].start + tStart, absEnd = tseq[ts].start + myEnd - 1;
        if (absStart > absEnd) {
            handleError("halGetMaf invalid range", errStr);
            return -1;
        }
        if (absEnd > tseq[ts].start + tseq[ts].length) { // MMapSequence::getEndPosition() == start + length
            handleError("halGetMaf target end position outside of target sequence", errStr);
            return -1;
        }
        std::ostringstream mafBuffer;
        halgpu::GpuMafExport mafExport(ctx);
        mafExport.setNoDupes(doDupes == 0);
        mafExport.setUcscNames(true);
        mafExport.setMaxBlockLengthRaw((int64_t)maxBlockLength);
        // sic: the reference hands convertSequence the ABSOLUTE start (:452), which only equals the sequence-relative one
        // it expects for the first sequence of a genome; kept
        mafExport.convertSequence(mafBuffer, t, ts, absStart, (uint64_t)(1 + absEnd - absStart), std::vector<int>(qSet.begin(), qSet.end()));
        const std::string buf = mafBuffer.str();
        if (!buf.empty()) numBytes = (hal_int_t)fwrite(buf.c_str(), buf.length(), sizeof(char), outFile); // (fwrite's ITEM count, i.e. 1)
    } catch (std::exception &e) {
        handleError("halGetMaf error writing MAF blocks: " + std::string(e.what()), errStr);
        return -1;
    }
    return numBytes;
}
hal_int_t halGetMAF(FILE *outFile, int halHandle, struct hal_species_t *qSpeciesNames, char *tSpecies, char *tChrom, hal_int_t tStart,
                    hal_int_t tEnd, int doDupes, char **errStr) { // deprecated spelling (halBlockViz.cpp:475-479)
    return halGetMaf(outFile, halHandle, qSpeciesNames, tSpecies, tChrom, tStart, tEnd, 0, 0, doDupes, errStr);
}

struct hal_species_t *halGetSpecies(int halHandle, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try {
        halgpu_ctx *ctx = ctxOf(halHandle);
        const std::map<std::string, double> bl = branchLengths(halgpu_newick(ctx));
        hal_species_t *head = nullptr, *prev = nullptr;
        const int n = halgpu_num_genomes(ctx);
        int root = -1;
        for (int gidx = 0; gidx < n; ++gidx) if (halgpu_genome_parent(ctx, gidx) < 0) root = gidx;
        if (root < 0) return nullptr;
        std::deque<int> bfQueue(1, root); // breadth first, children in newick order (halBlockViz.cpp:490-520)
        while (!bfQueue.empty()) {
            const int gi = bfQueue.back();
            bfQueue.pop_back();
            hal_species_t *cur = speciesOf(ctx, gi, bl);
            if (head == nullptr) head = cur; else prev->next = cur;
            prev = cur;
            for (int k = 0; k < halgpu_genome_num_children(ctx, gi); ++k) bfQueue.push_front(halgpu_genome_child(ctx, gi, k));
        }
        return head;
    } catch (std::exception &e) {
        handleError("halGetSpecies: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

struct hal_species_t *halGetPossibleCoalescenceLimits(int halHandle, const char *qSpecies, const char *tSpecies, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try {
        halgpu_ctx *ctx = ctxOf(halHandle);
        const int q = halgpu_genome_id(ctx, qSpecies), t = halgpu_genome_id(ctx, tSpecies);
        if (q < 0 || t < 0) throw std::runtime_error("genome not found");
        const std::map<std::string, double> bl = branchLengths(halgpu_newick(ctx));
        hal_species_t *head = nullptr, *prev = nullptr;
        for (int gi = halgpu_mrca(ctx, q, t); gi >= 0; gi = halgpu_genome_parent(ctx, gi)) { // the MRCA and all its ancestors
            hal_species_t *cur = speciesOf(ctx, gi, bl);
            if (head == nullptr) head = cur; else prev->next = cur;
            prev = cur;
        }
        return head;
    } catch (std::exception &e) {
        handleError("halGetPossibleCoalescenceLimits: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

struct hal_chromosome_t *halGetChroms(int halHandle, char *speciesName, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try {
        halgpu_ctx *ctx = ctxOf(halHandle);
        const int gi = halgpu_genome_id(ctx, speciesName);
        if (gi < 0) {
            handleError("halGetChroms: species with name " + std::string(speciesName) + " not found in alignment with handle " + std::to_string(halHandle), errStr);
            return nullptr;
        }
        const halgpu_seq *seqs = nullptr;
        size_t ns = 0;
        halgpu_sequence_table(ctx, gi, &seqs, &ns);
        hal_chromosome_t *head = nullptr, *prev = nullptr;
        for (size_t i = 0; i < ns; ++i) {
            hal_chromosome_t *cur = static_cast<hal_chromosome_t *>(calloc(1, sizeof(hal_chromosome_t)));
            cur->name = copyCString(seqs[i].name);
            cur->length = (hal_int_t)seqs[i].length;
            if (head == nullptr) head = cur; else prev->next = cur;
            prev = cur;
        }
        return head;
    } catch (std::exception &e) {
        handleError("halGetChroms: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

char *halGetDna(int halHandle, char *speciesName, char *chromName, hal_int_t start, hal_int_t end, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try {
        halgpu_ctx *ctx = alignmentFor(halHandle, 0, true);
        const int gi = halgpu_genome_id(ctx, speciesName);
        if (gi < 0) {
            handleError("halGetChroms: species with name " + std::string(speciesName) + " not found in alignment with handle " + std::to_string(halHandle), errStr);
            return nullptr;
        }
        const halgpu_seq *seqs = nullptr;
        size_t ns = 0;
        halgpu_sequence_table(ctx, gi, &seqs, &ns);
        const int si = findSeq(seqs, ns, chromName);
        if (si < 0) {
            handleError("halGetDna: chromosome with name " + std::string(chromName) + " not found in species " + speciesName, errStr);
            return nullptr;
        }
        if (start > end || end > (hal_int_t)seqs[si].length) {
            handleError("halGetDna: specified range [" + std::to_string(start) + "," + std::to_string(end) + ") is invalid for chromsome " + chromName +
                            " in species " + speciesName + " which is of length " + std::to_string(seqs[si].length), errStr);
            return nullptr;
        }
        return copyCString(dnaOf(ctx, gi, seqs[si].start + start, end - start));
    } catch (std::exception &e) {
        handleError("halGetDna: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

hal_int_t halGetMaxLODQueryLength(int halHandle, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    if (gHandles.find(halHandle) == gHandles.end()) {
        handleError("halGetMaxLODQueryLength error getting Max LOD Query Length.  handle " + std::to_string(halHandle) + ": not found", errStr);
        return -1;
    }
    return (hal_int_t)(gHandles[halHandle].maxLodLowerBound - 1); // LodManager::getMaxQueryLength()
}

struct hal_metadata_t *halGetGenomeMetadata(int halHandle, const char *genomeName, char **errStr) {
    std::lock_guard<std::mutex> g(gLock);
    try { // halBlockViz.cpp:1177-1213
        halgpu_ctx *ctx = ctxOf(halHandle);
        const int gi = halgpu_genome_id(ctx, genomeName);
        if (gi < 0) throw std::runtime_error("Genome " + std::string(genomeName) + " not found in alignment");
        hal_metadata_t *head = nullptr, *prev = nullptr;
        const char *k = nullptr, *v = nullptr;
        for (size_t i = 0; halgpu_genome_metadata(ctx, gi, i, &k, &v) == 0; ++i) {
            hal_metadata_t *cur = static_cast<hal_metadata_t *>(calloc(1, sizeof(hal_metadata_t)));
            cur->key = copyCString(k);
            cur->value = copyCString(v);
            if (prev != nullptr) prev->next = cur; else head = cur;
            prev = cur;
        }
        return head;
    } catch (std::exception &e) {
        handleError("halGetGenomeMetadata: " + std::string(e.what()), errStr);
        return nullptr;
    }
}

} // extern "C"
