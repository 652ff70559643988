/* TEST INFRASTRUCTURE ONLY -- CPU oracle: halWiggleLiftover restated, text in -> text out.
 *
 * Follows the reference line by line:
 *   WiggleScanner::scan / scanHeader / scanLine        liftover/impl/halWiggleScanner.cpp:39-167
 *     (incl. its quirks: a header needs text after its first token; variableStep positions are used as read,
 *      i.e. 0-based -- :143 decrements _start instead of _first; span applies when > 1)
 *   WiggleLiftover::visitHeader / visitLine / mapSegment / mapFragments / write   halWiggleLiftover.cpp:63-198
 *     (batches = lines inside the "current segment"; "Coordinate out of order" only inside a batch; every batch
 *      range is mapped like one '+' BED interval -- halMapSegment per source segment, extractSegment runs -- and
 *      every run's bases are assigned with the _cvIdx cursor; value = max(lifted, existing-or-0.0))
 *   WiggleLoader (--append preload)                     halWiggleLoader.cpp:24-48
 *   WiggleTiles get/set/exists                          liftover/inc/halWiggleTiles.h:96-137
 * The reference passes the src/tgt SPANNING tree as genomesOnPath (halWiggleLiftover.cpp:51-54), so
 * mapRecursiveDown (api/impl/halSegmentMapper.cpp:208-224) descends from the MRCA into the FIRST child that is on
 * that set: when the source-side child precedes the target-side child it walks back down to the source genome and
 * throws "Could not find correct child that leads from <src> to <tgt>" as soon as anything maps up to the MRCA.
 * Restated here (wrongTurn) so the oracle and the reference agree on which inputs fail.
 */
#include "wiggle.h"
#include <algorithm>
#include <sstream>

namespace oracle {

namespace {

struct CoordVal {
    int64_t first, last;
    double val;
};

struct WigScan { /* WiggleScanner state */
    bool fixedStep = false;
    bool haveHeader = false;
    std::string sequenceName;
    int64_t start = 0, step = 0, span = -1, offset = 0;
    int64_t first = 0, last = 0;
    double value = 0;

    bool scanHeader(const std::string &line) {
        std::stringstream ss(line);
        std::string buf;
        ss >> buf;
        if (ss.good() && buf == "variableStep") {
            fixedStep = false;
            ss >> buf;
            if (!ss || buf.length() <= 6 || buf.substr(0, 6) != "chrom=") throw std::runtime_error("Error parsing chrom in variableStep header");
            sequenceName = buf.substr(6);
            parseSpan(ss);
            return true;
        }
        if (ss.good() && buf == "fixedStep") {
            fixedStep = true;
            offset = 0;
            ss >> buf;
            if (!ss || buf.length() <= 6 || buf.substr(0, 6) != "chrom=") throw std::runtime_error("Error parsing chrom in fixedStep header");
            sequenceName = buf.substr(6);
            ss >> buf;
            if (!ss || buf.length() <= 6 || buf.substr(0, 6) != "start=") throw std::runtime_error("Error parsing start in fixedStep header");
            {
                std::stringstream s1(buf.substr(6));
                s1 >> start;
                if (!s1) throw std::runtime_error("Error parsing start in fixedStep header");
            }
            --start; /* 0-based internally */
            ss >> buf;
            if (!ss || buf.length() <= 5 || buf.substr(0, 5) != "step=") throw std::runtime_error("Error parsing step in fixedStep header");
            {
                std::stringstream s2(buf.substr(5));
                s2 >> step;
                if (!s2) throw std::runtime_error("Error parsing step in fixedStep header");
            }
            parseSpan(ss);
            return true;
        }
        return false;
    }
    void parseSpan(std::stringstream &ss) {
        std::string buf;
        ss >> buf;
        if (!ss || buf.length() <= 5 || buf.substr(0, 5) != "span=") {
            span = -1;
        } else {
            std::stringstream s1(buf.substr(5));
            s1 >> span;
            if (!s1) span = -1;
        }
    }
    void scanLine(const std::string &line) {
        std::stringstream ss(line);
        if (fixedStep) {
            first = start + offset * step;
            ++offset;
        } else {
            ss >> first;
            if (!ss) throw std::runtime_error("Error parsing position for " + sequenceName);
            --start; /* sic: the reference decrements _start, so variableStep positions stay as read */
        }
        ss >> value;
        if (!ss) {
            std::stringstream m;
            m << "Error parsing value for " << sequenceName << " pos " << start;
            throw std::runtime_error(m.str());
        }
        last = first;
        if (span > 1) last += span - 1;
    }
};

/* calls visitor(kind, scan): kind 0 header, 1 data line, 2 EOF.  Exceptions of kinds 0/1 get the line number suffix. */
template <class V> void scanWig(const std::string &text, V &&visit) {
    WigScan sc;
    size_t p = 0, lineNumber = 0;
    auto skipWs = [&]() { while (p < text.size() && std::isspace((unsigned char)text[p])) ++p; };
    try {
        skipWs();
        while (p < text.size()) {
            ++lineNumber;
            size_t nl = text.find('\n', p);
            std::string line = text.substr(p, nl == std::string::npos ? std::string::npos : nl - p);
            p = nl == std::string::npos ? text.size() : nl + 1;
            if (sc.scanHeader(line)) {
                sc.haveHeader = true;
                visit(0, sc);
            } else {
                if (!sc.haveHeader) throw std::runtime_error("Missing Wig header"); /* (the reference reads an uninitialised _fixedStep first) */
                sc.scanLine(line);
                visit(1, sc);
            }
            skipWs();
        }
    } catch (std::exception &e) {
        throw std::runtime_error(std::string(e.what()) + " in input wiggle line " + std::to_string(lineNumber));
    }
    visit(2, sc);
}

const SeqView *findSeq(const GenomeView &g, const std::string &name) {
    for (const SeqView &s : g.seqs) if (s.name == name) return &s;
    return nullptr;
}

} // namespace

std::string wiggleLiftover(const HalView &v, int src, int tgt, bool dupes, const std::string &inText, const std::string *preloadText,
                           bool correctPath) {
    const GenomeView &S = v.genomes[src], &T = v.genomes[tgt];
    std::vector<double> vals((size_t)T.len, 0.0);
    std::vector<char> exists((size_t)T.len, 0);
    if (preloadText) { /* WiggleLoader::visitLine: plain set(), later lines overwrite */
        const SeqView *seq = nullptr;
        scanWig(*preloadText, [&](int kind, WigScan &sc) {
            if (kind == 0) {
                seq = findSeq(T, sc.sequenceName);
                if (!seq) throw std::runtime_error("Sequence " + sc.sequenceName + " not found in genome " + T.name);
            } else if (kind == 1) {
                for (int64_t p = sc.first + seq->start; p <= sc.last + seq->start; ++p) {
                    vals.at((size_t)p) = sc.value;
                    exists[(size_t)p] = 1;
                }
            }
        });
    }
    const Plan plan = makePlan(v, src, tgt);
    /* which child of the MRCA does mapRecursiveDown pick?  the first one on the spanning set of {src, tgt} */
    bool wrongTurn = false;
    if (!correctPath && plan.mrca != tgt && plan.mrca != src) {
        const int srcSide = plan.up[plan.up.size() - 2], tgtSide = plan.down[1];
        wrongTurn = v.genomes[srcSide].slot < v.genomes[tgtSide].slot;
    }
    /* the wrong turn leads from the MRCA straight back down to the source genome; the exception is raised there, when the
     * list that arrives is not empty (without dupes the way down can lose everything: only canonical copies are followed) */
    Plan backPlan = makePlan(v, src, plan.mrca);
    backPlan.tgt = src;
    backPlan.down.assign(backPlan.up.rbegin(), backPlan.up.rend());
    const bool srcTop = S.numTop > 0;
    const int64_t N = srcTop ? S.numTop : S.numBot;
    auto sstart = [&](int64_t i) { return srcTop ? S.tStart(i) : S.bStart(i); };
    auto segOf = [&](int64_t pos) {
        int64_t lo = 0, hi = N - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) / 2;
            if (sstart(mid) <= pos) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    std::vector<CoordVal> cvals;
    int64_t curSeg = 0; /* array index of _segment */
    const SeqView *srcSeq = nullptr;

    auto mapSegment = [&]() {
        if (cvals.empty()) return;
        const int64_t first = cvals[0].first, last = cvals.back().last;
        if (first < 0 || last >= S.len) throw std::runtime_error("wiggle coordinate outside the source genome");
        std::vector<OutLine> lines;
        std::vector<Frag> frags;
        std::vector<size_t> runSizes;
        if (wrongTurn) {
            liftInterval(v, backPlan, dupes, first, last, '+', lines, nullptr);
            if (!lines.empty()) throw std::runtime_error("Could not find correct child that leads from " + S.name + " to " + T.name);
        } else {
            liftInterval(v, plan, dupes, first, last, '+', lines, nullptr, &frags, &runSizes);
        }
        size_t at = 0;
        for (size_t rs : runSizes) { /* mapFragments, once per extractSegment run */
            std::vector<Frag> run(frags.begin() + at, frags.begin() + at + rs);
            at += rs;
            std::sort(run.begin(), run.end(), [](const Frag &a, const Frag &b) {
                if (a.sLo != b.sLo) return a.sLo < b.sLo;
                if (a.sHi != b.sHi) return a.sHi < b.sHi;
                if (a.tLo != b.tLo) return a.tLo < b.tLo;
                return a.tHi < b.tHi;
            });
            size_t cvIdx = 0;
            for (size_t i = 0; i < run.size() && cvIdx < cvals.size(); ++i) {
                const Frag &f = run[i];
                const int64_t len = f.sHi - f.sLo + 1;
                for (int64_t j = 0; j < len && cvIdx < cvals.size(); ++j) {
                    const int64_t pos = f.sLo + j;
                    while (cvIdx < cvals.size() && cvals[cvIdx].last < pos) ++cvIdx;
                    if (cvIdx < cvals.size() && pos >= cvals[cvIdx].first && pos <= cvals[cvIdx].last) {
                        const int64_t mpos = f.tRev ? f.tHi - j : f.tLo + j;
                        const double cur = vals[(size_t)mpos]; /* get(): the default 0.0 until set */
                        vals[(size_t)mpos] = std::max(cvals[cvIdx].val, cur);
                        exists[(size_t)mpos] = 1;
                    }
                }
            }
        }
        /* where the reference's _segment ends up: past the segment that contains `last`, or on it when the range
         * stopped short of its end (SegmentIterator::toRight, api/impl/halSegmentIterator.cpp:208-238) */
        const int64_t s = segOf(last);
        curSeg = (last == sstart(s + 1) - 1) ? s + 1 : s;
        cvals.clear();
    };

    scanWig(inText, [&](int kind, WigScan &sc) {
        if (kind == 0) {
            mapSegment();
            srcSeq = findSeq(S, sc.sequenceName);
            if (!srcSeq) throw std::runtime_error("Sequence " + sc.sequenceName + " not found in genome " + S.name);
        } else if (kind == 1) {
            if (curSeg >= N) curSeg = 0;
            const int64_t absFirst = sc.first + srcSeq->start, absLast = sc.last + srcSeq->start;
            if (absFirst < sstart(curSeg) || absLast > sstart(curSeg + 1) - 1) mapSegment();
            if (!cvals.empty() && cvals.back().last >= absFirst) throw std::runtime_error("Coordinate out of order");
            cvals.push_back(CoordVal{absFirst, absLast, sc.value});
        } else {
            mapSegment();
        }
    });

    /* WiggleLiftover::write */
    std::ostringstream out;
    int outSeq = -1;
    bool needHeader = true;
    int64_t prevPos = -1;
    for (int64_t pos = 0; pos < T.len; ++pos) {
        if (!exists[(size_t)pos]) continue;
        /* MMapSequence::getEndPosition() is start + length, ONE PAST the last base (api/mmap_impl/mmapSequence.h:50-52; the
         * interface documents start + len - 1, api/inc/halSequence.h:76): a run that continues contiguously into the next
         * sequence gets no header there, and a header printed at exactly that position still names the previous sequence */
        if (outSeq < 0 || pos < T.seqs[outSeq].start || pos > T.seqs[outSeq].start + T.seqs[outSeq].length) {
            outSeq = T.seqOf(pos);
            needHeader = true;
        } else if (pos != prevPos + 1) {
            needHeader = true;
        }
        if (needHeader) {
            out << "fixedStep" << "\tchrom=" << T.seqs[outSeq].name << "\tstart=" << (1 + pos - T.seqs[outSeq].start) << "\tstep=1\n";
            needHeader = false;
        }
        out << vals[(size_t)pos] << '\n';
        prevPos = pos;
    }
    return out.str();
}

} // namespace oracle
