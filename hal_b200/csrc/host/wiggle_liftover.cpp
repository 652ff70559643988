#include "wiggle_liftover.hpp"
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <map>
#include <ostream>
#include <stdexcept>
#include <thread>

namespace halgpu {

const double GpuWiggleLiftover::DefaultValue = 0.0;

namespace {

inline bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// ---- the extraction operators WiggleScanner relies on, restated over character ranges ----
// `ss >> std::string`: skip blanks, take the run of non-blanks.  Returns false when nothing is left (the stream fails).
inline bool nextToken(const char *&p, const char *e, const char *&tb, const char *&te) {
    while (p < e && isSpace(*p)) ++p;
    if (p >= e) return false;
    tb = p;
    while (p < e && !isSpace(*p)) ++p;
    te = p;
    return true;
}
// `ss >> int64`: skip blanks, optional sign, digits; trailing text is left alone; no digits or overflow fail
inline bool extractInt(const char *&p, const char *e, int64_t &v) {
    while (p < e && isSpace(*p)) ++p;
    const char *b = p;
    if (b < e && *b == '+') ++b;
    auto r = std::from_chars(b, e, v);
    if (r.ec != std::errc() || (b > p && b < e && *b == '-')) return false;
    p = r.ptr;
    return true;
}
// `ss >> double`: num_get gathers [sign] digits [. digits] [e|E [sign] digits] and converts the gathered text, which must
// be consumed completely ("1e", ".", "-" fail; "inf"/"nan" are not gathered at all; out of range fails)
inline bool extractDouble(const char *&p, const char *e, double &v) {
    while (p < e && isSpace(*p)) ++p;
    const char *b = p, *q = p;
    if (q < e && (*q == '+' || *q == '-')) ++q;
    while (q < e && *q >= '0' && *q <= '9') ++q;
    if (q < e && *q == '.') {
        ++q;
        while (q < e && *q >= '0' && *q <= '9') ++q;
    }
    if (q < e && (*q == 'e' || *q == 'E')) {
        ++q;
        if (q < e && (*q == '+' || *q == '-')) ++q;
        while (q < e && *q >= '0' && *q <= '9') ++q;
    }
    const char *nb = (b < q && *b == '+') ? b + 1 : b;
    auto r = std::from_chars(nb, q, v);
    if (r.ec != std::errc() || r.ptr != q || nb == q) return false;
    p = q;
    return true;
}
inline bool hasPrefix(const char *tb, const char *te, const char *pre, size_t n) { return (size_t)(te - tb) > n && std::memcmp(tb, pre, n) == 0; }

struct WigScanner { // WiggleScanner's members (liftover/inc/halWiggleScanner.h:44-57)
    bool fixedStep = false, haveHeader = false;
    std::string sequenceName;
    int64_t start = 0, step = 0, span = -1, offset = 0;
    int64_t first = 0, last = 0;
    double value = 0;

    void scanSpan(const char *&p, const char *e) { // optional trailing span=N (halWiggleScanner.cpp:84-93, 125-134)
        const char *tb, *te;
        span = -1;
        if (nextToken(p, e, tb, te) && hasPrefix(tb, te, "span=", 5)) {
            const char *q = tb + 5;
            if (!extractInt(q, te, span)) span = -1;
        }
    }
    // WiggleScanner::scanHeader (halWiggleScanner.cpp:71-139)
    bool scanHeader(const char *p, const char *e) {
        const char *tb, *te;
        if (!nextToken(p, e, tb, te) || p >= e) return false; // the reference requires ss.good() after the first token
        const size_t n = (size_t)(te - tb);
        const bool var = n == 12 && std::memcmp(tb, "variableStep", 12) == 0;
        const bool fix = n == 9 && std::memcmp(tb, "fixedStep", 9) == 0;
        if (!var && !fix) return false;
        if (var) {
            fixedStep = false;
            if (!nextToken(p, e, tb, te) || !hasPrefix(tb, te, "chrom=", 6)) throw std::runtime_error("Error parsing chrom in variableStep header");
            sequenceName.assign(tb + 6, te);
            scanSpan(p, e);
            return true;
        }
        fixedStep = true;
        offset = 0;
        if (!nextToken(p, e, tb, te) || !hasPrefix(tb, te, "chrom=", 6)) throw std::runtime_error("Error parsing chrom in fixedStep header");
        sequenceName.assign(tb + 6, te);
        if (!nextToken(p, e, tb, te) || !hasPrefix(tb, te, "start=", 6)) throw std::runtime_error("Error parsing start in fixedStep header");
        {
            const char *q = tb + 6;
            if (!extractInt(q, te, start)) throw std::runtime_error("Error parsing start in fixedStep header");
        }
        --start; // store internally in 0-based coordinates
        if (!nextToken(p, e, tb, te) || !hasPrefix(tb, te, "step=", 5)) throw std::runtime_error("Error parsing step in fixedStep header");
        {
            const char *q = tb + 5;
            if (!extractInt(q, te, step)) throw std::runtime_error("Error parsing step in fixedStep header");
        }
        scanSpan(p, e);
        return true;
    }
    // WiggleScanner::scanLine (halWiggleScanner.cpp:141-167)
    void scanLine(const char *p, const char *e) {
        if (fixedStep) {
            first = start + offset * step;
            ++offset;
        } else {
            if (!extractInt(p, e, first)) throw std::runtime_error("Error parsing position for " + sequenceName);
            --start; // sic (:156): the reference decrements _start, not _first, so variableStep positions are used as read
        }
        if (!extractDouble(p, e, value)) {
            throw std::runtime_error("Error parsing value for " + sequenceName + " pos " + std::to_string(start));
        }
        last = first;
        if (span > 1) last += span - 1;
    }
};

// WiggleScanner::scan (halWiggleScanner.cpp:39-69): visit(0) after a header, visit(1) after a data line, visit(2) at EOF.
// Blank lines and leading blanks are skipped between lines and do not count as lines.
template <class V> size_t scanWiggle(const char *text, size_t n, WigScanner &sc, V &&visit) {
    const char *p = text, *const e = text + n;
    size_t lineNumber = 0;
    try {
        while (p < e && isSpace(*p)) ++p;
        while (p < e) {
            ++lineNumber;
            const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(e - p)));
            const char *le = nl ? nl : e;
            // fast test: a data line starts with a digit, sign or dot far more often than with 'f'/'v'
            if ((*p == 'f' || *p == 'v') && sc.scanHeader(p, le)) {
                sc.haveHeader = true;
                visit(0);
            } else {
                if (!sc.haveHeader) throw std::runtime_error("Missing Wig header"); // (the reference reads an uninitialised _fixedStep here)
                sc.scanLine(p, le);
                visit(1);
            }
            p = nl ? nl + 1 : e;
            while (p < e && isSpace(*p)) ++p;
        }
    } catch (std::exception &ex) {
        throw std::runtime_error(std::string(ex.what()) + " in input wiggle line " + std::to_string(lineNumber));
    }
    visit(2);
    return lineNumber;
}


// Multi-threaded pre-parse of the data lines (SURVEY.md 8(f) rank 1 applied to wiggle text).  The lines are cut into chunks,
// every thread classifies its lines -- header candidate / "value" / "position value" -- and converts the numbers in a strict
// dialect (whole tokens, nothing else on the line); headers are then parsed by the very same WigScanner::scanHeader and the
// visitors are called in input order with the state scanLine would have produced.  Anything the strict pre-parse does not
// recognise (or a line form that does not fit its section) makes it return false BEFORE any visitor ran, and the caller
// falls back to scanWiggle(), which owns the reference's tolerant parsing and its messages.
struct FastWigChunk {
    std::vector<uint8_t> kind;    // per non-blank line: 0 header candidate, 1 "value", 2 "position value"
    std::vector<double> value;    // per line of kind 1 / 2 (kind 0: unused slot)
    std::vector<int64_t> pos;     // per line of kind 2
    std::vector<std::pair<const char *, const char *>> header; // text of the kind-0 lines
    bool bad = false;
};

inline bool strictDouble(const char *b, const char *e, double &v) { // the whole token, in the grammar `ss >> double` gathers
    const char *q = b;
    return b < e && !isSpace(*b) && extractDouble(q, e, v) && q == e;
}
inline bool strictInt64(const char *b, const char *e, int64_t &v) {
    const char *q = b;
    return b < e && !isSpace(*b) && extractInt(q, e, v) && q == e;
}

template <class V> bool fastScanWiggle(const char *text, size_t n, unsigned nThreads, WigScanner &sc, size_t &linesOut, V &&visit) {
    size_t grain = (size_t)1 << 20;
    if (const char *gs = std::getenv("HALGPU_WIG_GRAIN")) grain = (size_t)std::max(1L, std::atol(gs)); // test hook
    nThreads = std::max(1u, std::min<unsigned>(nThreads, (unsigned)(n / grain) + 1));
    std::vector<size_t> cut(nThreads + 1, n);
    cut[0] = 0;
    for (unsigned t = 1; t < nThreads; ++t) {
        const size_t c = std::max(cut[t - 1], n * t / nThreads);
        const char *nl = c < n ? static_cast<const char *>(std::memchr(text + c, '\n', n - c)) : nullptr;
        cut[t] = nl ? (size_t)(nl - text) + 1 : n;
    }
    std::vector<FastWigChunk> chunks(nThreads);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nThreads; ++t) {
        th.emplace_back([&, t] {
            FastWigChunk &C = chunks[t];
            const char *p = text + cut[t], *const e = text + cut[t + 1];
            const size_t guess = (size_t)(e - p) / 6 + 16;
            C.kind.reserve(guess);
            C.value.reserve(guess);
            while (true) {
                while (p < e && isSpace(*p)) ++p;
                if (p >= e) break;
                const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(e - p)));
                const char *le = nl ? nl : e;
                if (*p == 'f' || *p == 'v') {
                    C.kind.push_back(0);
                    C.value.push_back(0);
                    C.header.emplace_back(p, le);
                } else {
                    const char *t1 = p;
                    while (t1 < le && !isSpace(*t1)) ++t1;
                    const char *q = t1;
                    while (q < le && isSpace(*q)) ++q;
                    double v = 0;
                    if (q == le) { // one token
                        if (!strictDouble(p, t1, v)) { C.bad = true; return; }
                        C.kind.push_back(1);
                        C.value.push_back(v);
                    } else { // two tokens, then only blanks
                        const char *t2 = q;
                        while (t2 < le && !isSpace(*t2)) ++t2;
                        const char *r = t2;
                        while (r < le && isSpace(*r)) ++r;
                        int64_t pos = 0;
                        if (r != le || !strictInt64(p, t1, pos) || !strictDouble(q, t2, v)) { C.bad = true; return; }
                        C.kind.push_back(2);
                        C.value.push_back(v);
                        C.pos.push_back(pos);
                    }
                }
                if (!nl) break;
                p = nl + 1;
            }
        });
    }
    for (auto &x : th) x.join();
    for (const FastWigChunk &C : chunks) if (C.bad) return false;
    // dry run over the headers and the line forms: every section must consist of lines of the form its header announces
    {
        WigScanner probe = sc;
        bool have = probe.haveHeader;
        try {
            for (const FastWigChunk &C : chunks) {
                size_t h = 0;
                for (size_t i = 0; i < C.kind.size(); ++i) {
                    if (C.kind[i] == 0) {
                        if (!probe.scanHeader(C.header[h].first, C.header[h].second)) return false;
                        ++h;
                        have = true;
                    } else if (!have || C.kind[i] != (probe.fixedStep ? 1 : 2)) {
                        return false;
                    }
                }
            }
        } catch (std::exception &) {
            return false; // a malformed header: the serial scanner reports it with the right line number
        }
    }
    // the real pass: visitors in input order
    size_t lineNumber = 0;
    try {
        for (const FastWigChunk &C : chunks) {
            size_t h = 0, k2 = 0;
            for (size_t i = 0; i < C.kind.size(); ++i) {
                ++lineNumber;
                if (C.kind[i] == 0) {
                    sc.scanHeader(C.header[h].first, C.header[h].second);
                    ++h;
                    sc.haveHeader = true;
                    visit(0);
                    continue;
                }
                if (sc.fixedStep) { // WiggleScanner::scanLine
                    sc.first = sc.start + sc.offset * sc.step;
                    ++sc.offset;
                } else {
                    sc.first = C.pos[k2++];
                    --sc.start;
                }
                sc.value = C.value[i];
                sc.last = sc.first;
                if (sc.span > 1) sc.last += sc.span - 1;
                visit(1);
            }
        }
    } catch (std::exception &ex) {
        throw std::runtime_error(std::string(ex.what()) + " in input wiggle line " + std::to_string(lineNumber));
    }
    visit(2);
    linesOut = lineNumber;
    return true;
}

std::string slurp(std::istream &in) {
    std::string s;
    char buf[1 << 16];
    while (in.read(buf, sizeof buf) || in.gcount() > 0) s.append(buf, (size_t)in.gcount());
    return s;
}

double seconds(std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }

} // namespace

unsigned GpuWiggleLiftover::defaultTextThreads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hw ? hw : 1u, 32u));
}

void GpuWiggleLiftover::preloadOutput(int tgtGenome, std::istream *inputFile) {
    if (_ctx == nullptr || inputFile == nullptr) throw std::runtime_error("GpuWiggleLiftover::preloadOutput: null argument");
    if (inputFile->bad()) throw std::runtime_error("Error reading wiggle input stream");
    const halgpu_seq *seqs = nullptr;
    size_t ns = 0;
    if (halgpu_sequence_table(_ctx, tgtGenome, &seqs, &ns) != 0) throw std::runtime_error("genome index out of range");
    std::map<std::string, size_t> byName;
    for (size_t i = 0; i < ns; ++i) byName[seqs[i].name] = i;
    const std::string genomeName = halgpu_genome_name(_ctx, tgtGenome);
    const int64_t genomeLen = halgpu_genome_length(_ctx, tgtGenome);
    const std::string text = slurp(*inputFile);
    // WiggleLoader::visitLine (halWiggleLoader.cpp:37-48): set() every base of the line; a later line overwrites
    std::vector<std::pair<int64_t, size_t>> order; // (position, ordinal)
    std::vector<double> vals;
    const halgpu_seq *seq = nullptr;
    WigScanner sc;
    scanWiggle(text.data(), text.size(), sc, [&](int kind) {
        if (kind == 0) {
            auto it = byName.find(sc.sequenceName);
            if (it == byName.end()) throw std::runtime_error("Sequence " + sc.sequenceName + " not found in genome " + genomeName);
            seq = &seqs[it->second];
        } else if (kind == 1) {
            const int64_t a = sc.first + seq->start, b = sc.last + seq->start;
            if (a < 0 || b >= genomeLen) throw std::runtime_error("wiggle position outside genome " + genomeName);
            for (int64_t p = a; p <= b; ++p) {
                order.emplace_back(p, vals.size());
                vals.push_back(sc.value);
            }
        }
    });
    std::sort(order.begin(), order.end());
    _prePos.clear();
    _preVal.clear();
    for (size_t i = 0; i < order.size(); ++i) {
        if (i + 1 < order.size() && order[i + 1].first == order[i].first) continue; // keep the last write to a position
        _prePos.push_back(order[i].first);
        _preVal.push_back(vals[order[i].second]);
    }
}

void GpuWiggleLiftover::convert(int srcGenome, std::istream *inputFile, int tgtGenome, std::ostream *outputFile, bool traverseDupes,
                                bool /*unique: stored but never read by the reference either*/) {
    if (_ctx == nullptr || inputFile == nullptr || outputFile == nullptr) throw std::runtime_error("GpuWiggleLiftover::convert: null argument");
    const halgpu_seq *sseq = nullptr, *tseq = nullptr;
    size_t ns = 0, nt = 0;
    if (halgpu_sequence_table(_ctx, srcGenome, &sseq, &ns) != 0 || halgpu_sequence_table(_ctx, tgtGenome, &tseq, &nt) != 0) {
        throw std::runtime_error("genome index out of range");
    }
    std::map<std::string, size_t> byName;
    for (size_t i = 0; i < ns; ++i) byName[sseq[i].name] = i;
    const std::string srcName = halgpu_genome_name(_ctx, srcGenome);
    const int64_t srcLen = halgpu_genome_length(_ctx, srcGenome);
    linesIn = runs = basesIn = basesOut = 0;
    parseSeconds = gpuSeconds = writeSeconds = 0;
    kernelMs = 0;

    // boundaries of the source segments the reference iterates: the top array if there is one, else the bottom array
    // (halWiggleLiftover.cpp:40-46)
    const int64_t numTop = halgpu_genome_num_top(_ctx, srcGenome);
    size_t stride = 40;
    const uint8_t *segs = static_cast<const uint8_t *>(halgpu_genome_top_segments(_ctx, srcGenome));
    int64_t numSegs = numTop;
    if (numTop <= 0) {
        segs = static_cast<const uint8_t *>(halgpu_genome_bottom_segments(_ctx, srcGenome, &stride));
        numSegs = halgpu_genome_num_bottom(_ctx, srcGenome);
    }
    if (segs == nullptr || numSegs <= 0) throw std::runtime_error("source genome " + srcName + " has no segments");
    auto segStart = [&](int64_t i) {
        int64_t v;
        std::memcpy(&v, segs + stride * (size_t)i, 8);
        return v;
    };
    // index of the segment containing pos, galloping from a nearby index first
    auto segOf = [&](int64_t pos, int64_t hint) {
        int64_t lo = 0, hi = numSegs; // start(lo) <= pos < start(hi)
        if (hint >= 0 && hint < numSegs && segStart(hint) <= pos) {
            lo = hint;
            int64_t stepw = 1;
            while (lo + stepw < numSegs && segStart(lo + stepw) <= pos) { lo += stepw; stepw <<= 1; }
            hi = std::min(numSegs, lo + stepw);
        }
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (segStart(mid) <= pos) lo = mid; else hi = mid;
        }
        return lo;
    };

    auto t0 = std::chrono::steady_clock::now();
    if (inputFile->bad()) throw std::runtime_error("Error reading wiggle input stream");
    const std::string text = slurp(*inputFile);
    std::vector<int64_t> runFirst, runLast, runValOff;
    std::vector<double> vals;
    bool runOpen = false;
    // replay of WiggleLiftover::visitLine / mapSegment (halWiggleLiftover.cpp:72-131) as far as it decides errors:
    // curSeg = array index of the reference's _segment, batchLines / batchLast = its _cvals
    int64_t curSeg = 0, batchLast = -1;
    size_t batchLines = 0;
    auto flushBatch = [&]() {
        if (batchLines == 0) return;
        // mapSegment leaves _segment past the segment holding the batch's last base, or on it when the batch stopped short
        // of its end (SegmentIterator::toRight, api/impl/halSegmentIterator.cpp:208-238)
        const int64_t s = segOf(batchLast, curSeg);
        curSeg = (batchLast == segStart(s + 1) - 1) ? s + 1 : s;
        batchLines = 0;
    };
    const halgpu_seq *srcSeq = nullptr;
    WigScanner sc;
    auto visitor = [&](int kind) {
        if (kind == 0) { // WiggleLiftover::visitHeader
            flushBatch();
            auto it = byName.find(sc.sequenceName);
            if (it == byName.end()) throw std::runtime_error("Sequence " + sc.sequenceName + " not found in genome " + srcName);
            srcSeq = &sseq[it->second];
            return;
        }
        if (kind == 2) { // visitEOF
            flushBatch();
            return;
        }
        const int64_t absFirst = sc.first + srcSeq->start, absLast = sc.last + srcSeq->start;
        if (absFirst < 0 || absLast >= srcLen) {
            throw std::runtime_error("wiggle position " + std::to_string(sc.last) + " of " + sc.sequenceName + " lies outside genome " + srcName);
        }
        if (curSeg >= numSegs) curSeg = 0;
        if (absFirst < segStart(curSeg) || absLast > segStart(curSeg + 1) - 1) flushBatch();
        if (batchLines > 0 && batchLast >= absFirst) throw std::runtime_error("Coordinate out of order");
        ++batchLines;
        batchLast = absLast;
        // pack into runs for the GPU
        const int64_t span = absLast - absFirst + 1;
        basesIn += (size_t)span;
        if (span >= singleValueSpan) {
            runFirst.push_back(absFirst);
            runLast.push_back(absLast);
            runValOff.push_back(~(int64_t)vals.size());
            vals.push_back(sc.value);
            runOpen = false;
        } else {
            if (!(runOpen && absFirst == runLast.back() + 1 && (size_t)(absLast - runFirst.back() + 1) <= maxRunBases)) {
                runFirst.push_back(absFirst);
                runLast.push_back(absFirst - 1);
                runValOff.push_back((int64_t)vals.size());
                runOpen = true;
            }
            for (int64_t k = 0; k < span; ++k) vals.push_back(sc.value);
            runLast.back() = absLast;
        }
    };
    unsigned threads = textThreads;
    if (const char *tt = std::getenv("HALGPU_TEXT_THREADS")) threads = (unsigned)std::max(0L, std::atol(tt));
    size_t fastLinesSeen = 0;
    if (threads > 0 && fastScanWiggle(text.data(), text.size(), threads, sc, fastLinesSeen, visitor)) {
        linesIn = fastLinesSeen;
        fastParsed = true;
    } else {
        linesIn = scanWiggle(text.data(), text.size(), sc, visitor);
    }
    runs = runFirst.size();
    parseSeconds = seconds(t0);

    t0 = std::chrono::steady_clock::now();
    halgpu_wig_result *res = nullptr;
    char *err = nullptr;
    const int rc = halgpu_wiggle_liftover(_ctx, srcGenome, tgtGenome, traverseDupes ? 0u : (uint32_t)HALGPU_NO_DUPES, runFirst.size(),
                                          runFirst.data(), runLast.data(), runValOff.data(), vals.data(), vals.size(), _prePos.size(),
                                          _prePos.data(), _preVal.data(), &res, &err);
    gpuSeconds = seconds(t0);
    if (rc != 0) {
        std::string m = err ? err : "halgpu_wiggle_liftover failed";
        halgpu_free_string(err);
        throw std::runtime_error(m);
    }
    kernelMs = res->kernel_ms;
    basesOut = res->n;

    // WiggleLiftover::write (halWiggleLiftover.cpp:160-198).  Whether position i opens a new "fixedStep" block depends on the
    // previous position and on the writer's current sequence, which is sticky: MMapSequence::getEndPosition() is start +
    // length, one PAST the last base (api/mmap_impl/mmapSequence.h:50-52), so a run that continues contiguously into the next
    // sequence gets no header there and a header printed at exactly that position still names the previous sequence (kept:
    // HAL-MMAP files are what this build reads).  The state after position j is simply "the sequence containing it" unless j
    // is the first base of its sequence, so every formatting thread can recover its starting state by looking back.
    t0 = std::chrono::steady_clock::now();
    auto seqOfPos = [&](int64_t pos) { // Genome::getSequenceBySite
        size_t lo = 0, hi = nt;
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) >> 1;
            if (tseq[mid].start <= pos) lo = mid; else hi = mid;
        }
        return (int64_t)lo;
    };
    auto step = [&](int64_t &seqIdx, int64_t &prevPos, int64_t pos, bool &needHeader) { // one iteration of the reference's loop
        if (seqIdx < 0 || pos < tseq[seqIdx].start || pos > tseq[seqIdx].start + tseq[seqIdx].length) {
            seqIdx = seqOfPos(pos);
            needHeader = true;
        } else if (pos != prevPos + 1) {
            needHeader = true;
        }
        prevPos = pos;
    };
    size_t fgrain = (size_t)1 << 18;
    if (const char *gs = std::getenv("HALGPU_WIG_GRAIN")) fgrain = (size_t)std::max(1L, std::atol(gs)); // test hook
    unsigned fthreads = std::max(1u, std::min<unsigned>(threads, (unsigned)(res->n / fgrain) + 1));
    std::vector<std::string> outs(fthreads);
    auto formatRange = [&](unsigned t) {
        const size_t a = res->n * t / fthreads, b = res->n * (t + 1) / fthreads;
        std::string &out = outs[t];
        out.reserve((b - a) * 9 + 256);
        int64_t seqIdx = -1, prevPos = -1;
        if (a > 0) { // recover the writer's state after element a - 1
            size_t j = a - 1;
            while (j > 0 && res->pos[j] == tseq[seqOfPos(res->pos[j])].start) --j;
            bool dummy = false;
            if (!(j == 0 && res->pos[0] == tseq[seqOfPos(res->pos[0])].start)) {
                seqIdx = seqOfPos(res->pos[j]);
                prevPos = res->pos[j];
                ++j;
            }
            for (; j < a; ++j) step(seqIdx, prevPos, res->pos[j], dummy);
        }
        for (size_t i = a; i < b; ++i) {
            const int64_t pos = res->pos[i];
            bool needHeader = false;
            step(seqIdx, prevPos, pos, needHeader);
            if (needHeader) {
                out += "fixedStep\tchrom=";
                out += tseq[seqIdx].name;
                out += "\tstart=";
                out += std::to_string(1 + pos - tseq[seqIdx].start);
                out += "\tstep=1\n";
            }
            char buf[40]; // operator<<(double): %g with precision 6
            auto r = std::to_chars(buf, buf + sizeof buf, res->val[i], std::chars_format::general, 6);
            out.append(buf, r.ptr);
            out += '\n';
        }
    };
    if (fthreads == 1) {
        formatRange(0);
    } else {
        std::vector<std::thread> fth;
        for (unsigned t = 0; t < fthreads; ++t) fth.emplace_back(formatRange, t);
        for (auto &x : fth) x.join();
    }
    for (const std::string &o : outs) outputFile->write(o.data(), (std::streamsize)o.size());
    writeSeconds = seconds(t0);
    halgpu_free_wig_result(res);
    _prePos.clear();
    _preVal.clear();
}

} // namespace halgpu
