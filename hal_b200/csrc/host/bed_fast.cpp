#include "bed_fast.hpp"
#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <thread>

namespace halgpu {

void parallelFor(unsigned nThreads, const std::function<void(unsigned)> &fn) {
    if (nThreads <= 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> errs(nThreads);
    th.reserve(nThreads);
    for (unsigned t = 0; t < nThreads; ++t) {
        th.emplace_back([&, t] {
            try {
                fn(t);
            } catch (...) {
                errs[t] = std::current_exception();
            }
        });
    }
    for (auto &x : th) x.join();
    for (auto &e : errs) {
        if (e) std::rethrow_exception(e);
    }
}

namespace {

inline bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); } // std::isspace, "C" locale

// Strict integer: -?[0-9]+ over the whole field.  Anything looser (blanks, '+', trailing text -- all of which
// hal::strToInt tolerates, api/impl/halCommon.cpp:45-53) is left to the serial path.
inline bool strictInt(const char *b, const char *e, int64_t &v) {
    if (b == e) return false;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e && *b != '+';
}

struct Fields {
    const char *tab[13]; // tab[i] = position of the i-th TAB (or end of line): field i is [i ? tab[i-1]+1 : line, tab[i])
    int n;               // number of fields found, counting at most 13
};

// Splits [p, e) at TABs, stopping after `want` fields (the rest of the line is the pass-through tail).
inline void splitTabs(const char *p, const char *e, int want, Fields &f) {
    f.n = 0;
    const char *q = p;
    while (f.n < want) {
        const char *t = static_cast<const char *>(std::memchr(q, '\t', (size_t)(e - q)));
        f.tab[f.n++] = t ? t : e;
        if (!t) break;
        q = t + 1;
    }
}

struct Parsed { // the fields of one accepted line the formatter needs again
    int64_t start, end, score, thickStart, thickEnd, rgb[3];
    char strand;
};

// Validates one line as BED `bedType` in the strict dialect.  nf = total number of columns.
inline bool parseStrict(const char *p, const char *e, int bedType, const Fields &f, Parsed &o) {
    auto fb = [&](int i) { return i == 0 ? p : f.tab[i - 1] + 1; };
    auto fe = [&](int i) { return f.tab[i]; };
    if (!strictInt(fb(1), fe(1), o.start) || !strictInt(fb(2), fe(2), o.end) || o.start >= o.end) return false;
    if (bedType > 4 && !strictInt(fb(4), fe(4), o.score)) return false;
    if (bedType > 5) {
        if (fe(5) - fb(5) != 1) return false;
        o.strand = *fb(5);
        if (o.strand != '+' && o.strand != '-' && o.strand != '.') return false;
    }
    if (bedType > 6 && !strictInt(fb(6), fe(6), o.thickStart)) return false;
    if (bedType > 7 && !strictInt(fb(7), fe(7), o.thickEnd)) return false;
    if (bedType > 8) { // BedLine::read itemRGB: 1 to 3 comma separated values, missing ones copy the first (halBedLine.cpp:73-86)
        const char *q = fb(8), *qe = fe(8);
        int k = 0;
        while (true) {
            const char *c = static_cast<const char *>(std::memchr(q, ',', (size_t)(qe - q)));
            if (k == 3 || !strictInt(q, c ? c : qe, o.rgb[k])) return false;
            ++k;
            if (!c) break;
            q = c + 1;
        }
        if (k == 1) o.rgb[1] = o.rgb[2] = o.rgb[0];
        if (k == 2) o.rgb[2] = o.rgb[0];
    }
    (void)e;
    return true;
}

inline void putInt(std::string &out, int64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    out.append(buf, r.ptr);
}

struct ThreadParse {
    std::vector<FastLine> lines;
    std::vector<int64_t> gs, ge;
    std::vector<uint8_t> st;
    std::vector<FastEvent> events;
    size_t linesSeen = 0;
    uint64_t lastOff = 0;
    uint32_t lastLen = 0;
    int bedType = 0; // 0: none yet, -1: mixed
    bool bail = false;
};

} // namespace

FastBedBlock::FastBedBlock(const halgpu_seq *srcSeqs, size_t nSrc, const halgpu_seq *tgtSeqs, size_t nTgt)
    : _src(srcSeqs), _tgt(tgtSeqs), _nTgt(nTgt) {
    _seqByName.reserve(nSrc * 2);
    // like the std::map of the serial path, a duplicated name resolves to its LAST index
    for (size_t i = 0; i < nSrc; ++i) _seqByName[std::string_view(srcSeqs[i].name)] = (int32_t)i;
}

FastBedBlock::~FastBedBlock() {
    halgpu_host_free(_gs);
    halgpu_host_free(_ge);
    halgpu_host_free(_st);
}

void FastBedBlock::reservePinned(size_t n) {
    if (n <= _cap) return;
    halgpu_host_free(_gs);
    halgpu_host_free(_ge);
    halgpu_host_free(_st);
    _cap = n + n / 8 + 1024;
    _gs = static_cast<int64_t *>(halgpu_host_alloc(_cap * 8));
    _ge = static_cast<int64_t *>(halgpu_host_alloc(_cap * 8));
    _st = static_cast<uint8_t *>(halgpu_host_alloc(_cap));
    if (!_gs || !_ge || !_st) {
        _cap = 0;
        throw std::runtime_error("cannot allocate page-locked staging buffers");
    }
}

bool FastBedBlock::parseBlock(const char *block, size_t n, int forcedBedType, const BedLine &sticky, unsigned nThreads) {
    _lines.clear();
    _events.clear();
    _n = _linesSeen = 0;
    _bedType = 0;
    _stickyStrand = sticky.strand;
    _stickyThickEnd = sticky.thickEnd;
    nThreads = std::max(1u, std::min<unsigned>(nThreads, (unsigned)(n / (1u << 16)) + 1));
    // slice boundaries: just after a newline
    std::vector<size_t> cut(nThreads + 1, n);
    cut[0] = 0;
    for (unsigned t = 1; t < nThreads; ++t) {
        size_t c = std::max(cut[t - 1], n * t / nThreads);
        const char *nl = c < n ? static_cast<const char *>(std::memchr(block + c, '\n', n - c)) : nullptr;
        cut[t] = nl ? (size_t)(nl - block) + 1 : n;
    }
    std::vector<ThreadParse> tp(nThreads);
    parallelFor(nThreads, [&](unsigned t) {
        ThreadParse &T = tp[t];
        const char *p = block + cut[t], *e = block + cut[t + 1];
        const size_t guess = (size_t)(e - p) / 20 + 16;
        T.lines.reserve(guess); T.gs.reserve(guess); T.ge.reserve(guess); T.st.reserve(guess);
        std::string_view lastName;
        int32_t lastSeq = -1;
        Fields f;
        Parsed v;
        while (true) {
            while (p < e && isSpace(*p)) ++p; // BedScanner::skipWhiteSpaces (halBedScanner.cpp:63-69): blank lines, leading blanks
            if (p >= e) break;
            const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(e - p)));
            const char *le = nl ? nl : e;
            ++T.linesSeen;
            T.lastOff = (uint64_t)(p - block);
            T.lastLen = (uint32_t)(le - p);
            if ((size_t)(le - p) > 0xffffffffu || le[-1] == '\t') { T.bail = true; return; } // chopString drops a trailing empty field
            splitTabs(p, le, 13, f);
            const int nf = f.n;
            const int bt = forcedBedType ? forcedBedType : std::min(nf, 12);
            if (nf < 3 || bt > 9 || bt > nf) { T.bail = true; return; }
            if (T.bedType == 0) T.bedType = bt;
            else if (T.bedType != bt) { T.bail = true; return; }
            if (!parseStrict(p, le, bt, f, v)) { T.bail = true; return; }
            const std::string_view chr(p, (size_t)(f.tab[0] - p));
            if (chr != lastName || lastSeq < 0) {
                auto it = _seqByName.find(chr);
                lastSeq = it == _seqByName.end() ? -1 : it->second;
                lastName = chr;
            }
            if (lastSeq < 0) {
                T.events.push_back(FastEvent{FastEvent::MISSING_SEQUENCE, std::string(chr), v.end, 0, 0});
            } else if (v.end > _src[lastSeq].length) {
                T.events.push_back(FastEvent{FastEvent::ENDPOINT_BEYOND_SEQUENCE, std::string(chr), v.end, _src[lastSeq].length, 0});
            } else {
                const int64_t base = _src[lastSeq].start;
                T.lines.push_back(FastLine{T.lastOff, T.lastLen, lastSeq});
                T.gs.push_back(v.start + base);        // halBlockLiftover.cpp:48-49
                T.ge.push_back(v.end - 1 + base);
                T.st.push_back((uint8_t)(bt > 5 ? v.strand : _stickyStrand));
            }
            if (!nl) break;
            p = nl + 1;
        }
    });
    size_t total = 0;
    for (const ThreadParse &T : tp) {
        if (T.bail) return false;
        if (T.bedType != 0) {
            if (_bedType == 0) _bedType = T.bedType;
            else if (_bedType != T.bedType) return false;
        }
        total += T.lines.size();
    }
    reservePinned(total);
    _lines.resize(total);
    std::vector<size_t> at(nThreads + 1, 0);
    for (unsigned t = 0; t < nThreads; ++t) at[t + 1] = at[t] + tp[t].lines.size();
    parallelFor(nThreads, [&](unsigned t) {
        const ThreadParse &T = tp[t];
        const size_t c = T.lines.size();
        if (c == 0) return;
        std::memcpy(_lines.data() + at[t], T.lines.data(), c * sizeof(FastLine));
        std::memcpy(_gs + at[t], T.gs.data(), c * 8);
        std::memcpy(_ge + at[t], T.ge.data(), c * 8);
        std::memcpy(_st + at[t], T.st.data(), c);
    });
    _n = total;
    for (unsigned t = 0; t < nThreads; ++t) {
        ThreadParse &T = tp[t];
        _linesSeen += T.linesSeen;
        if (T.linesSeen > 0) { _lastOff = T.lastOff; _lastLen = T.lastLen; }
        for (FastEvent &ev : T.events) _events.push_back(std::move(ev));
    }
    return true;
}

TextBuf::~TextBuf() { std::free(_p); }

void TextBuf::grow(size_t need) {
    const size_t cap = std::max(_cap + _cap / 2, _n + need + (1u << 16));
    char *q = static_cast<char *>(std::realloc(_p, cap));
    if (!q) throw std::bad_alloc();
    _p = q;
    _cap = cap;
}

namespace {
inline char *wInt(char *w, int64_t v) { return std::to_chars(w, w + 24, v).ptr; }
inline char *wBytes(char *w, const char *b, size_t n) {
    std::memcpy(w, b, n);
    return w + n;
}
} // namespace

size_t FastBedBlock::formatBlock(const char *block, const halgpu_lift_result *res, unsigned nThreads, std::vector<TextBuf> &text) const {
    const size_t n = _n;
    nThreads = std::max(1u, std::min<unsigned>(nThreads, (unsigned)(n / 4096) + 1));
    if (text.size() < nThreads) text.resize(nThreads);
    for (TextBuf &b : text) b.clear();
    std::vector<size_t> outLines(nThreads, 0);
    std::vector<size_t> tgtNameLen(_nTgt);
    for (size_t i = 0; i < _nTgt; ++i) tgtNameLen[i] = std::strlen(_tgt[i].name);
    const int bt = _bedType;
    parallelFor(nThreads, [&](unsigned t) {
        const size_t la = n * t / nThreads, lb = n * (t + 1) / nThreads;
        TextBuf &out = text[t];
        if (lb > la) out.room((size_t)((res->offsets[lb] - res->offsets[la]) * 36 + 64));
        Fields f;
        Parsed v;
        std::vector<uint64_t> order;
        std::string mid, tail; // per input line: "\tname\tscore" and everything after the thick columns
        for (size_t i = la; i < lb; ++i) {
            const uint64_t r0 = res->offsets[i], r1 = res->offsets[i + 1];
            if (r0 == r1) continue;
            const FastLine &L = _lines[i];
            const char *p = block + L.off, *le = p + L.len;
            splitTabs(p, le, bt + 1, f);
            v.thickStart = 0;
            v.thickEnd = _stickyThickEnd;
            parseStrict(p, le, bt, f, v); // accepted before: cannot fail
            mid.clear();
            tail.clear();
            if (bt > 3) { mid += '\t'; mid.append(f.tab[2] + 1, f.tab[3]); }
            if (bt > 4) { mid += '\t'; putInt(mid, v.score); }
            const bool thick = v.thickStart != 0 || v.thickEnd != 0; // cleanResults (halLiftover.cpp:318-325)
            if (bt > 8) { tail += '\t'; putInt(tail, v.rgb[0]); tail += ','; putInt(tail, v.rgb[1]); tail += ','; putInt(tail, v.rgb[2]); }
            if (f.tab[bt - 1] != le) tail.append(f.tab[bt - 1], le); // pass-through columns, verbatim (with their leading TAB)
            // Liftover::visitLine sorts the mapped lines (stable) by source start (halLiftover.cpp:90); the library
            // already returns them in that order -- verified here, re-sorted only if it ever is not
            bool sorted = true;
            for (uint64_t r = r0 + 1; r < r1; ++r) sorted &= res->recs[r - 1].src_start <= res->recs[r].src_start;
            order.clear();
            if (!sorted) {
                for (uint64_t r = r0; r < r1; ++r) order.push_back(r);
                std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return res->recs[a].src_start < res->recs[b].src_start; });
            }
            const size_t fixed = mid.size() + tail.size() + 6 * 24;
            for (uint64_t k = r0; k < r1; ++k) {
                const halgpu_lift_rec &rec = res->recs[sorted ? k : order[k - r0]];
                const size_t nl = tgtNameLen[rec.tgt_seq];
                char *w = out.room(fixed + nl);
                w = wBytes(w, _tgt[rec.tgt_seq].name, nl); *w++ = '\t'; w = wInt(w, rec.start); *w++ = '\t'; w = wInt(w, rec.end);
                w = wBytes(w, mid.data(), mid.size());
                if (bt > 5) { *w++ = '\t'; *w++ = (char)rec.strand; }
                if (bt > 6) { *w++ = '\t'; w = wInt(w, thick ? rec.start : v.thickStart); }
                if (bt > 7) { *w++ = '\t'; w = wInt(w, thick ? rec.end : v.thickEnd); }
                w = wBytes(w, tail.data(), tail.size());
                *w++ = '\n';
                out.advanceTo(w);
            }
            outLines[t] += (size_t)(r1 - r0);
        }
    });
    size_t total = 0;
    for (size_t c : outLines) total += c;
    return total;
}

} // namespace halgpu
