#include "gpu_liftover.hpp"
#include "bed_fast.hpp"
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <chrono>
#include <cstring>
#include <iostream>
#include <istream>
#include <list>
#include <map>
#include <ostream>
#include <stdexcept>
#include <vector>

namespace halgpu {

namespace {

struct PendingLine {
    BedLine bed;          // the parsed input line (a full copy: sticky fields already resolved)
    size_t firstInterval; // first element of the batch arrays that belongs to this line
    size_t numIntervals;  // 1 for BED<=9, one per non-empty block for BED12
};

// Liftover::compatible (liftover/impl/halLiftover.cpp:169-195)
bool compatible(const BedLine &tgtBed, const BedLine &newBlock, char inputStrand) {
    if (tgtBed.strand != newBlock.strand) return false;
    if (tgtBed.srcStart == newBlock.srcStart) return false;
    const BedBlock &tb = tgtBed.blocks.back();
    int64_t delta;
    if (tgtBed.strand != inputStrand) delta = tb.start - newBlock.end;
    else delta = newBlock.start - (tb.start + tb.length);
    if (delta < 0) return false;
    return tgtBed.chrName == newBlock.chrName;
}

// Hands the input stream out in blocks of whole lines: about `target` bytes, extended to the end of the line the
// cut falls into.  The final block may lack a trailing newline (as the final line of a BED file may).
struct BlockBuf { // text of one block, owned by a pipeline slot
    std::unique_ptr<char[]> buf;
    size_t cap = 0;
};
class BlockReader {
  public:
    BlockReader(std::istream &in, size_t target) : _in(in), _target(std::max<size_t>(target, 1)) {}
    bool next(BlockBuf &b, const char *&p, size_t &n) {
        if (!_in.good()) return false;
        if (b.cap < _target) {
            b.buf.reset(new char[_target]); // uninitialised on purpose: pages are touched only as far as the input reaches
            b.cap = _target;
        }
        _in.read(b.buf.get(), (std::streamsize)_target);
        size_t got = (size_t)_in.gcount();
        if (got == _target && _in.good() && b.buf[got - 1] != '\n') {
            std::string tail;
            std::getline(_in, tail); // consumes the newline; it only separated lines
            if (got + tail.size() > b.cap) {
                std::unique_ptr<char[]> bigger(new char[got + tail.size()]);
                std::memcpy(bigger.get(), b.buf.get(), got);
                b.buf = std::move(bigger);
                b.cap = got + tail.size();
            }
            std::memcpy(b.buf.get() + got, tail.data(), tail.size());
            got += tail.size();
        }
        p = b.buf.get();
        n = got;
        return got > 0;
    }

  private:
    std::istream &_in;
    size_t _target;
};

// The fast text path as a three-stage pipeline over blocks: the calling thread reads and tokenises block k+2 while one
// worker lifts block k+1 (halgpu_liftover: PCIe + GPU) and another formats and writes block k.  Blocks leave in input order.
struct FastSlot {
    BlockBuf text;
    const char *block = nullptr;
    size_t blockLen = 0;
    std::unique_ptr<FastBedBlock> fast;
    std::vector<TextBuf> out;
    halgpu_lift_result *res = nullptr;
    size_t firstLine = 0; // input line number of the block's first line
    int state = 0; // 0 free, 1 parsed (waits for the lift), 2 lifted (waits for format + write)
};
class FastPipeline {
  public:
    FastPipeline(size_t nSlots, std::function<halgpu_lift_result *(FastSlot &)> lift, std::function<void(FastSlot &)> emit)
        : _slots(nSlots), _lift(std::move(lift)), _emit(std::move(emit)) {
        _lifter = std::thread([this] { run(1, 2, _lift2); });
        _writer = std::thread([this] { run(2, 0, _emit2); });
    }
    ~FastPipeline() {
        {
            std::lock_guard<std::mutex> g(_m);
            _stop = true;
        }
        _cv.notify_all();
        _lifter.join();
        _writer.join();
        for (FastSlot &s : _slots) if (s.res) halgpu_free_result(s.res);
    }
    FastSlot &acquire() { // the next slot in ring order, once it is free again
        std::unique_lock<std::mutex> g(_m);
        FastSlot &s = _slots[_head % _slots.size()];
        _cv.wait(g, [&] { return s.state == 0 || !_error.empty(); });
        rethrow();
        return s;
    }
    void submit() { // the slot handed out by the last acquire() is parsed
        {
            std::lock_guard<std::mutex> g(_m);
            _slots[_head % _slots.size()].state = 1;
            ++_head;
        }
        _cv.notify_all();
    }
    void drain() { // everything submitted so far has been written
        std::unique_lock<std::mutex> g(_m);
        _cv.wait(g, [&] {
            if (!_error.empty()) return true;
            for (const FastSlot &s : _slots) if (s.state != 0) return false;
            return true;
        });
        rethrow();
    }

  private:
    void rethrow() {
        if (!_error.empty()) throw std::runtime_error(_error);
    }
    void run(int from, int to, const std::function<void(FastSlot &)> &work) {
        size_t next = 0;
        while (true) {
            FastSlot *s;
            {
                std::unique_lock<std::mutex> g(_m);
                s = &_slots[next % _slots.size()];
                _cv.wait(g, [&] { return _stop || s->state == from; });
                if (s->state != from) return; // stop requested and nothing left for this stage
            }
            try {
                work(*s);
            } catch (const std::exception &e) {
                std::lock_guard<std::mutex> g(_m);
                if (_error.empty()) _error = e.what();
                _stop = true;
                _cv.notify_all();
                return;
            }
            {
                std::lock_guard<std::mutex> g(_m);
                s->state = to;
            }
            _cv.notify_all();
            ++next;
        }
    }
    std::vector<FastSlot> _slots;
    std::function<halgpu_lift_result *(FastSlot &)> _lift;
    std::function<void(FastSlot &)> _emit;
    std::function<void(FastSlot &)> _lift2 = [this](FastSlot &s) { s.res = _lift(s); };
    std::function<void(FastSlot &)> _emit2 = [this](FastSlot &s) { _emit(s); };
    std::mutex _m;
    std::condition_variable _cv;
    std::thread _lifter, _writer;
    size_t _head = 0;
    bool _stop = false;
    std::string _error;
};

} // namespace

unsigned GpuBlockLiftover::defaultTextThreads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hw ? hw : 1u, 32u));
}

void GpuBlockLiftover::convert(int srcGenome, std::istream *in, int tgtGenome, std::ostream *out, int bedType,
                               bool traverseDupes, bool outPSL, bool outPSLWithName, int coalescenceLimit) {
    if (_ctx == nullptr || in == nullptr || out == nullptr) throw std::runtime_error("GpuBlockLiftover::convert: null argument");
    if (outPSLWithName) outPSL = true;
    if (columnLiftover && outPSL) throw std::runtime_error("PSL output needs BlockLiftover (ColumnLiftover has no source coordinates)");
    const halgpu_seq *sseq = nullptr, *tseq = nullptr;
    size_t ns = 0, nt = 0;
    if (halgpu_sequence_table(_ctx, srcGenome, &sseq, &ns) != 0 || halgpu_sequence_table(_ctx, tgtGenome, &tseq, &nt) != 0) {
        throw std::runtime_error("genome index out of range");
    }
    std::map<std::string, size_t> seqByName;
    for (size_t i = 0; i < ns; ++i) seqByName[sseq[i].name] = i;
    const std::string srcName = halgpu_genome_name(_ctx, srcGenome);
    _missed.clear();
    linesIn = intervalsLifted = linesOut = fastLines = 0;
    gpuSeconds = textSeconds = writeSeconds = parseSeconds = readSeconds = 0;

    if (in->bad()) throw std::runtime_error("Error reading bed input stream");
    BedLine cur; // persists across lines like BedScanner::_bedLine
    std::string lineBuf, outBuf;
    size_t lineNumber = 0;
    unsigned threads = textThreads;
    if (const char *tt = std::getenv("HALGPU_TEXT_THREADS")) threads = (unsigned)std::max(0L, std::atol(tt));
    size_t blockTarget = blockBytes;
    if (const char *bb = std::getenv("HALGPU_BLOCK_BYTES")) blockTarget = (size_t)std::max(1L, std::atol(bb)); // test hook
    BlockReader reader(*in, blockTarget);
    const char *block = nullptr;
    size_t blockLen = 0;
    const uint32_t liftFlags = (traverseDupes ? 0u : (uint32_t)HALGPU_NO_DUPES) | (outPSL ? (uint32_t)HALGPU_PSL : 0u) |
                               (columnLiftover ? (uint32_t)HALGPU_COLUMN_LIFTOVER : 0u);
    std::mutex statMutex; // the pipeline's workers add to the totals too
    auto lift = [&](size_t n, const int64_t *gs, const int64_t *ge, const uint8_t *st) {
        halgpu_lift_result *res = nullptr;
        char *err = nullptr;
        auto t0 = std::chrono::steady_clock::now();
        const int rc = halgpu_liftover(_ctx, srcGenome, tgtGenome, coalescenceLimit, liftFlags, n, gs, ge, st, &res, &err);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        {
            std::lock_guard<std::mutex> g(statMutex);
            gpuSeconds += dt;
            if (rc == 0) intervalsLifted += n;
        }
        if (rc != 0) {
            std::string m = err ? err : "halgpu_liftover failed";
            halgpu_free_string(err);
            throw std::runtime_error(m);
        }
        return res;
    };
    // the fast text path runs as a pipeline over blocks (FastPipeline); its two workers
    std::unique_ptr<FastPipeline> pipe;
    BlockBuf serialText; // block buffer of the serial path (and of the fast path's look-ahead read)
    auto liftSlot = [&](FastSlot &sl) -> halgpu_lift_result * {
        if (sl.fast->numIntervals() == 0) return nullptr;
        try {
            return lift(sl.fast->numIntervals(), sl.fast->starts(), sl.fast->endsIncl(), sl.fast->strands());
        } catch (std::exception &e) { // the reference raises inside liftInterval of the first line it maps (halBedScanner.cpp:53-58)
            throw std::runtime_error(std::string(e.what()) + " in input bed line " + std::to_string(sl.firstLine));
        }
    };
    auto emitSlot = [&](FastSlot &sl) {
        if (sl.res == nullptr) return;
        auto t0 = std::chrono::steady_clock::now();
        const size_t lines = sl.fast->formatBlock(sl.block, sl.res, threads, sl.out);
        const double tf = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        halgpu_free_result(sl.res);
        sl.res = nullptr;
        t0 = std::chrono::steady_clock::now();
        for (const TextBuf &tb : sl.out) out->write(tb.data(), (std::streamsize)tb.size());
        const double tw = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::lock_guard<std::mutex> g(statMutex);
        linesOut += lines; textSeconds += tf; writeSeconds += tw;
    };
    const bool fastEligible = threads > 0 && !outPSL;
    if (fastEligible) {
        pipe.reset(new FastPipeline(3, liftSlot, emitSlot));
    }
    while (true) {
        FastSlot *sl = nullptr;
        if (fastEligible) {
            sl = &pipe->acquire();
        }
        BlockBuf &textBuf = sl ? sl->text : serialText;
        {
            auto tr = std::chrono::steady_clock::now();
            const bool more = reader.next(textBuf, block, blockLen);
            readSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tr).count();
            if (!more) break;
        }
        // ---- fast path: the whole block tokenised and formatted by `threads` threads (bed_fast.hpp) ----
        if (sl != nullptr) {
            if (!sl->fast) sl->fast.reset(new FastBedBlock(sseq, ns, tseq, nt));
            FastBedBlock &fast = *sl->fast;
            auto t0 = std::chrono::steady_clock::now();
            const bool ok = fast.parseBlock(block, blockLen, bedType, cur, threads);
            const double tp = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> g(statMutex);
                textSeconds += tp;
            }
            parseSeconds += tp;
            if (ok) {
                for (const FastEvent &ev : fast.events()) { // Liftover::visitLine's messages (halLiftover.cpp:53-69), in input order
                    if (ev.kind == FastEvent::MISSING_SEQUENCE) {
                        if (_missed.insert(ev.chrName).second) {
                            std::cerr << "Unable to find sequence " << ev.chrName << " in genome " << srcName << std::endl;
                        }
                    } else {
                        std::cerr << "Skipping interval with endpoint " << ev.end << "because sequence " << ev.chrName << " has length "
                                  << ev.seqLength << std::endl;
                    }
                }
                sl->firstLine = lineNumber + 1;
                lineNumber += fast.linesSeen();
                linesIn += fast.linesSeen();
                fastLines += fast.linesSeen();
                uint64_t lo = 0;
                uint32_t ll = 0;
                if (fast.lastLine(lo, ll)) { // carry the scanner's sticky BedLine on for the next block
                    lineBuf.assign(block + lo, ll);
                    cur.parse(lineBuf, bedType);
                }
                sl->block = block; sl->blockLen = blockLen;
                pipe->submit();
                continue;
            }
            pipe->drain(); // the serial path below writes to the same stream: everything before this block goes out first
        }
        // ---- serial path over the block: BedLine::parse per line, every BED flavour, the reference's messages ----
        const char *p = block, *const blockEnd = block + blockLen;
        auto skipWs = [&]() {
            while (p < blockEnd && std::isspace((unsigned char)*p)) ++p;
        };
        skipWs();
        while (p < blockEnd) {
        std::vector<PendingLine> pending;
        std::vector<int64_t> gs, ge;
        std::vector<uint8_t> st;
        std::string deferredError; // a malformed line ends the run, but only after the lines before it are lifted and written:
                                   // the reference streams every line's result before it reads the next (halBedScanner.cpp:40-61)
        size_t batchFirstLine = 0;
        // ---- read one batch (Liftover::visitLine up to the liftInterval call) ----
        while (pending.size() < batchLines && p < blockEnd) {
            ++lineNumber;
            const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(blockEnd - p)));
            lineBuf.assign(p, nl ? nl : blockEnd);
            p = nl ? nl + 1 : blockEnd;
            try {
                cur.parse(lineBuf, bedType);
            } catch (std::exception &e) {
                deferredError = std::string(e.what()) + " in input bed line " + std::to_string(lineNumber);
                break;
            }
            skipWs();
            ++linesIn;
            if (outPSL && cur.bedType < 12) cur.expandToBed12(); // forcing to BED12 makes PSL code simpler (halLiftover.cpp:47-50)
            auto it = seqByName.find(cur.chrName);
            if (it == seqByName.end()) {
                if (_missed.insert(cur.chrName).second) {
                    std::cerr << "Unable to find sequence " << cur.chrName << " in genome " << srcName << std::endl;
                }
                continue;
            }
            const halgpu_seq &sq = sseq[it->second];
            if (cur.end > sq.length) {
                std::cerr << "Skipping interval with endpoint " << cur.end << "because sequence " << cur.chrName << " has length "
                          << sq.length << std::endl;
                continue;
            }
            if (cur.bedType > 9 && cur.blocks.empty()) {
                std::cerr << "Skipping input line with 0 blocks" << std::endl;
                continue;
            }
            PendingLine p;
            p.firstInterval = gs.size();
            if (cur.bedType <= 9) {
                gs.push_back(cur.start + sq.start);
                ge.push_back(cur.end - 1 + sq.start);
                st.push_back((uint8_t)cur.strand);
            } else { // liftBlockIntervals (halLiftover.cpp:296-310): blocks in ascending start order
                std::sort(cur.blocks.begin(), cur.blocks.end(), [](const BedBlock &a, const BedBlock &b) { return a.start < b.start; });
                for (const BedBlock &b : cur.blocks) {
                    const int64_t s0 = b.start + cur.start, e0 = s0 + b.length;
                    if (e0 > s0) {
                        gs.push_back(s0 + sq.start);
                        ge.push_back(e0 - 1 + sq.start);
                        st.push_back((uint8_t)cur.strand);
                    }
                }
            }
            p.numIntervals = gs.size() - p.firstInterval;
            p.bed = cur;
            if (pending.empty()) batchFirstLine = lineNumber;
            pending.push_back(std::move(p));
        }
        if (pending.empty()) {
            if (!deferredError.empty()) { out->flush(); throw std::runtime_error(deferredError); }
            continue;
        }

        // ---- one GPU call for the whole batch ----
        halgpu_lift_result *res = nullptr;
        try {
            res = lift(gs.size(), gs.data(), ge.data(), st.data());
        } catch (std::exception &e) { // the reference raises inside liftInterval of the first line it maps
            throw std::runtime_error(std::string(e.what()) + " in input bed line " + std::to_string(batchFirstLine));
        }

        // ---- per-line post-processing and output ----
        outBuf.clear();
        std::vector<BedLine> mapped;
        std::list<BedLine> outLines;
        for (const PendingLine &p : pending) {
            const BedLine &src = p.bed;
            mapped.clear();
            for (size_t k = 0; k < p.numIntervals; ++k) {
                const size_t iv = p.firstInterval + k;
                for (uint64_t r = res->offsets[iv]; r < res->offsets[iv + 1]; ++r) {
                    const halgpu_lift_rec &rec = res->recs[r];
                    mapped.push_back(src); // outBedLine = _bedLine (halBlockLiftover.cpp:84-86)
                    BedLine &o = mapped.back();
                    o.blocks.clear();
                    o.chrName = tseq[rec.tgt_seq].name;
                    o.start = rec.start;
                    o.end = rec.end;
                    o.strand = (char)rec.strand;
                    o.srcStart = rec.src_start;
                    o.srcStrand = (char)rec.src_strand;
                    if (outPSL) { // BlockLiftover::readPSLInfo (halBlockLiftover.cpp:115-162); base counts come from the GPU
                        const halgpu_seq &ssq = sseq[seqByName[src.chrName]];
                        o.psl.assign(1, PslInfo());
                        PslInfo &ps = o.psl[0];
                        ps.matches = res->psl[4 * r]; ps.misMatches = res->psl[4 * r + 1];
                        ps.repMatches = res->psl[4 * r + 2]; ps.nCount = res->psl[4 * r + 3];
                        ps.qSeqName = ssq.name;
                        ps.qSeqSize = (uint64_t)ssq.length;
                        ps.qStrand = src.strand == '-' ? '-' : '+'; // the source is reversed iff the BED strand is '-' (:50,64-70)
                        ps.qChromOffset = (uint64_t)ssq.start;
                        ps.qEnd = (uint64_t)(o.srcStart + (o.end - o.start));
                        ps.tSeqSize = (uint64_t)tseq[rec.tgt_seq].length;
                    }
                }
            }
            outLines.clear();
            if (src.bedType <= 9) { // writeBlocksAsIntervals
                outLines.assign(mapped.begin(), mapped.end());
            } else if (!mapped.empty()) { // assignBlocksToIntervals (halLiftover.cpp:108-167), BED output
                std::stable_sort(mapped.begin(), mapped.end(), [](const BedLine &a, const BedLine &b) { return a.srcStart < b.srcStart; });
                int64_t prevSrcBlockEnd = -1;
                for (size_t bi = 0; bi < mapped.size(); ++bi) {
                    const BedLine &blk = mapped[bi];
                    const int64_t srcBlockEnd = blk.srcStart + (blk.end - blk.start);
                    const bool dupe = blk.srcStart < prevSrcBlockEnd || (bi + 1 < mapped.size() && mapped[bi + 1].srcStart < srcBlockEnd);
                    // filter dupes in psl but let them be on single bed line
                    if (outLines.empty() || (outPSL && dupe) || !compatible(outLines.back(), blk, src.strand)) {
                        outLines.push_back(blk);
                    }
                    prevSrcBlockEnd = srcBlockEnd;
                    BedLine &t = outLines.back();
                    t.start = std::min(t.start, blk.start);
                    t.end = std::max(t.end, blk.end);
                    BedBlock b;
                    b.start = blk.start; // absolute for now
                    b.length = blk.end - blk.start;
                    t.blocks.push_back(b);
                    if (outPSL) {
                        PslInfo &tp = t.psl[0];
                        tp.qBlockStarts.push_back(blk.srcStart);
                        if (t.blocks.size() > 1) { // the first block's counts came with the copy
                            tp.matches += blk.psl[0].matches; tp.misMatches += blk.psl[0].misMatches;
                            tp.repMatches += blk.psl[0].repMatches; tp.nCount += blk.psl[0].nCount;
                        }
                    }
                }
                for (BedLine &t : outLines) {
                    for (BedBlock &b : t.blocks) b.start -= t.start;
                    if (t.blocks.size() > 1) { // flipBlocks (halLiftover.cpp:197-234)
                        const int64_t delta = t.blocks[1].start - (t.blocks[0].start + t.blocks[0].length);
                        const bool mustFlip = !outPSL ? delta < 0 : ((t.strand == '-' && delta >= 0) || (t.strand != '-' && delta < 0));
                        if (mustFlip) {
                            std::reverse(t.blocks.begin(), t.blocks.end());
                            if (outPSL) std::reverse(t.psl[0].qBlockStarts.begin(), t.psl[0].qBlockStarts.end());
                        }
                    }
                    if (outPSL) { // computePSLInserts (halLiftover.cpp:236-290)
                        PslInfo &ps = t.psl[0];
                        ps.qNumInsert = ps.qBaseInsert = ps.tNumInsert = ps.tBaseInsert = 0;
                        for (size_t i = 1; i < t.blocks.size(); ++i) {
                            const BedBlock *cur = &t.blocks[i], *prev = &t.blocks[i - 1];
                            const BedBlock *tc = cur, *tp = prev;
                            if (t.strand == '-') std::swap(tc, tp);
                            const int64_t tgap = tc->start - (tp->start + tp->length);
                            if (tgap > 0) { ++ps.tNumInsert; ps.tBaseInsert += (uint64_t)tgap; }
                            int64_t qc = ps.qBlockStarts[i], qp = ps.qBlockStarts[i - 1];
                            const BedBlock *qpb = prev;
                            if (ps.qStrand == '-') { std::swap(qc, qp); qpb = cur; }
                            const int64_t qgap = qc >= qp + qpb->length ? qc - (qp + qpb->length) : 0; // duplicated blocks can overlap
                            if (qgap > 0) { ++ps.qNumInsert; ps.qBaseInsert += (uint64_t)qgap; }
                        }
                    }
                }
            }
            // cleanResults (halLiftover.cpp:313-355)
            if (src.bedType > 6) {
                for (auto it = outLines.begin(); it != outLines.end();) {
                    if (src.thickStart != 0 || src.thickEnd != 0) {
                        it->thickStart = it->start;
                        it->thickEnd = it->end;
                    }
                    if (src.bedType > 9 && it->blocks.empty()) { it = outLines.erase(it); continue; }
                    if (src.bedType > 9 && outPSL) { // halLiftover.cpp:336-346
                        PslInfo &ps = it->psl[0];
                        it->srcStart = INT64_MAX;
                        ps.qEnd = 0;
                        for (size_t j = 0; j < ps.qBlockStarts.size(); ++j) {
                            it->srcStart = std::min(it->srcStart, ps.qBlockStarts[j]);
                            ps.qEnd = std::max<uint64_t>(ps.qEnd, (uint64_t)(ps.qBlockStarts[j] + it->blocks[j].length));
                        }
                    }
                    ++it;
                }
            }
            outLines.sort([](const BedLine &a, const BedLine &b) { return a.srcStart < b.srcStart; }); // stable
            for (const BedLine &l : outLines) {
                if (outPSL) l.appendPSL(outBuf, outPSLWithName); else l.append(outBuf);
                ++linesOut;
            }
            if (outBuf.size() > (8u << 20)) {
                out->write(outBuf.data(), (std::streamsize)outBuf.size());
                outBuf.clear();
            }
        }
        out->write(outBuf.data(), (std::streamsize)outBuf.size());
        halgpu_free_result(res);
        if (!deferredError.empty()) { out->flush(); throw std::runtime_error(deferredError); }
        } // batches of the block
    }
    if (pipe) pipe->drain();
}

} // namespace halgpu
