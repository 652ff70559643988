// Multi-threaded, allocation-free BED text layer for the liftover CLI.
//
// The reference parses and prints one line at a time through iostreams (BedLine::read / write,
// liftover/impl/halBedLine.cpp:27-151; BedScanner::scan, halBedScanner.cpp:40-61).  With the mapping itself at
// hundreds of millions of intervals per second on the GPU, that text layer is the wall clock of the CLI, so for the
// common case -- BED3..BED9 (+ pass-through columns), every line of a block of the same width, BED output -- a block
// of input text is tokenised by N threads straight into the (pinned) coordinate arrays of one halgpu_liftover call and
// the result records are formatted by N threads into per-thread text buffers that are written out in input order.
//
// The fast path accepts a strict subset of what BedLine::parse accepts and produces byte-identical text for it;
// ANYTHING else in a block (BED12 blocks, mixed widths, a malformed field, a line ending in TAB ...) makes
// parseBlock() return false and the caller re-runs that block through the serial BedLine code, which owns the
// reference's exact error messages and sticky-field behaviour.
#pragma once
#include "../../../include/halgpu.h"
#include "bed.hpp"
#include <cstdint>
#include <functional>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace halgpu {

void parallelFor(unsigned nThreads, const std::function<void(unsigned)> &fn);

struct FastLine {
    uint64_t off; // offset of the line's first character in the block
    uint32_t len; // characters up to (not including) the newline
    int32_t seq;  // source sequence index
};

struct FastEvent { // a line the reference skips with a message on stderr (halLiftover.cpp:53-69)
    enum Kind : uint8_t { MISSING_SEQUENCE, ENDPOINT_BEYOND_SEQUENCE } kind;
    std::string chrName;
    int64_t end;
    int64_t seqLength;
    uint64_t order; // (thread << 40) | ordinal: events are reported in input order
};

// Growable output text buffer with raw-pointer appends (capacity survives clear(), so blocks after the first reuse
// already-faulted pages).
class TextBuf {
  public:
    TextBuf() = default;
    TextBuf(const TextBuf &) = delete;
    TextBuf &operator=(const TextBuf &) = delete;
    TextBuf(TextBuf &&o) noexcept : _p(o._p), _n(o._n), _cap(o._cap) { o._p = nullptr; o._n = o._cap = 0; }
    ~TextBuf();
    void clear() { _n = 0; }
    const char *data() const { return _p; }
    size_t size() const { return _n; }
    char *room(size_t need) { // returns the write position with at least `need` bytes available
        if (_n + need > _cap) grow(need);
        return _p + _n;
    }
    void advanceTo(char *w) { _n = (size_t)(w - _p); }

  private:
    void grow(size_t need);
    char *_p = nullptr;
    size_t _n = 0, _cap = 0;
};

class FastBedBlock {
  public:
    FastBedBlock(const halgpu_seq *srcSeqs, size_t nSrc, const halgpu_seq *tgtSeqs, size_t nTgt);
    ~FastBedBlock();
    FastBedBlock(const FastBedBlock &) = delete;
    FastBedBlock &operator=(const FastBedBlock &) = delete;

    // Tokenises block[0..n) (whole lines).  `sticky` is the scanner's BedLine before the block: the fields a line of
    // this block's width does not carry keep its values (strand for BED<6, thickEnd for BED7).  Returns false when the
    // block needs the serial path (nothing has been emitted then).
    bool parseBlock(const char *block, size_t n, int forcedBedType, const BedLine &sticky, unsigned nThreads);

    // inputs of the halgpu_liftover call (pinned)
    size_t numIntervals() const { return _n; }
    const int64_t *starts() const { return _gs; }
    const int64_t *endsIncl() const { return _ge; }
    const uint8_t *strands() const { return _st; }
    size_t linesSeen() const { return _linesSeen; } // non-blank lines of the block, lifted or skipped
    int bedType() const { return _bedType; }
    const std::vector<FastEvent> &events() const { return _events; }
    // offset and length of the block's last line (the caller re-parses it serially to carry the sticky state on)
    bool lastLine(uint64_t &off, uint32_t &len) const { off = _lastOff; len = _lastLen; return _linesSeen > 0; }

    // Formats the output lines of the block: text[t] holds the lines of the t-th slice of input lines, in order.
    // Returns the number of output lines.
    size_t formatBlock(const char *block, const halgpu_lift_result *res, unsigned nThreads, std::vector<TextBuf> &text) const;

  private:
    void reservePinned(size_t n);
    std::unordered_map<std::string_view, int32_t> _seqByName;
    const halgpu_seq *_src, *_tgt;
    size_t _nTgt;
    std::vector<FastLine> _lines;
    std::vector<FastEvent> _events;
    int64_t *_gs = nullptr, *_ge = nullptr;
    uint8_t *_st = nullptr;
    size_t _cap = 0, _n = 0, _linesSeen = 0;
    uint64_t _lastOff = 0;
    uint32_t _lastLen = 0;
    int _bedType = 0;
    char _stickyStrand = '+';
    int64_t _stickyThickEnd = 0;
};

} // namespace halgpu
