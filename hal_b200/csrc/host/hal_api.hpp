// hal_api.hpp -- read-only hal::Alignment / Genome / Sequence / Top|BottomSegmentIterator query surface over a
// HAL-MMAP file, with the reference's names and semantics for everything the liftover / column path touches
// (api/inc/halAlignment.h, halGenome.h, halSequence.h, halSegmentIterator.h, halTopSegmentIterator.h,
// halBottomSegmentIterator.h).  Host code written against the reference's iterators (toSite / toRight / slice /
// toParent / toChild / toParseUp / toParseDown / toNextParalogy ...) compiles against this header unchanged in
// spirit: it is the CPU-side companion of the staged GPU index, NOT a compute fallback -- nothing in the GPU path
// calls it; tools use it to inspect alignments and to prepare / interpret batches.
//
// Semantics restated from api/impl/halSegmentIterator.cpp:46-299, halTopSegmentIterator.cpp:36-107,
// halBottomSegmentIterator.cpp:40-94, api/mmap_impl/mmapTopSegment.h:60-90, mmapBottomSegment.h:60-100.
#pragma once
#include "../halmmap.hpp"
#include <algorithm>
#include <cstring>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace hal {

typedef int64_t hal_index_t;
typedef uint64_t hal_size_t;
typedef uint64_t hal_offset_t;
static const hal_index_t NULL_INDEX = -1;

class hal_exception : public std::runtime_error {
  public:
    explicit hal_exception(const std::string &m) : std::runtime_error(m) {}
};

class Alignment;
class Genome;
class TopSegmentIterator;
class BottomSegmentIterator;
typedef std::shared_ptr<TopSegmentIterator> TopSegmentIteratorPtr;
typedef std::shared_ptr<BottomSegmentIterator> BottomSegmentIteratorPtr;
typedef std::shared_ptr<const Alignment> AlignmentConstPtr;

class Sequence {
  public:
    Sequence(const Genome *g, int idx, const halgpu::SequenceInfo *i) : _genome(g), _index(idx), _info(i) {}
    const std::string &getName() const { return _info->name; }
    std::string getFullName() const;
    const Genome *getGenome() const { return _genome; }
    hal_index_t getStartPosition() const { return _info->start; }
    hal_index_t getEndPosition() const { return _info->start + _info->length - 1; }
    hal_index_t getArrayIndex() const { return _index; }
    hal_size_t getSequenceLength() const { return (hal_size_t)_info->length; }
    hal_size_t getNumTopSegments() const { return (hal_size_t)_info->numTop; }
    hal_size_t getNumBottomSegments() const { return (hal_size_t)_info->numBottom; }
    hal_index_t getTopSegmentArrayIndex() const { return _info->topFirst; }
    hal_index_t getBottomSegmentArrayIndex() const { return _info->bottomFirst; }
    void getSubString(std::string &out, hal_size_t start, hal_size_t length) const;
    void getString(std::string &out) const { getSubString(out, 0, getSequenceLength()); }

  private:
    const Genome *_genome;
    int _index;
    const halgpu::SequenceInfo *_info;
};

class Genome {
  public:
    Genome(const Alignment *a, int id, const halgpu::GenomeInfo *i) : _alignment(a), _id(id), _info(i) {
        for (size_t s = 0; s < i->sequences.size(); ++s) _seqs.emplace_back(this, (int)s, &i->sequences[s]);
    }
    const std::string &getName() const { return _info->name; }
    const Alignment *getAlignment() const { return _alignment; }
    int getArrayIndex() const { return _id; }
    hal_size_t getSequenceLength() const { return (hal_size_t)_info->length; }
    hal_size_t getNumSequences() const { return _seqs.size(); }
    hal_size_t getNumTopSegments() const { return (hal_size_t)_info->numTop; }
    hal_size_t getNumBottomSegments() const { return (hal_size_t)_info->numBottom; }
    hal_size_t getNumChildren() const { return _info->children.size(); }
    const Genome *getParent() const;
    const Genome *getChild(hal_size_t i) const;
    hal_index_t getChildIndex(const Genome *child) const { // api/inc/halGenome.h:247-256
        for (size_t i = 0; i < _info->children.size(); ++i)
            if (getChild(i) == child) return (hal_index_t)i;
        return NULL_INDEX;
    }
    const Sequence *getSequence(const std::string &name) const {
        auto it = _info->sequenceByName.find(name);
        return it == _info->sequenceByName.end() ? nullptr : &_seqs[it->second];
    }
    const Sequence *getSequenceByIndex(hal_index_t i) const { return &_seqs[(size_t)i]; }
    const Sequence *getSequenceBySite(hal_size_t pos) const { // mmapGenomeSiteMap.cpp:99-113 (same answer)
        if ((hal_index_t)pos >= _info->length) return nullptr;
        size_t lo = 0, hi = _seqs.size();
        while (hi - lo > 1) { size_t m = (lo + hi) / 2; if (_info->sequences[m].start <= (hal_index_t)pos) lo = m; else hi = m; }
        return &_seqs[lo];
    }
    TopSegmentIteratorPtr getTopSegmentIterator(hal_index_t segmentIndex = 0) const;
    BottomSegmentIteratorPtr getBottomSegmentIterator(hal_index_t segmentIndex = 0) const;
    void getSubString(std::string &out, hal_size_t start, hal_size_t length) const {
        static const char tbl[16] = {'a', 'c', 'g', 't', 'n', 0, 0, 0, 'A', 'C', 'G', 'T', 'N', 0, 0, 0}; // halCommon.cpp:233-235
        out.resize(length);
        for (hal_size_t i = 0; i < length; ++i) {
            const hal_size_t p = start + i;
            const uint8_t b = _info->dna[p >> 1];
            out[i] = tbl[(p & 1) ? (b & 0xF) : (b >> 4)];
        }
    }
    void getString(std::string &out) const { getSubString(out, 0, getSequenceLength()); }
    const halgpu::GenomeInfo &info() const { return *_info; }

  private:
    const Alignment *_alignment;
    int _id;
    const halgpu::GenomeInfo *_info;
    std::vector<Sequence> _seqs;
};

inline std::string Sequence::getFullName() const { return _genome->getName() + "." + _info->name; }
inline void Sequence::getSubString(std::string &out, hal_size_t start, hal_size_t length) const {
    _genome->getSubString(out, (hal_size_t)_info->start + start, length);
}

class Alignment {
  public:
    explicit Alignment(const std::string &path) : _file(path) {
        for (size_t g = 0; g < _file.genomes().size(); ++g) _genomes.emplace_back(new Genome(this, (int)g, &_file.genomes()[g]));
    }
    hal_size_t getNumGenomes() const { return _genomes.size(); }
    std::string getRootName() const { return _file.genomes()[_file.root()].name; }
    std::string getParentName(const std::string &name) const {
        const Genome *g = openGenome(name);
        if (g == nullptr) throw hal_exception("genome " + name + " not found in alignment.");
        return g->getParent() ? g->getParent()->getName() : "";
    }
    std::vector<std::string> getChildNames(const std::string &name) const {
        const Genome *g = openGenome(name);
        if (g == nullptr) throw hal_exception("genome " + name + " not found in alignment.");
        std::vector<std::string> out;
        for (hal_size_t i = 0; i < g->getNumChildren(); ++i) out.push_back(g->getChild(i)->getName());
        return out;
    }
    const Genome *openGenome(const std::string &name) const {
        int id = _file.genomeId(name);
        return id < 0 ? nullptr : _genomes[id].get();
    }
    const Genome *genomeByIndex(int id) const { return _genomes[id].get(); }
    void closeGenome(const Genome *) const {} // mmap: intentionally blank (api/mmap_impl/mmapAlignment.h:90-92)
    std::string getNewickTree() const { return _file.newick(); }
    std::string getVersion() const { return _file.version(); }
    const std::string &getStorageFormat() const { static const std::string f = "mmap"; return f; }
    bool isReadOnly() const { return true; }
    const halgpu::HalFile &file() const { return _file; }

  private:
    halgpu::HalFile _file;
    std::vector<std::unique_ptr<Genome>> _genomes;
};

inline const Genome *Genome::getParent() const { return _info->parent < 0 ? nullptr : _alignment->genomeByIndex(_info->parent); }
inline const Genome *Genome::getChild(hal_size_t i) const {
    return i < _info->children.size() ? _alignment->genomeByIndex(_info->children[i]) : nullptr;
}

inline AlignmentConstPtr openHalAlignment(const std::string &path) {
    try {
        return AlignmentConstPtr(new Alignment(path));
    } catch (halgpu::HalError &e) {
        throw hal_exception(e.what());
    }
}

/* getLowestCommonAncestor / getGenomesInSpanningTree / getGenomesInSubTree (api/impl/halCommon.cpp:123-195) */
inline const Genome *getLowestCommonAncestor(const std::set<const Genome *> &in) {
    if (in.empty()) return nullptr;
    const Alignment *a = (*in.begin())->getAlignment();
    int m = (*in.begin())->getArrayIndex();
    for (const Genome *g : in) m = a->file().mrca(m, g->getArrayIndex());
    return a->genomeByIndex(m);
}
inline void getGenomesInSpanningTree(const std::set<const Genome *> &in, std::set<const Genome *> &out) {
    const Genome *m = getLowestCommonAncestor(in);
    for (const Genome *g : in)
        for (const Genome *x = g;; x = x->getParent()) { out.insert(x); if (x == m) break; }
}
inline void getGenomesInSubTree(const Genome *root, std::set<const Genome *> &out) {
    out.insert(root);
    for (hal_size_t i = 0; i < root->getNumChildren(); ++i) getGenomesInSubTree(root->getChild(i), out);
}

// ---- iterators -------------------------------------------------------------------------------------------
class SegmentIterator {
  public:
    virtual ~SegmentIterator() {}
    virtual bool isTop() const = 0;
    const Genome *getGenome() const { return _genome; }
    hal_index_t getArrayIndex() const { return _index; }
    void setArrayIndex(const Genome *g, hal_index_t i) { _genome = g; _index = i; }
    hal_offset_t getStartOffset() const { return _startOffset; }
    hal_offset_t getEndOffset() const { return _endOffset; }
    bool getReversed() const { return _reversed; }
    hal_size_t getNumSegmentsInGenome() const { return isTop() ? _genome->getNumTopSegments() : _genome->getNumBottomSegments(); }
    bool inRange() const { return _index >= 0 && _index < (hal_index_t)getNumSegmentsInGenome(); }
    bool atEnd() const { return !inRange(); }
    // the underlying (unsliced) segment
    hal_index_t segStart() const { return rd(recordPtr(_index), 0); }
    hal_size_t segLength() const { return (hal_size_t)(rd(recordPtr(_index + 1), 0) - segStart()); } // +1 sentinel, mmapTopSegment.h:78-80
    const Sequence *getSequence() const { return _genome->getSequenceBySite((hal_size_t)segStart()); }
    // sliced view (halSegmentIterator.cpp:46-68)
    hal_index_t getStartPosition() const {
        return !_reversed ? segStart() + (hal_index_t)_startOffset : segStart() + (hal_index_t)segLength() - (hal_index_t)_startOffset - 1;
    }
    hal_index_t getEndPosition() const {
        return !_reversed ? getStartPosition() + (hal_index_t)(getLength() - 1) : getStartPosition() - (hal_index_t)(getLength() - 1);
    }
    hal_size_t getLength() const { return segLength() - _endOffset - _startOffset; }
    void getString(std::string &out) const {
        _genome->getSubString(out, (hal_size_t)segStart(), segLength());
        if (_reversed) reverseComplement(out);
        out = out.substr(_startOffset, getLength());
    }
    bool leftOf(hal_index_t pos) const {
        return !_reversed ? (hal_index_t)(getStartPosition() + getLength()) <= pos : getStartPosition() < pos;
    }
    bool rightOf(hal_index_t pos) const {
        return !_reversed ? getStartPosition() > pos : getStartPosition() - (hal_index_t)getLength() >= pos;
    }
    bool overlaps(hal_index_t pos) const { return !leftOf(pos) && !rightOf(pos); }
    bool isFirst() const { return !_reversed ? _index == 0 : _index == (hal_index_t)getNumSegmentsInGenome() - 1; }
    bool isLast() const { return !_reversed ? _index == (hal_index_t)getNumSegmentsInGenome() - 1 : _index == 0; }
    void toReverse() { _reversed = !_reversed; }
    void toReverseInPlace() { _reversed = !_reversed; std::swap(_startOffset, _endOffset); }
    void slice(hal_offset_t so = 0, hal_offset_t eo = 0) { _startOffset = so; _endOffset = eo; }
    void toLeft(hal_index_t leftCutoff = NULL_INDEX) { // halSegmentIterator.cpp:177-206
        if (!_reversed) {
            if (_startOffset == 0) { --_index; _endOffset = 0; }
            else { _endOffset = segLength() - _startOffset; _startOffset = 0; }
            if (_index >= 0 && leftCutoff != NULL_INDEX && overlaps(leftCutoff)) _startOffset = (hal_offset_t)(leftCutoff - segStart());
        } else {
            if (_startOffset == 0) { ++_index; _endOffset = 0; }
            else { _endOffset = segLength() - _startOffset; _startOffset = 0; }
            if ((hal_size_t)_index < getNumSegmentsInGenome() && leftCutoff != NULL_INDEX && overlaps(leftCutoff))
                _startOffset = (hal_offset_t)(segStart() + (hal_index_t)segLength() - 1 - leftCutoff);
        }
    }
    void toRight(hal_index_t rightCutoff = NULL_INDEX) { // halSegmentIterator.cpp:208-238
        if (!_reversed) {
            if (_endOffset == 0) { ++_index; _startOffset = 0; }
            else { _startOffset = segLength() - _endOffset; _endOffset = 0; }
            if ((hal_size_t)_index < getNumSegmentsInGenome() && rightCutoff != NULL_INDEX && overlaps(rightCutoff))
                _endOffset = (hal_offset_t)(segStart() + (hal_index_t)segLength() - rightCutoff - 1);
        } else {
            if (_endOffset == 0) { --_index; _startOffset = 0; }
            else { _startOffset = segLength() - _endOffset; _endOffset = 0; }
            if (_index >= 0 && rightCutoff != NULL_INDEX && overlaps(rightCutoff)) _endOffset = (hal_offset_t)(rightCutoff - segStart());
        }
    }
    void toSite(hal_index_t position, bool doSlice = true) { // result of halSegmentIterator.cpp:240-299 (any exact search)
        const hal_index_t len = (hal_index_t)_genome->getSequenceLength(), nseg = (hal_index_t)getNumSegmentsInGenome();
        _startOffset = _endOffset = 0;
        if (position < 0) { _index = NULL_INDEX; return; }
        if (position >= len) { _index = len; return; }
        hal_index_t lo = 0, hi = nseg;
        while (hi - lo > 1) { hal_index_t m = (lo + hi) / 2; if (rd(recordPtr(m), 0) <= position) lo = m; else hi = m; }
        _index = lo;
        if (doSlice) {
            _startOffset = (hal_offset_t)(position - segStart());
            _endOffset = (hal_offset_t)(segStart() + (hal_index_t)segLength() - position - 1);
        }
    }
    static void reverseComplement(std::string &s) { // api/impl/halCommon.cpp:56-75
        std::reverse(s.begin(), s.end());
        for (char &c : s) {
            switch (c) {
            case 'A': c = 'T'; break; case 'a': c = 't'; break; case 'C': c = 'G'; break; case 'c': c = 'g'; break;
            case 'G': c = 'C'; break; case 'g': c = 'c'; break; case 'T': c = 'A'; break; case 't': c = 'a'; break;
            default: break;
            }
        }
    }

  protected:
    SegmentIterator(const Genome *g, hal_index_t i) : _genome(g), _index(i) {}
    virtual const uint8_t *recordPtr(hal_index_t i) const = 0;
    static int64_t rd(const uint8_t *p, size_t off) { int64_t v; std::memcpy(&v, p + off, 8); return v; }
    const Genome *_genome;
    hal_index_t _index;
    hal_offset_t _startOffset = 0, _endOffset = 0;
    bool _reversed = false;
};

class TopSegmentIterator : public SegmentIterator {
  public:
    TopSegmentIterator(const Genome *g, hal_index_t i) : SegmentIterator(g, i) {}
    bool isTop() const override { return true; }
    // TopSegment accessors (api/inc/halTopSegment.h); tseg() mirrors `it->tseg()->...`
    const TopSegmentIterator *tseg() const { return this; }
    const TopSegmentIterator *getTopSegment() const { return this; }
    hal_index_t getBottomParseIndex() const { return rd(recordPtr(_index), 8); }
    hal_index_t getNextParalogyIndex() const { return rd(recordPtr(_index), 16); }
    hal_index_t getParentIndex() const { return rd(recordPtr(_index), 24); }
    bool getParentReversed() const { return recordPtr(_index)[32] != 0; }
    bool hasParent() const { return getParentIndex() != NULL_INDEX; }
    bool hasParseDown() const { return getBottomParseIndex() != NULL_INDEX; }
    bool hasNextParalogy() const { return getNextParalogyIndex() != NULL_INDEX; }
    bool isCanonicalParalog() const; // api/mmap_impl/mmapTopSegment.cpp:30-40
    void toChild(const BottomSegmentIteratorPtr &bot, hal_size_t child);
    void toChildG(const BottomSegmentIteratorPtr &bot, const Genome *childGenome);
    void toParseUp(const BottomSegmentIteratorPtr &bot);
    void toNextParalogy() { // halTopSegmentIterator.cpp:99-107
        const bool rev = getParentReversed();
        _index = getNextParalogyIndex();
        if (getParentReversed() != rev) toReverse();
    }
    TopSegmentIteratorPtr clone() const { return TopSegmentIteratorPtr(new TopSegmentIterator(*this)); }
    void copy(const TopSegmentIteratorPtr &o) { *this = *o; }

  protected:
    const uint8_t *recordPtr(hal_index_t i) const override { return _genome->info().top + 40 * i; }
};

class BottomSegmentIterator : public SegmentIterator {
  public:
    BottomSegmentIterator(const Genome *g, hal_index_t i) : SegmentIterator(g, i) {}
    bool isTop() const override { return false; }
    const BottomSegmentIterator *bseg() const { return this; }
    const BottomSegmentIterator *getBottomSegment() const { return this; }
    hal_size_t getNumChildren() const { return _genome->getNumChildren(); }
    hal_index_t getTopParseIndex() const { return rd(recordPtr(_index), 8); }
    bool hasParseUp() const { return getTopParseIndex() != NULL_INDEX; }
    hal_index_t getChildIndex(hal_size_t i) const { return rd(recordPtr(_index), 16 + 8 * i); }
    hal_index_t getChildIndexG(const Genome *child) const { return getChildIndex((hal_size_t)_genome->getChildIndex(child)); }
    bool hasChild(hal_size_t i) const { return getChildIndex(i) != NULL_INDEX; }
    bool hasChildG(const Genome *child) const { return getChildIndexG(child) != NULL_INDEX; }
    bool getChildReversed(hal_size_t i) const { return recordPtr(_index)[16 + 8 * getNumChildren() + i] != 0; }
    void toParent(const TopSegmentIteratorPtr &top) { // halBottomSegmentIterator.cpp:40-49
        _genome = top->getGenome()->getParent();
        _index = top->getParentIndex();
        _startOffset = top->getStartOffset(); _endOffset = top->getEndOffset(); _reversed = top->getReversed();
        if (top->getParentReversed()) toReverse();
    }
    void toParseDown(const TopSegmentIteratorPtr &top) { // halBottomSegmentIterator.cpp:51-76
        _genome = top->getGenome();
        _index = top->getBottomParseIndex();
        _reversed = top->getReversed();
        const hal_index_t startPos = top->getStartPosition();
        while (startPos >= segStart() + (hal_index_t)segLength()) ++_index;
        if (!_reversed) {
            _startOffset = (hal_offset_t)(startPos - segStart());
            const hal_index_t myEnd = segStart() + (hal_index_t)segLength(), otherEnd = top->getStartPosition() + (hal_index_t)top->getLength();
            _endOffset = (hal_offset_t)std::max<hal_index_t>(0, myEnd - otherEnd);
        } else {
            _startOffset = (hal_offset_t)(segStart() + (hal_index_t)segLength() - 1 - startPos);
            const hal_index_t myEnd = segStart(), otherEnd = top->getStartPosition() - (hal_index_t)top->getLength() + 1;
            _endOffset = (hal_offset_t)std::max<hal_index_t>(0, otherEnd - myEnd);
        }
    }
    BottomSegmentIteratorPtr clone() const { return BottomSegmentIteratorPtr(new BottomSegmentIterator(*this)); }
    void copy(const BottomSegmentIteratorPtr &o) { *this = *o; }

  protected:
    const uint8_t *recordPtr(hal_index_t i) const override { return _genome->info().bottom + _genome->info().bottomStride * (size_t)i; }
};

inline bool TopSegmentIterator::isCanonicalParalog() const {
    if (!hasParent()) return false;
    const Genome *p = _genome->getParent();
    BottomSegmentIterator b(p, getParentIndex());
    return b.getChildIndex((hal_size_t)p->getChildIndex(_genome)) == _index;
}
inline void TopSegmentIterator::toChild(const BottomSegmentIteratorPtr &bot, hal_size_t child) { // halTopSegmentIterator.cpp:36-45
    _genome = bot->getGenome()->getChild(child);
    _index = bot->getChildIndex(child);
    _startOffset = bot->getStartOffset(); _endOffset = bot->getEndOffset(); _reversed = bot->getReversed();
    if (bot->getChildReversed(child)) toReverse();
}
inline void TopSegmentIterator::toChildG(const BottomSegmentIteratorPtr &bot, const Genome *childGenome) {
    toChild(bot, (hal_size_t)bot->getGenome()->getChildIndex(childGenome));
}
inline void TopSegmentIterator::toParseUp(const BottomSegmentIteratorPtr &bot) { // halTopSegmentIterator.cpp:55-81
    _genome = bot->getGenome();
    _index = bot->getTopParseIndex();
    _reversed = bot->getReversed();
    const hal_index_t startPos = bot->getStartPosition();
    while (startPos >= segStart() + (hal_index_t)segLength()) ++_index;
    if (!_reversed) {
        _startOffset = (hal_offset_t)(startPos - segStart());
        const hal_index_t myEnd = segStart() + (hal_index_t)segLength(), otherEnd = bot->getStartPosition() + (hal_index_t)bot->getLength();
        _endOffset = (hal_offset_t)std::max<hal_index_t>(0, myEnd - otherEnd);
    } else {
        _startOffset = (hal_offset_t)(segStart() + (hal_index_t)segLength() - 1 - startPos);
        const hal_index_t myEnd = segStart(), otherEnd = bot->getStartPosition() - (hal_index_t)bot->getLength() + 1;
        _endOffset = (hal_offset_t)std::max<hal_index_t>(0, otherEnd - myEnd);
    }
}
inline TopSegmentIteratorPtr Genome::getTopSegmentIterator(hal_index_t i) const { return TopSegmentIteratorPtr(new TopSegmentIterator(this, i)); }
inline BottomSegmentIteratorPtr Genome::getBottomSegmentIterator(hal_index_t i) const { return BottomSegmentIteratorPtr(new BottomSegmentIterator(this, i)); }

} // namespace hal
