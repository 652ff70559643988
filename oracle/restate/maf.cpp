/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
 *
 * Per-column restatement of
 *   MafExport::convertSequence / writeHeader         maf/impl/halMafExport.cpp:15-88
 *   MafBlock::resetEntries / initEntry / updateEntry  maf/impl/halMafBlock.cpp:36-132
 *   MafBlock::initBlock / appendColumn / canAppendColumn / printBlock   :294-450, 499-519
 *   ColumnMap ordering and persistence                api/inc/halColumnIterator.h:45-54, halColumnIterator.cpp:192-206,821-825
 * for the default ColumnIterator flags (unique=false, maxRefGap=0, no --printTree).
 */
#include "maf.h"
#include <algorithm>
#include <map>
#include <memory>

namespace oracle {

namespace {

struct SeqKey {
    int genome, seq;
};
struct KeyLess { /* ColumnIterator::SequenceLess: genome NAME bytes, then sequence array index */
    const HalView *v;
    bool operator()(const SeqKey &a, const SeqKey &b) const {
        int d = v->genomes[a.genome].name.compare(v->genomes[b.genome].name);
        return d < 0 || (d == 0 && a.seq < b.seq);
    }
};
inline bool sameKey(const SeqKey &a, const SeqKey &b) { return a.genome == b.genome && a.seq == b.seq; }

struct Entry { /* MafBlockEntry */
    std::string name;
    int genome = -1;
    int64_t start = -1, length = 0, srcLength = 0;
    char strand = '+';
    std::string text;
    int lastUsed = 0;
};

typedef std::map<SeqKey, std::vector<ColRow>, KeyLess> ColMap;
typedef std::multimap<SeqKey, std::unique_ptr<Entry>, KeyLess> Entries;

struct Block {
    const HalView &v;
    const MafOpts &o;
    Entries entries;
    Entry *reference = nullptr;
    int64_t refIndex = -1;

    Block(const HalView &vv, const MafOpts &oo) : v(vv), o(oo), entries(KeyLess{&vv}) {}

    std::string nameOf(const SeqKey &k) const {
        const GenomeView &g = v.genomes[k.genome];
        return o.fullNames ? g.name + "." + g.seqs[k.seq].name : g.seqs[k.seq].name;
    }
    char baseOf(const ColRow &r) const {
        char c = v.genomes[r.genome].base(r.pos);
        if (!r.rev) return c;
        switch (c) { /* reverseComplement(char), api/inc/halCommon.h:45-67 */
        case 'A': return 'T'; case 'a': return 't'; case 'C': return 'G'; case 'c': return 'g';
        case 'G': return 'C'; case 'g': return 'c'; case 'T': return 'A'; case 't': return 'a';
        default: return c;
        }
    }
    void resetEntries() {
        reference = nullptr;
        refIndex = -1;
        for (auto i = entries.begin(); i != entries.end();) {
            Entry *e = i->second.get();
            bool deleted = false;
            if (e->start == -1) {
                if (e->lastUsed > 10) { i = entries.erase(i); deleted = true; }
                else ++e->lastUsed;
            } else {
                e->lastUsed = 0;
            }
            if (!deleted) {
                e->start = -1; e->strand = '+'; e->length = 0; e->text.clear();
                ++i;
            }
        }
    }
    void initEntry(Entry *e, const SeqKey &k, const ColRow *row, bool clearText = true) {
        std::string nm = nameOf(k);
        if (e->name != nm || e->genome != k.genome) {
            e->name = nm; e->genome = k.genome; e->srcLength = v.genomes[k.genome].seqs[k.seq].length;
        }
        if (row) {
            e->start = row->pos - v.genomes[k.genome].seqs[k.seq].start;
            e->length = 0;
            e->strand = row->rev ? '-' : '+';
            if (row->rev) e->start = e->srcLength - 1 - e->start;
        } else {
            e->start = -1; e->length = 0; e->strand = '+';
        }
        if (clearText) e->text.clear();
    }
    void updateEntry(Entry *e, const SeqKey *k, const ColRow *row) {
        if (row) {
            if (e->start == -1) initEntry(e, *k, row, false);
            ++e->length;
            e->text.push_back(baseOf(*row));
        } else {
            e->text.push_back('-');
        }
    }
    void initBlock(const ColMap &cm, const SeqKey &refKey, int64_t refSeqPos) {
        resetEntries();
        auto e = entries.begin();
        for (auto c = cm.begin(); c != cm.end(); ++c) {
            const SeqKey &k = c->first;
            if (c->second.empty()) {
                e = entries.lower_bound(k);
                if (e == entries.end() || !sameKey(e->first, k)) {
                    std::unique_ptr<Entry> ne(new Entry);
                    initEntry(ne.get(), k, nullptr);
                    e = entries.insert(Entries::value_type(k, std::move(ne)));
                } else {
                    initEntry(e->second.get(), k, nullptr);
                }
            } else {
                for (const ColRow &d : c->second) {
                    if (e == entries.begin()) {
                        e = entries.lower_bound(k);
                        if (e == entries.end() || !sameKey(e->first, k)) e = entries.end();
                    } else {
                        while (e != entries.end() && !sameKey(e->first, k)) ++e;
                    }
                    if (e == entries.end()) {
                        std::unique_ptr<Entry> ne(new Entry);
                        initEntry(ne.get(), k, &d);
                        e = entries.insert(Entries::value_type(k, std::move(ne)));
                    } else {
                        initEntry(e->second.get(), k, &d);
                    }
                    ++e;
                }
            }
        }
        if (reference == nullptr) {
            e = entries.lower_bound(refKey);
            if (e == entries.end() || !sameKey(e->first, refKey)) e = entries.begin();
            reference = e->second.get();
            if (sameKey(e->first, refKey)) refIndex = refSeqPos;
        }
    }
    bool canAppend(const ColMap &cm) const {
        auto e = entries.begin();
        for (auto c = cm.begin(); c != cm.end(); ++c) {
            const SeqKey &k = c->first;
            const int64_t seqStart = v.genomes[k.genome].seqs[k.seq].start;
            for (const ColRow &d : c->second) {
                while (e != entries.end() && !sameKey(e->first, k)) ++e;
                if (e == entries.end()) return false;
                const Entry *en = e->second.get();
                if (en->start != -1) {
                    if (en->length >= o.maxBlockLen || (en->length > 0 && (en->strand == '-') != d.rev)) return false;
                    int64_t pos = d.pos - seqStart;
                    if (d.rev) pos = en->srcLength - 1 - pos;
                    if (pos - en->start != en->length) return false;
                }
                ++e;
            }
        }
        return true;
    }
    void append(const ColMap &cm) {
        auto e = entries.begin();
        for (auto c = cm.begin(); c != cm.end(); ++c) {
            for (const ColRow &d : c->second) {
                while (e != entries.end() && !sameKey(e->first, c->first)) { updateEntry(e->second.get(), nullptr, nullptr); ++e; }
                updateEntry(e->second.get(), &c->first, &d);
                ++e;
            }
        }
        for (; e != entries.end(); ++e) updateEntry(e->second.get(), nullptr, nullptr);
    }
    bool referenceIsAllGaps() const {
        if (!reference) return false;
        for (char c : reference->text) if (c != '-') return false;
        return true;
    }
    static void printEntry(std::string &out, const Entry &e, int64_t start) {
        out += "s\t" + e.name + "\t" + std::to_string(start) + "\t" + std::to_string(e.length) + "\t" + e.strand + "\t" +
               std::to_string(e.srcLength) + "\t" + e.text + "\n";
    }
    void print(std::string &out) const {
        out += "a\n";
        if (reference->start == -1) {
            if (refIndex != -1) printEntry(out, *reference, refIndex);
        } else {
            printEntry(out, *reference, reference->start);
        }
        for (auto e = entries.begin(); e != entries.end(); ++e) {
            if (e->second->start != -1 && e->second.get() != reference) printEntry(out, *e->second, e->second->start);
        }
    }
};

void loadColumn(const HalView &v, int ref, int64_t p, const MafOpts &o, ColMap &cm, std::vector<ColRow> &rows) {
    for (auto &kv : cm) kv.second.clear(); /* resetColMap: keys persist */
    column(v, ref, p, o.col, rows);
    for (const ColRow &r : rows) cm[SeqKey{r.genome, v.genomes[r.genome].seqOf(r.pos)}].push_back(r);
}

void convertSequence(const HalView &v, Block &blk, int ref, int seq, int64_t start, int64_t length, const MafOpts &o, std::string &out) {
    const SeqView &S = v.genomes[ref].seqs[seq];
    if (start >= S.length || start + length > S.length) throw std::runtime_error("Invalid range specified for convertGenome");
    if (length == 0) length = S.length - start;
    if (length == 0) throw std::runtime_error("Cannot convert zero length sequence");
    if (out.empty()) out += "##maf version=1 scoring=N/A\n# hal " + v.newick + "\n\n";
    ColMap cm(KeyLess{&v});
    std::vector<ColRow> rows;
    const SeqKey refKey{ref, seq};
    int64_t appendCount = 0;
    size_t numBlocks = 0;
    for (int64_t i = 0; i < length; ++i) {
        const int64_t sp = start + i;
        if (o.unique) {
            /* ColumnIterator(unique = true) + MafExport's isCanonicalOnRef test (maf/impl/halMafExport.cpp:52,61;
             * api/impl/halColumnIterator.cpp:208-212, 749-762, 771-818).  Every walked column puts its reference-genome bases
             * right of the window start into the visit cache and nextFreeIndex skips cached positions WITHOUT walking them:
             * a position is walked iff no reference-genome row of its column lies in [window start, position).  A walked
             * column is written iff its left-most reference-genome base is not left of the window start -- but it has been
             * walked, so its sequences are ColumnMap keys from then on (which matters to canAppendColumn). */
            const int64_t w0 = S.start + start, me = S.start + sp;
            column(v, ref, me, o.col, rows);
            bool walked = true, canonical = true;
            for (const ColRow &r : rows) {
                if (r.genome != ref) continue;
                walked &= !(r.pos >= w0 && r.pos < me);
                canonical &= !(r.pos < w0);
            }
            if (!walked) continue;
            loadColumn(v, ref, me, o, cm, rows);
            if (!canonical) continue;
        } else {
            loadColumn(v, ref, S.start + sp, o, cm, rows);
        }
        if (appendCount == 0) {
            blk.initBlock(cm, refKey, sp);
        } else if (!blk.canAppend(cm)) {
            if (numBlocks++ % 1000 == 0) { /* defragment: drop ColumnMap keys without rows */
                for (auto it = cm.begin(); it != cm.end();) { if (it->second.empty()) it = cm.erase(it); else ++it; }
            }
            if (appendCount > 0 && (o.keepEmptyRefBlocks || !blk.referenceIsAllGaps())) { blk.print(out); out += "\n"; }
            blk.initBlock(cm, refKey, sp);
        }
        blk.append(cm);
        ++appendCount;
    }
    if (appendCount > 0 && (o.keepEmptyRefBlocks || !blk.referenceIsAllGaps())) { blk.print(out); out += "\n"; }
}

} // namespace

void hal2maf(const HalView &v, int ref, int refSeq, int64_t start, int64_t length, const MafOpts &o, std::string &out) {
    Block blk(v, o);
    if (refSeq >= 0) {
        convertSequence(v, blk, ref, refSeq, start, length, o, out);
    } else {
        for (size_t s = 0; s < v.genomes[ref].seqs.size(); ++s) convertSequence(v, blk, ref, (int)s, start, length, o, out);
    }
}

} // namespace oracle
