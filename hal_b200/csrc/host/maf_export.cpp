#include "maf_export.hpp"
#include <algorithm>
#include <charconv>
#include <chrono>
#include <stdexcept>

namespace halgpu {

namespace {

struct Key { // ColumnIterator::SequenceLess order: genome name bytes, then sequence array index
    int32_t rank, seq;
    bool operator<(const Key &o) const { return rank < o.rank || (rank == o.rank && seq < o.seq); }
    bool operator==(const Key &o) const { return rank == o.rank && seq == o.seq; }
    bool operator!=(const Key &o) const { return !(*this == o); }
};

struct Entry { // MafBlockEntry (maf/inc/halMafBlock.h:70-115)
    const std::string *name = nullptr; // cached "Genome.Sequence" (or bare sequence name)
    std::string text;
    int genome = -1, seq = -1;
    int64_t start = -1, length = 0, srcLength = 0;
    char strand = '+';
    int lastUsed = 0;
};

struct Slot {
    Key key;
    std::unique_ptr<Entry> e;
};

struct Row { // one row of the current column
    Key key;
    int genome;
    int64_t pos; // forward genome coordinate in this column
    bool rev;
};

const char NIB[16] = {'a', 'c', 'g', 't', 'n', '?', '?', '?', 'A', 'C', 'G', 'T', 'N', '?', '?', '?'};       // dnaUnpack, halCommon.cpp:224-235
const char NIBRC[16] = {'t', 'g', 'c', 'a', 'n', '?', '?', '?', 'T', 'G', 'C', 'A', 'N', '?', '?', '?'};     // + reverseComplement(char)

} // namespace

struct GpuMafExport::Impl {
    halgpu_ctx *ctx;
    int numGenomes = 0;
    std::vector<int> rank;                         // genome -> name rank
    std::vector<int> genomeOfRank;
    std::vector<const halgpu_seq *> seqs;          // per genome
    std::vector<size_t> nseqs;
    std::vector<const uint8_t *> dna;
    std::vector<std::string> gname;
    // MafBlock state (persists across convertSequence calls: MafExport::_mafBlock is a member)
    std::vector<Slot> entries;                     // multimap<Sequence*, Entry*> in key order, equal keys in insertion order
    Entry *reference = nullptr;
    int64_t refIndex = -1;
    bool fullNames = true;
    int64_t maxLength = 1000;
    // ColumnIterator state of the current convertSequence call
    std::vector<Key> colKeys;                      // ColumnMap keys (sorted), persist until defragment
    std::vector<Row> rows;                         // current column, ColumnMap order
    std::vector<std::pair<Entry *, int>> pairing;  // appendColumn walk: entry -> row index or -1

    std::vector<std::vector<std::string>> names;   // per genome, per sequence (filled on first use)
    const std::string *nameOf(const Key &k) {
        const int g = genomeOfRank[k.rank];
        if (names.empty()) names.resize(numGenomes);
        if (names[g].empty()) {
            names[g].reserve(nseqs[g]);
            for (size_t s = 0; s < nseqs[g]; ++s) names[g].push_back(fullNames ? gname[g] + "." + seqs[g][s].name : std::string(seqs[g][s].name));
        }
        return &names[g][k.seq];
    }
    size_t lowerBound(const Key &k) const {
        size_t lo = 0, hi = entries.size();
        while (lo < hi) { size_t m = (lo + hi) / 2; if (entries[m].key < k) lo = m + 1; else hi = m; }
        return lo;
    }
    size_t insertEntry(const Key &k, std::unique_ptr<Entry> e) { // multimap::insert: after the last equal key
        size_t i = lowerBound(k);
        while (i < entries.size() && entries[i].key == k) ++i;
        Slot s;
        s.key = k; s.e = std::move(e);
        entries.insert(entries.begin() + i, std::move(s));
        return i;
    }
    void resetEntries() { // MafBlock::resetEntries (maf/impl/halMafBlock.cpp:36-79)
        reference = nullptr;
        refIndex = -1;
        size_t w = 0;
        for (size_t i = 0; i < entries.size(); ++i) {
            Entry *e = entries[i].e.get();
            bool deleted = false;
            if (e->start == -1) {
                if (e->lastUsed > 10) deleted = true; else ++e->lastUsed;
            } else {
                e->lastUsed = 0;
            }
            if (!deleted) {
                e->start = -1; e->strand = '+'; e->length = 0; e->text.clear();
                if (w != i) entries[w] = std::move(entries[i]);
                ++w;
            }
        }
        entries.resize(w);
    }
    void initEntry(Entry *e, const Key &k, const Row *row, bool clearText = true) { // :81-107
        const int g = genomeOfRank[k.rank];
        if (e->genome != g || e->seq != k.seq || e->name == nullptr) {
            e->name = nameOf(k); e->genome = g; e->seq = k.seq; e->srcLength = seqs[g][k.seq].length;
        }
        if (row != nullptr) {
            e->start = row->pos - seqs[g][k.seq].start;
            e->length = 0;
            e->strand = row->rev ? '-' : '+';
            if (row->rev) e->start = e->srcLength - 1 - e->start;
        } else {
            e->start = -1; e->length = 0; e->strand = '+';
        }
        if (clearText) e->text.clear();
    }
    void initBlock(const Key &refKey, int64_t refSeqPos) { // MafBlock::initBlock (:294-368)
        resetEntries();
        size_t e = 0; // cursor into entries (== iterator `e`)
        size_t r = 0, ck = 0;
        // merged walk over the ColumnMap: persistent keys, some of which have rows in this column
        while (ck < colKeys.size()) {
            const Key k = colKeys[ck++];
            size_t r1 = r;
            while (r1 < rows.size() && rows[r1].key == k) ++r1;
            if (r1 == r) { // no rows for this sequence: give it (or re-initialise) a blank entry, cursor stays ON it
                e = lowerBound(k);
                if (e == entries.size() || entries[e].key != k) {
                    std::unique_ptr<Entry> ne(new Entry);
                    initEntry(ne.get(), k, nullptr);
                    e = insertEntry(k, std::move(ne));
                } else {
                    initEntry(entries[e].e.get(), k, nullptr);
                }
            } else {
                for (; r < r1; ++r) {
                    if (e == 0) {
                        e = lowerBound(k);
                        if (e == entries.size() || entries[e].key != k) e = entries.size();
                    } else {
                        while (e < entries.size() && entries[e].key != k) ++e;
                    }
                    if (e == entries.size()) {
                        std::unique_ptr<Entry> ne(new Entry);
                        initEntry(ne.get(), k, &rows[r]);
                        e = insertEntry(k, std::move(ne));
                    } else {
                        initEntry(entries[e].e.get(), k, &rows[r]);
                    }
                    ++e;
                }
            }
        }
        if (reference == nullptr) {
            size_t i = lowerBound(refKey);
            if (i == entries.size() || entries[i].key != refKey) i = 0;
            reference = entries[i].e.get();
            if (entries[i].key == refKey) refIndex = refSeqPos;
        }
    }
    bool canAppend() const { // MafBlock::canAppendColumn (:401-450)
        size_t e = 0;
        for (const Row &d : rows) {
            while (e < entries.size() && entries[e].key != d.key) ++e;
            if (e == entries.size()) return false;
            const Entry *en = entries[e].e.get();
            if (en->start != -1) {
                if (en->length >= maxLength || (en->length > 0 && (en->strand == '-') != d.rev)) return false;
                int64_t pos = d.pos - seqs[d.genome][d.key.seq].start;
                if (d.rev) pos = en->srcLength - 1 - pos;
                if (pos - en->start != en->length) return false;
            }
            ++e;
        }
        return true;
    }
    void computePairing() { // the walk of MafBlock::appendColumn (:370-395)
        pairing.clear();
        size_t e = 0;
        for (size_t r = 0; r < rows.size(); ++r) {
            while (e < entries.size() && entries[e].key != rows[r].key) { pairing.emplace_back(entries[e].e.get(), -1); ++e; }
            pairing.emplace_back(entries[e].e.get(), (int)r);
            ++e;
        }
        for (; e < entries.size(); ++e) pairing.emplace_back(entries[e].e.get(), -1);
    }
    void appendBases(std::string &text, int g, int64_t pos, bool rev, int64_t count) const { // DnaIterator::getBase x count
        const uint8_t *d = dna[g];
        const size_t at = text.size();
        text.resize(at + (size_t)count);
        char *o = &text[at];
        if (!rev) {
            for (int64_t i = 0; i < count; ++i) { const int64_t p = pos + i; const uint8_t b = d[p >> 1]; o[i] = NIB[(p & 1) ? (b & 0xF) : (b >> 4)]; }
        } else {
            for (int64_t i = 0; i < count; ++i) { const int64_t p = pos - i; const uint8_t b = d[p >> 1]; o[i] = NIBRC[(p & 1) ? (b & 0xF) : (b >> 4)]; }
        }
    }
    // append `count` consecutive columns starting with the current one (rows hold the current column)
    void appendColumns(int64_t count) {
        for (auto &pr : pairing) {
            Entry *e = pr.first;
            if (pr.second >= 0) {
                const Row &d = rows[pr.second];
                if (e->start == -1) initEntry(e, d.key, &d, false); // updateEntry (:109-113): keeps the accumulated '-'
                e->length += count;
                appendBases(e->text, d.genome, d.pos, d.rev, count);
            } else {
                e->text.append((size_t)count, '-');
            }
        }
    }
    int64_t capacity() const { // columns that can still be appended before canAppendColumn trips on maxBlockLen
        int64_t cap = INT64_MAX;
        for (auto &pr : pairing)
            if (pr.second >= 0) cap = std::min(cap, maxLength - pr.first->length);
        return cap < 0 ? 0 : cap;
    }
    bool referenceIsAllGaps() const {
        if (reference == nullptr) return false;
        for (char c : reference->text) if (c != '-') return false;
        return true;
    }
    static void printEntry(std::string &out, const Entry &e, int64_t start) { // operator<<(MafBlockEntry) (:452-456)
        char buf[24];
        out += "s\t"; out += *e.name; out += '\t';
        auto r = std::to_chars(buf, buf + sizeof buf, start); out.append(buf, r.ptr); out += '\t';
        r = std::to_chars(buf, buf + sizeof buf, e.length); out.append(buf, r.ptr); out += '\t';
        out += e.strand; out += '\t';
        r = std::to_chars(buf, buf + sizeof buf, e.srcLength); out.append(buf, r.ptr); out += '\t';
        out += e.text; out += '\n';
    }
    void printBlock(std::string &out) const { // MafBlock::printBlock (:499-519)
        out += "a\n";
        if (reference->start == -1) {
            if (refIndex != -1) printEntry(out, *reference, refIndex);
        } else {
            printEntry(out, *reference, reference->start);
        }
        for (const Slot &s : entries) {
            if (s.e->start != -1 && s.e.get() != reference) printEntry(out, *s.e, s.e->start);
        }
    }
    void defragment() { // ColumnIterator::defragment (api/impl/halColumnIterator.cpp:192-206): drop keys without rows
        std::vector<Key> keep;
        size_t r = 0;
        for (const Key &k : colKeys) {
            while (r < rows.size() && rows[r].key < k) ++r;
            if (r < rows.size() && rows[r].key == k) keep.push_back(k);
        }
        colKeys.swap(keep);
    }
    void noteKeys() { // colMapInsert creates the ColumnMap key of every row's sequence
        for (const Row &d : rows) {
            auto it = std::lower_bound(colKeys.begin(), colKeys.end(), d.key);
            if (it == colKeys.end() || *it != d.key) colKeys.insert(it, d.key);
        }
    }
};

GpuMafExport::GpuMafExport(halgpu_ctx *ctx) : _impl(new Impl), _ctx(ctx) {
    Impl &m = *_impl;
    m.ctx = ctx;
    m.numGenomes = halgpu_num_genomes(ctx);
    m.rank.resize(m.numGenomes);
    m.seqs.resize(m.numGenomes);
    m.nseqs.resize(m.numGenomes);
    m.dna.resize(m.numGenomes);
    std::vector<int> order(m.numGenomes);
    for (int g = 0; g < m.numGenomes; ++g) {
        order[g] = g;
        m.gname.push_back(halgpu_genome_name(ctx, g));
        halgpu_sequence_table(ctx, g, &m.seqs[g], &m.nseqs[g]);
        m.dna[g] = halgpu_genome_dna(ctx, g);
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return m.gname[a] < m.gname[b]; });
    m.genomeOfRank = order;
    for (int i = 0; i < m.numGenomes; ++i) m.rank[order[i]] = i;
}

GpuMafExport::~GpuMafExport() {}

void GpuMafExport::convertSequence(std::ostream &mafStream, int refGenome, int refSequence, int64_t startPosition, uint64_t length,
                                   const std::vector<int> &targets) {
    Impl &m = *_impl;
    if (refGenome < 0 || refGenome >= m.numGenomes || refSequence < 0 || (size_t)refSequence >= m.nseqs[refGenome]) {
        throw std::runtime_error("reference sequence out of range");
    }
    const halgpu_seq &S = m.seqs[refGenome][refSequence];
    if (startPosition >= S.length || (uint64_t)startPosition + length > (uint64_t)S.length) {
        throw std::runtime_error("Invalid range specified for convertGenome");
    }
    if (length == 0) length = (uint64_t)(S.length - startPosition);
    if (length == 0) throw std::runtime_error("Cannot convert zero length sequence");
    if (m.fullNames != _ucscNames) m.names.clear();
    m.fullNames = _ucscNames;
    m.maxLength = _maxLength;
    if (!_append && mafStream.tellp() <= std::streampos(0)) { // MafExport::writeHeader (maf/impl/halMafExport.cpp:15-23)
        mafStream << "##maf version=1 scoring=N/A\n" << "# hal " << halgpu_newick(_ctx) << std::endl << std::endl;
    }
    const uint32_t flags = (_noDupes ? (uint32_t)HALGPU_COL_NO_DUPES : 0u) | (_noAncestors ? (uint32_t)HALGPU_NO_ANCESTORS : 0u) |
                           (_onlyOrthologs ? (uint32_t)HALGPU_ONLY_ORTHOLOGS : 0u) | (_unique ? (uint32_t)HALGPU_COL_UNIQUE : 0u);
    const int64_t sweepFirst = S.start + startPosition; // the ColumnIterator of this call starts here (its visit cache spans all chunks)
    const Key refKey{m.rank[refGenome], refSequence};
    m.colKeys.clear(); // a fresh ColumnIterator per call
    uint64_t appendCount = 0;
    size_t numBlocks = 0;
    std::string out;
    auto flush = [&]() {
        if (appendCount > 0 && (_keepEmptyRefBlocks || !m.referenceIsAllGaps())) {
            m.printBlock(out);
            out += '\n';
            ++blocks;
        }
        if (out.size() > (8u << 20)) { mafStream.write(out.data(), (std::streamsize)out.size()); out.clear(); }
    };
    for (uint64_t done = 0; done < length;) {
        const uint64_t chunk = std::min<uint64_t>(chunkColumns, length - done);
        const int64_t gFirst = S.start + startPosition + (int64_t)done;
        halgpu_col_runs *cr = nullptr;
        char *err = nullptr;
        auto t0 = std::chrono::steady_clock::now();
        if (halgpu_column_runs_in_sweep(_ctx, refGenome, gFirst, gFirst + (int64_t)chunk - 1, sweepFirst, targets.empty() ? nullptr : targets.data(),
                                        targets.size(), flags, &cr, &err) != 0) {
            std::string msg = err ? err : "halgpu_column_runs failed";
            halgpu_free_string(err);
            throw std::runtime_error(msg);
        }
        gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (size_t r = 0; r < cr->n_runs; ++r) {
            const int64_t col0 = cr->run_col[r], runLen = cr->run_col[r + 1] - col0;
            const int cls = cr->run_class ? cr->run_class[r] : 0;
            if (cls == 2) continue; // --unique: positions the iterator skips without walking them (nextFreeIndex)
            const halgpu_col_row *rr = cr->rows + cr->row_offset[r];
            const size_t nr = (size_t)(cr->row_offset[r + 1] - cr->row_offset[r]);
            m.rows.resize(nr);
            for (size_t i = 0; i < nr; ++i) {
                m.rows[i].key = Key{m.rank[rr[i].genome], rr[i].seq};
                m.rows[i].genome = rr[i].genome;
                m.rows[i].pos = rr[i].pos;
                m.rows[i].rev = rr[i].rev != 0;
            }
            m.noteKeys();
            if (cls == 1) continue; // --unique: walked (its sequences are ColumnMap keys now) but isCanonicalOnRef() is false
            int64_t j = 0;
            while (j < runLen) {
                // exact per-column step (MafExport::convertSequence loop body, maf/impl/halMafExport.cpp:48-81)
                const int64_t seqPos = startPosition + (int64_t)done + col0 + j;
                if (appendCount == 0) {
                    m.initBlock(refKey, seqPos);
                } else if (!m.canAppend()) {
                    if (numBlocks++ % 1000 == 0) m.defragment();
                    flush();
                    m.initBlock(refKey, seqPos);
                }
                m.computePairing();
                // this column plus as many of the run's following columns as maxBlockLen allows
                int64_t take = std::min<int64_t>(runLen - j, std::max<int64_t>(1, m.capacity()));
                m.appendColumns(take);
                appendCount += (uint64_t)take;
                j += take;
                for (Row &d : m.rows) d.pos += d.rev ? -take : take;
            }
        }
        columns += cr->n_cols;
        runs += cr->n_runs;
        halgpu_free_col_runs(cr);
        done += chunk;
    }
    if (appendCount > 0 && (_keepEmptyRefBlocks || !m.referenceIsAllGaps())) {
        m.printBlock(out);
        ++blocks;
        mafStream.write(out.data(), (std::streamsize)out.size());
        mafStream << std::endl;
    } else {
        mafStream.write(out.data(), (std::streamsize)out.size());
    }
}

} // namespace halgpu
