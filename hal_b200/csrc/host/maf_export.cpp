#include "maf_export.hpp"
#include <algorithm>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace halgpu {

namespace {

struct Key { // ColumnIterator::SequenceLess order: genome name bytes, then sequence array index
    int32_t rank, seq;
    bool operator<(const Key &o) const { return rank < o.rank || (rank == o.rank && seq < o.seq); }
    bool operator==(const Key &o) const { return rank == o.rank && seq == o.seq; }
    bool operator!=(const Key &o) const { return !(*this == o); }
};

// The row text of an entry is kept as a list of pieces -- runs of gaps, runs of consecutive bases of one strand -- while the
// (sequential, history dependent) block state machine runs; the characters themselves are decoded from the packed DNA only
// when a finished block is formatted, by many threads at once ("MAF row formatter").
struct Piece {
    int64_t pos;   // genome coordinate of the first base (walks down when rev); unused for gaps
    int64_t count;
    int8_t kind;   // 0 gap, 1 forward bases, 2 reverse-complemented bases
};

struct Entry { // MafBlockEntry (maf/inc/halMafBlock.h:70-115)
    const std::string *name = nullptr; // cached "Genome.Sequence" (or bare sequence name)
    std::vector<Piece> text;
    bool anyBase = false;
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
                e->start = -1; e->strand = '+'; e->length = 0; e->text.clear(); e->anyBase = false;
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
        if (clearText) { e->text.clear(); e->anyBase = false; }
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
    static void decodeBases(char *o, const uint8_t *d, int64_t pos, bool rev, int64_t count) { // DnaIterator::getBase x count
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
                const int8_t kind = d.rev ? 2 : 1;
                if (!e->text.empty() && e->text.back().kind == kind &&
                    d.pos == (d.rev ? e->text.back().pos - e->text.back().count : e->text.back().pos + e->text.back().count)) {
                    e->text.back().count += count;
                } else {
                    e->text.push_back(Piece{d.pos, count, kind});
                }
                e->anyBase = true;
            } else if (!e->text.empty() && e->text.back().kind == 0) {
                e->text.back().count += count;
            } else {
                e->text.push_back(Piece{0, count, 0});
            }
        }
    }
    int64_t capacity() const { // columns that can still be appended before canAppendColumn trips on maxBlockLen
        int64_t cap = INT64_MAX;
        for (auto &pr : pairing)
            if (pr.second >= 0) cap = std::min(cap, maxLength - pr.first->length);
        return cap < 0 ? 0 : cap;
    }
    bool referenceIsAllGaps() const { // MafBlock::referenceIsAllGaps: the reference row's text holds nothing but '-'
        return reference != nullptr && !reference->anyBase;
    }
    // ---- deferred printing: finished blocks are queued as row descriptors and formatted by formatQueued() ----
    struct RowJob {
        const std::string *name;
        int64_t start, length, srcLength;
        size_t firstPiece, numPieces;
        int genome;
        char strand;
    };
    struct BlockJob {
        size_t firstRow, numRows, textBytes;
    };
    std::vector<Piece> jobPieces;
    std::vector<RowJob> jobRows;
    std::vector<BlockJob> jobBlocks;
    size_t queuedBytes = 0;
    unsigned formatThreads = 1;

    void queueEntry(const Entry &e, int64_t start) { // operator<<(MafBlockEntry) (:452-456), text deferred
        RowJob r;
        r.name = e.name; r.start = start; r.length = e.length; r.srcLength = e.srcLength; r.genome = e.genome; r.strand = e.strand;
        r.firstPiece = jobPieces.size(); r.numPieces = e.text.size();
        jobPieces.insert(jobPieces.end(), e.text.begin(), e.text.end());
        size_t len = 0;
        for (const Piece &p : e.text) len += (size_t)p.count;
        jobBlocks.back().textBytes += len + e.name->size() + 64;
        jobRows.push_back(r);
        ++jobBlocks.back().numRows;
    }
    void queueBlock() { // MafBlock::printBlock (:499-519)
        jobBlocks.push_back(BlockJob{jobRows.size(), 0, 2});
        if (reference->start == -1) {
            if (refIndex != -1) queueEntry(*reference, refIndex);
        } else {
            queueEntry(*reference, reference->start);
        }
        for (const Slot &s : entries) {
            if (s.e->start != -1 && s.e.get() != reference) queueEntry(*s.e, s.e->start);
        }
        queuedBytes += jobBlocks.back().textBytes;
    }
    void formatRow(std::string &out, const RowJob &r) const {
        char buf[24];
        out += "s\t"; out += *r.name; out += '\t';
        auto c = std::to_chars(buf, buf + sizeof buf, r.start); out.append(buf, c.ptr); out += '\t';
        c = std::to_chars(buf, buf + sizeof buf, r.length); out.append(buf, c.ptr); out += '\t';
        out += r.strand; out += '\t';
        c = std::to_chars(buf, buf + sizeof buf, r.srcLength); out.append(buf, c.ptr); out += '\t';
        for (size_t i = 0; i < r.numPieces; ++i) {
            const Piece &p = jobPieces[r.firstPiece + i];
            const size_t at = out.size();
            if (p.kind == 0) {
                out.append((size_t)p.count, '-');
            } else {
                out.resize(at + (size_t)p.count);
                decodeBases(&out[at], dna[r.genome], p.pos, p.kind == 2, p.count);
            }
        }
        out += '\n';
    }
    // Writes every queued block ("a\n" + rows + blank line) except that the LAST one gets no blank line when `last` is set
    // (MafExport::convertSequence ends with `mafStream << _mafBlock << endl`).
    // The same bytes produced on the device (halgpu_maf_text, csrc/maf_kernels.cuh): the host formats only the row prefixes
    // and says where every row goes; bases and gap runs are written by the GPU straight from the staged packed DNA.
    bool deviceText = true;
    double deviceTextSeconds = 0, prefixSeconds = 0, writeSeconds = 0, hostTextSeconds = 0;
    // finished text leaves through one writer thread, so the next batch of blocks is assembled while this one is written
    struct WriteJob { char *buf; size_t n; };
    std::thread writer;
    std::mutex wm;
    std::condition_variable wcv;
    std::deque<WriteJob> wq;
    bool wstop = false, wbusy = false;
    std::string werror;
    std::ostream *wos = nullptr;
    void writerLoop() {
        while (true) {
            WriteJob j;
            {
                std::unique_lock<std::mutex> g(wm);
                wcv.wait(g, [&] { return wstop || !wq.empty(); });
                if (wq.empty()) return;
                j = wq.front(); wq.pop_front();
                wbusy = true;
            }
            const auto t0 = std::chrono::steady_clock::now();
            wos->write(j.buf, (std::streamsize)j.n);
            halgpu_host_free(j.buf);
            {
                std::lock_guard<std::mutex> g(wm);
                writeSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                wbusy = false;
            }
            wcv.notify_all();
        }
    }
    void enqueueWrite(std::ostream &os, char *buf, size_t n) {
        {
            std::unique_lock<std::mutex> g(wm);
            wcv.wait(g, [&] { return wq.size() < 2; }); // at most two finished buffers wait for the disk
            wos = &os;
            wq.push_back(WriteJob{buf, n});
        }
        if (!writer.joinable()) writer = std::thread([this] { writerLoop(); });
        wcv.notify_all();
    }
    void drainWriter() { // everything handed to the writer is in the stream
        std::unique_lock<std::mutex> g(wm);
        wcv.wait(g, [&] { return wq.empty() && !wbusy; });
    }
    ~Impl() {
        {
            std::lock_guard<std::mutex> g(wm);
            wstop = true;
        }
        wcv.notify_all();
        if (writer.joinable()) writer.join();
    }
    void formatQueuedOnDevice(std::ostream &os, bool last) {
        const size_t nb = jobBlocks.size();
        auto t0 = std::chrono::steady_clock::now();
        // rows in output order; (block, row) -> descriptor.  The prefixes are formatted by several threads, each into its own
        // string; offsets follow from two prefix sums.
        struct RowRef { size_t row; bool first, lastOfBlock, blank; };
        std::vector<RowRef> refs;
        refs.reserve(jobRows.size());
        for (size_t b = 0; b < nb; ++b) {
            const BlockJob &B = jobBlocks[b];
            const bool blank = !(last && b + 1 == nb);
            for (size_t r = B.firstRow; r < B.firstRow + B.numRows; ++r) refs.push_back(RowRef{r, r == B.firstRow, r + 1 == B.firstRow + B.numRows, blank});
        }
        const size_t nr = refs.size();
        std::vector<halgpu_maf_row> rows(nr);
        std::vector<uint64_t> textLen(nr);
        const unsigned T = std::max(1u, std::min<unsigned>(formatThreads, (unsigned)(nr >> 12) + 1));
        std::vector<std::string> parts(T);
        auto work = [&](unsigned k) {
            const size_t lo = nr * k / T, hi = nr * (k + 1) / T;
            std::string &px = parts[k];
            px.reserve((hi - lo) * 48);
            char buf[24];
            for (size_t i = lo; i < hi; ++i) {
                const RowJob &R = jobRows[refs[i].row];
                halgpu_maf_row &w = rows[i];
                const size_t at0 = px.size();
                if (refs[i].first) px += "a\n";
                px += "s\t"; px += *R.name; px += '\t';
                auto c = std::to_chars(buf, buf + sizeof buf, R.start); px.append(buf, c.ptr); px += '\t';
                c = std::to_chars(buf, buf + sizeof buf, R.length); px.append(buf, c.ptr); px += '\t';
                px += R.strand; px += '\t';
                c = std::to_chars(buf, buf + sizeof buf, R.srcLength); px.append(buf, c.ptr); px += '\t';
                w.prefix_offset = (uint32_t)at0; // relative to this thread's string for now
                w.prefix_len = (uint32_t)(px.size() - at0);
                w.first_piece = (uint32_t)R.firstPiece; w.num_pieces = (uint32_t)R.numPieces; w.genome = R.genome;
                w.tail_newlines = (refs[i].lastOfBlock && refs[i].blank) ? 2u : 1u;
                uint64_t text = 0;
                for (size_t q = 0; q < R.numPieces; ++q) text += (uint64_t)jobPieces[R.firstPiece + q].count;
                textLen[i] = text;
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (unsigned k = 0; k < T; ++k) th.emplace_back(work, k);
            for (auto &x : th) x.join();
        }
        std::string prefix;
        size_t total = 0;
        for (const std::string &px : parts) total += px.size();
        prefix.reserve(total);
        uint64_t at = 0;
        for (unsigned k = 0; k < T; ++k) {
            const size_t lo = nr * k / T, hi = nr * (k + 1) / T, base = prefix.size();
            prefix += parts[k];
            for (size_t i = lo; i < hi; ++i) {
                rows[i].prefix_offset += (uint32_t)base;
                rows[i].out_offset = at;
                at += rows[i].prefix_len + textLen[i] + rows[i].tail_newlines;
            }
        }
        std::vector<halgpu_maf_piece> pieces(jobPieces.size());
        for (size_t i = 0; i < jobPieces.size(); ++i) {
            pieces[i].pos = jobPieces[i].pos;
            pieces[i].count_kind = (jobPieces[i].count << 2) | (int64_t)jobPieces[i].kind;
        }
        if (prefix.size() >= 0xffffffffull || jobPieces.size() >= 0xffffffffull) throw std::runtime_error("too much queued MAF text for one device call");
        prefixSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        t0 = std::chrono::steady_clock::now();
        char *out = static_cast<char *>(halgpu_host_alloc(std::max<uint64_t>(at, 1)));
        if (out == nullptr) throw std::runtime_error("out of page-locked memory for the MAF text");
        char *err = nullptr;
        const int rc = halgpu_maf_text(ctx, rows.size(), rows.data(), pieces.size(), pieces.data(), prefix.data(), prefix.size(), at, out, nullptr, &err);
        deviceTextSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc != 0) {
            std::string msg = err ? err : "halgpu_maf_text failed";
            halgpu_free_string(err);
            halgpu_host_free(out);
            throw std::runtime_error(msg);
        }
        enqueueWrite(os, out, (size_t)at);
        jobPieces.clear(); jobRows.clear(); jobBlocks.clear();
        queuedBytes = 0;
    }
    void formatQueued(std::ostream &os, bool last) {
        const size_t nb = jobBlocks.size();
        if (nb == 0) return;
        if (deviceText) { formatQueuedOnDevice(os, last); return; }
        drainWriter();
        const auto tHost0 = std::chrono::steady_clock::now();
        unsigned T = std::max(1u, std::min<unsigned>(formatThreads, (unsigned)(queuedBytes >> 20) + 1));
        std::vector<size_t> cut(T + 1, nb); // contiguous groups of blocks with about equal text
        cut[0] = 0;
        size_t acc = 0, t = 1;
        for (size_t b = 0; b < nb && t < T; ++b) {
            acc += jobBlocks[b].textBytes;
            if (acc >= queuedBytes * t / T) cut[t++] = b + 1;
        }
        std::vector<std::string> outs(T);
        auto work = [&](unsigned k) {
            std::string &out = outs[k];
            size_t bytes = 0;
            for (size_t b = cut[k]; b < cut[k + 1]; ++b) bytes += jobBlocks[b].textBytes;
            out.reserve(bytes);
            for (size_t b = cut[k]; b < cut[k + 1]; ++b) {
                out += "a\n";
                const BlockJob &B = jobBlocks[b];
                for (size_t r = B.firstRow; r < B.firstRow + B.numRows; ++r) formatRow(out, jobRows[r]);
                if (!(last && b + 1 == nb)) out += '\n';
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (unsigned k = 0; k < T; ++k) th.emplace_back(work, k);
            for (auto &x : th) x.join();
        }
        hostTextSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tHost0).count();
        const auto tw0 = std::chrono::steady_clock::now();
        for (const std::string &o : outs) os.write(o.data(), (std::streamsize)o.size());
        writeSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
        jobPieces.clear(); jobRows.clear(); jobBlocks.clear();
        queuedBytes = 0;
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
double GpuMafExport::textSeconds() const { return _impl->deviceTextSeconds + _impl->prefixSeconds + _impl->hostTextSeconds; }
double GpuMafExport::writeSeconds() const { return _impl->writeSeconds; }

unsigned GpuMafExport::defaultFormatThreads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hw ? hw : 1u, 32u));
}

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
    m.formatThreads = formatThreads;
    m.deviceText = std::getenv("HALGPU_MAF_HOST_TEXT") == nullptr; // measurement / test switch: format on host threads instead
    auto flush = [&]() {
        if (appendCount > 0 && (_keepEmptyRefBlocks || !m.referenceIsAllGaps())) {
            m.queueBlock();
            ++blocks;
        }
        if (m.queuedBytes > queueBytes) m.formatQueued(mafStream, false);
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
        const auto tb0 = std::chrono::steady_clock::now();
        const double textBefore = textSeconds();
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
        blockerSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count() - (textSeconds() - textBefore);
        columns += cr->n_cols;
        runs += cr->n_runs;
        halgpu_free_col_runs(cr);
        done += chunk;
    }
    if (appendCount > 0 && (_keepEmptyRefBlocks || !m.referenceIsAllGaps())) {
        m.queueBlock();
        ++blocks;
        m.formatQueued(mafStream, true);
        m.drainWriter();
        mafStream << std::endl;
    } else {
        m.formatQueued(mafStream, false);
        m.drainWriter();
    }
}

} // namespace halgpu
