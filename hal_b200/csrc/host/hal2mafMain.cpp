// hal2maf -- GPU build of the reference CLI (maf/impl/hal2maf.cpp): same arguments, options, MAF text and
// messages for the ColumnIterator flags the GPU column walk implements (maxRefGap=0; --unique included).
// --refTargets <bed|stdin> drives one convertSequence per BED interval / BED12 block like MafBed (maf/impl/halMafBed.cpp:24-54).
// Not implemented (rejected with an error): --maxRefGap > 0, --global, --printTree.
#include <unistd.h>
#include "bed.hpp"
#include "maf_export.hpp"
#include <chrono>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <cctype>
#include <cstring>
#include <iostream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace std;

static void usage(ostream &os, const char *prog) {
    os << prog << " v-b200: Convert hal database to maf on a B200 GPU.\n\nUSAGE:\n" << prog << " [Options] <halFile> <mafFile>\n\n"
       << "OPTIONS:\n--refGenome <name>, --refSequence <name>, --refTargets <bed|stdin>, --start <n>, --length <n>, --rootGenome <name>, --targetGenomes <a,b,..>,\n"
       << "--noDupes, --noAncestors, --onlySequenceNames, --onlyOrthologs, --unique, --keepEmptyRefBlocks, --append, --maxBlockLen <n>,\n"
       << "--device <n>, --help\n";
}

static void subTree(halgpu_ctx *ctx, int g, vector<int> &out) { // getGenomesInSubTree (api/impl/halCommon.cpp:189-195)
    out.push_back(g);
    for (int k = 0; k < halgpu_genome_num_children(ctx, g); ++k) subTree(ctx, halgpu_genome_child(ctx, g, k), out);
}

static string fixString(const string &s) { // maf/impl/hal2maf.cpp:94-101
    if (s == "\"\"") {
        cerr << "WARNING missing string arguments should be specified as empty strings, not the obsolete '\"\"'" << endl;
        return "";
    }
    return s;
}

// hal2mafWithTargets + MafBed::visitLine (maf/impl/hal2maf.cpp:105-119, halMafBed.cpp:24-54) under BedScanner::scan
// (liftover/impl/halBedScanner.cpp:40-61): one convertSequence per BED<=9 line or per BED12 block, with the reference's
// messages.  The reference never tests the stream it opens (it tests a second, unopened one), so an unreadable path is an
// empty scan there and here.
static void refTargets(halgpu::GpuMafExport &ex, ostream &maf, halgpu_ctx *ctx, int ref, const halgpu_seq *seqs, size_t nseq, const string &path,
                       const vector<int> &targets) {
    ifstream bedFile;
    if (path != "stdin") bedFile.open(path);
    istream &bed = path != "stdin" ? static_cast<istream &>(bedFile) : cin;
    if (bed.bad()) throw runtime_error("Error reading bed input stream");
    const string text((istreambuf_iterator<char>(bed)), istreambuf_iterator<char>());
    map<string, int> seqByName;
    for (size_t i = 0; i < nseq; ++i) seqByName[seqs[i].name] = (int)i;
    halgpu::BedLine cur; // one object for the whole scan, like BedScanner::_bedLine
    string lineBuf;
    size_t lineNumber = 0;
    const char *p = text.data(), *const end = p + text.size();
    auto skipWs = [&]() { while (p < end && isspace((unsigned char)*p)) ++p; };
    skipWs();
    while (p < end) {
        ++lineNumber;
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
        lineBuf.assign(p, nl ? nl : end);
        p = nl ? nl + 1 : end;
        try {
            cur.parse(lineBuf, 0);
            const auto it = seqByName.find(cur.chrName);
            if (it == seqByName.end()) {
                cerr << "Line " << lineNumber << ": BED sequence " << cur.chrName << " not found in genome " << halgpu_genome_name(ctx, ref) << '\n';
            } else if (cur.bedType <= 9) {
                if (cur.end <= cur.start || cur.end > seqs[it->second].length) {
                    cerr << "Line " << lineNumber << ": BED coordinates invalid\n";
                } else {
                    ex.convertSequence(maf, ref, it->second, cur.start, (uint64_t)(cur.end - cur.start), targets);
                }
            } else {
                for (size_t i = 0; i < cur.blocks.size(); ++i) {
                    const int64_t bs = cur.start + cur.blocks[i].start, bl = cur.blocks[i].length;
                    if (bl == 0 || bs + bl >= seqs[it->second].length) {
                        cerr << "Line " << lineNumber << ", block " << i << ": BED coordinates invalid\n";
                    } else {
                        ex.convertSequence(maf, ref, it->second, bs, (uint64_t)bl, targets);
                    }
                }
            }
        } catch (exception &e) {
            throw runtime_error(string(e.what()) + " in input bed line " + to_string(lineNumber));
        }
        skipWs();
    }
}

int main(int argc, char **argv) {
    const auto tMain = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    double openSeconds = 0;
    string halPath, mafPath, refGenomeName, rootGenomeName, targetGenomes, refSequenceName, refTargetsPath;
    int64_t start = 0, maxBlockLen = 1000;
    uint64_t length = 0;
    bool noDupes = false, noAncestors = false, onlySequenceNames = false, append = false, onlyOrthologs = false, keepEmptyRefBlocks = false,
         unique = false;
    int device = 0;
    vector<string> pos;
    try {
        for (int i = 1; i < argc; ++i) {
            string a = argv[i];
            auto val = [&]() -> string { if (i + 1 >= argc) throw runtime_error("Option " + a + " requires a value"); return argv[++i]; };
            if (a == "--refGenome") refGenomeName = fixString(val());
            else if (a == "--refSequence") refSequenceName = fixString(val());
            else if (a == "--start") start = strtoll(val().c_str(), nullptr, 10);
            else if (a == "--length") length = strtoull(val().c_str(), nullptr, 10);
            else if (a == "--rootGenome") rootGenomeName = fixString(val());
            else if (a == "--targetGenomes") targetGenomes = fixString(val());
            else if (a == "--maxBlockLen") maxBlockLen = strtoll(val().c_str(), nullptr, 10);
            else if (a == "--device") device = atoi(val().c_str());
            else if (a == "--noDupes") noDupes = true;
            else if (a == "--noAncestors") noAncestors = true;
            else if (a == "--onlySequenceNames") onlySequenceNames = true;
            else if (a == "--append") append = true;
            else if (a == "--onlyOrthologs") onlyOrthologs = true;
            else if (a == "--keepEmptyRefBlocks") keepEmptyRefBlocks = true;
            else if (a == "--help") { usage(cerr, argv[0]); return 1; }
            else if (a == "--maxRefGap") { if (atoll(val().c_str()) != 0) throw runtime_error("--maxRefGap > 0 is not implemented in the GPU build"); }
            else if (a == "--unique") unique = true;
            else if (a == "--global" || a == "--printTree") throw runtime_error(a + " is not implemented in the GPU build");
            else if (a == "--refTargets") refTargetsPath = fixString(val());
            else if (a == "--format" || a == "--cacheMDC" || a == "--cacheRDC" || a == "--cacheBytes" || a == "--cacheW0" ||
                     a == "--mmapFileSize" || a == "--mmapSizeIncrease" || a == "--udcCacheDir") val();
            else if (a == "--inMemory") {}
            else if (a.rfind("--", 0) == 0) throw runtime_error("Unrecognized option: " + a);
            else pos.push_back(a);
        }
        if (pos.size() != 2) throw runtime_error(pos.size() < 2 ? "Too few (required positional) arguments" : "Too many (required positional) arguments");
        halPath = pos[0];
        mafPath = pos[1];
        if (rootGenomeName != "" && targetGenomes != "") throw runtime_error("--rootGenome and --targetGenomes options are mutually exclusive");
        if (refSequenceName == "" && (start != 0 || length != 0)) throw runtime_error("--start and --length require --refSequenceName");
        if (!refTargetsPath.empty() && (start != 0 || length != 0 || !refSequenceName.empty())) {
            throw runtime_error("--refSequence, --start, and --length options are unsupported when using BED input");
        }
    } catch (exception &e) {
        cerr << e.what() << endl;
        usage(cerr, argv[0]);
        return 1;
    }
    halgpu_ctx *ctx = nullptr;
    int rc = 0;
    try {
        char *err = nullptr;
        const auto tOpen = std::chrono::steady_clock::now();
        const int openRc = halgpu_open(halPath.c_str(), device, &ctx, &err);
        openSeconds = since(tOpen);
        if (openRc != 0) {
            string m = err ? err : "cannot open";
            halgpu_free_string(err);
            throw runtime_error(m);
        }
        int rootId = -1;
        for (int g = 0; g < halgpu_num_genomes(ctx); ++g) if (halgpu_genome_parent(ctx, g) < 0) rootId = g;
        vector<int> targets;
        if (rootGenomeName != "") {
            const int r = halgpu_genome_id(ctx, rootGenomeName.c_str());
            if (r < 0) throw runtime_error("Root genome " + rootGenomeName + ", not found in alignment");
            if (r != rootId) subTree(ctx, r, targets);
        }
        if (targetGenomes != "") {
            size_t b = 0;
            while (b <= targetGenomes.size()) {
                size_t e = targetGenomes.find(',', b);
                if (e == string::npos) e = targetGenomes.size();
                if (e > b) {
                    const string name = targetGenomes.substr(b, e - b);
                    const int t = halgpu_genome_id(ctx, name.c_str());
                    if (t < 0) throw runtime_error("Target genome, " + name + ", not found in alignment");
                    targets.push_back(t);
                }
                b = e + 1;
            }
        }
        int ref = rootId;
        if (refGenomeName != "") {
            ref = halgpu_genome_id(ctx, refGenomeName.c_str());
            if (ref < 0) throw runtime_error("Reference genome, " + refGenomeName + ", not found in alignment");
        }
        if (noAncestors && halgpu_genome_num_children(ctx, ref) != 0) {
            throw runtime_error(string("Since the reference genome to be used for the MAF is ancestral (") + halgpu_genome_name(ctx, ref) +
                                "), the --noAncestors option is invalid.  The --refGenome option can be used to specify a different reference.");
        }
        const halgpu_seq *seqs = nullptr;
        size_t nseq = 0;
        halgpu_sequence_table(ctx, ref, &seqs, &nseq);
        int refSeq = -1;
        if (refSequenceName != "") {
            for (size_t i = 0; i < nseq; ++i) if (refSequenceName == seqs[i].name) refSeq = (int)i;
            if (refSeq < 0) {
                throw runtime_error("Reference sequence, " + refSequenceName + ", not found in reference genome, " + halgpu_genome_name(ctx, ref));
            }
        }
        ofstream mafFile;
        if (mafPath != "stdout") {
            mafFile.open(mafPath, append ? ios_base::out | ios_base::app : ios_base::out);
            if (!mafFile) throw runtime_error("Error opening " + mafPath);
        }
        ostream &maf = mafPath != "stdout" ? mafFile : cout;
        halgpu::GpuMafExport ex(ctx);
        ex.setNoDupes(noDupes); ex.setNoAncestors(noAncestors); ex.setUcscNames(!onlySequenceNames); ex.setAppend(append);
        ex.setMaxBlockLength(maxBlockLen); ex.setOnlyOrthologs(onlyOrthologs); ex.setKeepEmptyRefBlocks(keepEmptyRefBlocks); ex.setUnique(unique);
        if (const char *cc = getenv("HALGPU_MAF_CHUNK_COLUMNS")) ex.chunkColumns = (size_t)std::max(1L, atol(cc)); // test hook
        if (const char *cc = getenv("HALGPU_MAF_QUEUE_BYTES")) ex.queueBytes = (size_t)std::max(0L, atol(cc));      // test hook
        if (const char *cc = getenv("HALGPU_TEXT_THREADS")) ex.formatThreads = (unsigned)std::max(1L, atol(cc));
        if (!refTargetsPath.empty()) {
            refTargets(ex, maf, ctx, ref, seqs, nseq, refTargetsPath, targets);
        } else if (refSeq >= 0) {
            ex.convertSequence(maf, ref, refSeq, start, length, targets);
        } else {
            for (size_t i = 0; i < nseq; ++i) ex.convertSequence(maf, ref, (int)i, start, length, targets);
        }
        maf.flush();
        if (getenv("HALGPU_TIMING") != nullptr) {
            cerr << "[hal2maf] columns " << ex.columns << ", runs " << ex.runs << ", blocks " << ex.blocks << "; column runs (GPU) " << ex.gpuSeconds
                 << " s, block state machine " << ex.blockerSeconds << " s, row text " << ex.textSeconds() << " s, write (own thread) " << ex.writeSeconds()
                 << " s; open (CUDA context) " << openSeconds << " s, total so far " << since(tMain) << " s" << endl;
        }
        if (mafPath != "stdout" && mafFile.tellp() == (streampos)0) std::remove(mafPath.c_str()); // hal2maf.cpp:206-215
    } catch (exception &e) {
        cerr << "hal exception caught: " << e.what() << endl;
        rc = 1;
    }
    // Nothing is left to do but to tear down the CUDA context, the staged genomes and the pinned buffers (0.1 - 1 s of a 2 - 3 s
    // process): every output stream is flushed and closed by now, so the process ends here.  HALGPU_CLEAN_EXIT=1 keeps the
    // orderly teardown (leak checkers, tools/sanitize.sh).
    if (getenv("HALGPU_CLEAN_EXIT") == nullptr) {
        if (getenv("HALGPU_TIMING") != nullptr) cerr << "[hal2maf] main() " << since(tMain) << " s (no teardown)" << endl;
        cout.flush();
        cerr.flush();
        fflush(nullptr);
        _exit(rc);
    }
    const auto tClose = std::chrono::steady_clock::now();
    halgpu_close(ctx);
    if (getenv("HALGPU_TIMING") != nullptr) cerr << "[hal2maf] close " << since(tClose) << " s, main() " << since(tMain) << " s" << endl;
    return rc;
}
