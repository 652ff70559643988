// halAlignmentDepth -- GPU build of the reference CLI (alignmentDepth/halAlignmentDepth.cpp): same arguments,
// options, wiggle text and messages.  Each printSequence() call of the reference (one ColumnIterator sweep,
// :215-308) becomes one halgpu_columns_depth call; text formatting stays on the host.
//
// --step > 1 follows the reference's release-build behaviour: values are printed for start, start+step, ...
// up to and INCLUDING start+length when the step lands on it (the reference passes the one-past-the-end
// position as lastColumnIndex to ColumnIterator::toSite, :305), clamped to the end of the genome.
#include "../../../include/halgpu.h"
#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace std;

namespace {

struct Options {
    string halPath, refGenome, outWiggle = "stdout", refSequence = "\"\"", rootGenome = "\"\"", targetGenomes = "\"\"";
    uint64_t start = 0, length = 0, step = 1;
    bool countDupes = false, noAncestors = false;
    int device = 0;
};

void usage(ostream &os, const char *prog) {
    os << prog << " v-b200: Make alignment depth wiggle plot for a genome on a B200 GPU. By default, this is a count of the number "
       << "of other unique genomes each base aligns to, including ancestral genomes.\n\n"
       << "USAGE:\n" << prog << " [Options] <halPath> <refGenome>\n\nOPTIONS:\n"
       << "--countDupes, --noAncestors, --outWiggle <path>, --refSequence <name>, --start <n>, --length <n>, --step <n>,\n"
       << "--rootGenome <name>, --targetGenomes <a,b,...>, --device <n>, --help\n";
}

void subTree(halgpu_ctx *ctx, int g, vector<int> &out) { // getGenomesInSubTree (api/impl/halCommon.cpp:189-195)
    out.push_back(g);
    for (int k = 0; k < halgpu_genome_num_children(ctx, g); ++k) subTree(ctx, halgpu_genome_child(ctx, g, k), out);
}

struct Printer {
    halgpu_ctx *ctx;
    int ref;
    vector<int> targets;
    uint32_t flags;
    uint64_t step;
    ostream &os;
    string buf;
    vector<int32_t> depth;

    void printSequence(const halgpu_seq &seq, uint64_t start, uint64_t length) { // :215-308
        const uint64_t seqLen = (uint64_t)seq.length;
        if (seqLen == 0) return;
        if (length == 0) length = seqLen - start;
        const uint64_t last = start + length;
        if (last > seqLen) {
            throw runtime_error("Specified range [" + to_string(start) + "," + to_string(length) + "] isout of range for sequence " +
                                seq.name + ", which has length " + to_string(seqLen));
        }
        buf.clear();
        buf += "fixedStep chrom=";
        buf += seq.name;
        buf += " start=" + to_string(start + 1) + " step=" + to_string(step) + "\n";
        const int64_t gFirst = (int64_t)start + seq.start;
        int64_t gLast = (int64_t)last - 1 + seq.start;
        if (step > 1) gLast = min<int64_t>((int64_t)last + seq.start, halgpu_genome_length(ctx, ref) - 1);
        const size_t n = (size_t)((gLast - gFirst) / (int64_t)step + 1);
        depth.resize(n);
        char *err = nullptr;
        if (halgpu_columns_depth(ctx, ref, gFirst, gLast, (int64_t)step, targets.empty() ? nullptr : targets.data(), targets.size(), flags,
                                 depth.data(), nullptr, &err) != 0) {
            string m = err ? err : "halgpu_columns_depth failed";
            halgpu_free_string(err);
            throw runtime_error(m);
        }
        char tmp[16];
        for (size_t i = 0; i < n; ++i) {
            auto r = to_chars(tmp, tmp + sizeof tmp, depth[i]);
            buf.append(tmp, r.ptr);
            buf += '\n';
            if (buf.size() > (16u << 20)) { os.write(buf.data(), (streamsize)buf.size()); buf.clear(); }
        }
        os.write(buf.data(), (streamsize)buf.size());
    }
};

} // namespace

int main(int argc, char **argv) {
    Options o;
    vector<string> pos;
    try {
        for (int i = 1; i < argc; ++i) {
            string a = argv[i];
            auto val = [&]() -> string { if (i + 1 >= argc) throw runtime_error("Option " + a + " requires a value"); return argv[++i]; };
            if (a == "--countDupes") o.countDupes = true;
            else if (a == "--noAncestors") o.noAncestors = true;
            else if (a == "--help") { usage(cerr, argv[0]); return 1; }
            else if (a == "--outWiggle") o.outWiggle = val();
            else if (a == "--refSequence") o.refSequence = val();
            else if (a == "--start") o.start = strtoull(val().c_str(), nullptr, 10);
            else if (a == "--length") o.length = strtoull(val().c_str(), nullptr, 10);
            else if (a == "--step") o.step = strtoull(val().c_str(), nullptr, 10);
            else if (a == "--rootGenome") o.rootGenome = val();
            else if (a == "--targetGenomes") o.targetGenomes = val();
            else if (a == "--device") o.device = atoi(val().c_str());
            else if (a == "--format" || a == "--cacheMDC" || a == "--cacheRDC" || a == "--cacheBytes" || a == "--cacheW0" ||
                     a == "--mmapFileSize" || a == "--mmapSizeIncrease" || a == "--udcCacheDir") val(); // storage options: ignored
            else if (a == "--inMemory") {}
            else if (a.rfind("--", 0) == 0) throw runtime_error("Unrecognized option: " + a);
            else pos.push_back(a);
        }
        if (pos.size() != 2) throw runtime_error(pos.size() < 2 ? "Too few (required positional) arguments" : "Too many (required positional) arguments");
        o.halPath = pos[0];
        o.refGenome = pos[1];
        if (o.rootGenome != "\"\"" && o.targetGenomes != "\"\"") {
            throw runtime_error("--rootGenome and --targetGenomes options are  mutually exclusive");
        }
        if (o.step < 1) throw runtime_error("--step must be at least 1");
    } catch (exception &e) {
        cerr << e.what() << endl;
        usage(cerr, argv[0]);
        return 1;
    }
    halgpu_ctx *ctx = nullptr;
    int rc = 0;
    try {
        char *err = nullptr;
        if (halgpu_open(o.halPath.c_str(), o.device, &ctx, &err) != 0) {
            string m = err ? err : "cannot open";
            halgpu_free_string(err);
            throw runtime_error(m);
        }
        int rootId = -1;
        for (int g = 0; g < halgpu_num_genomes(ctx); ++g) if (halgpu_genome_parent(ctx, g) < 0) rootId = g;
        vector<int> targets;
        if (o.rootGenome != "\"\"") {
            const int r = halgpu_genome_id(ctx, o.rootGenome.c_str());
            if (r < 0) throw runtime_error("Root genome, " + o.rootGenome + ", not found in alignment");
            if (r != rootId) subTree(ctx, r, targets);
        }
        if (o.targetGenomes != "\"\"") {
            size_t b = 0;
            while (b <= o.targetGenomes.size()) {
                size_t e = o.targetGenomes.find(',', b);
                if (e == string::npos) e = o.targetGenomes.size();
                if (e > b) {
                    const string name = o.targetGenomes.substr(b, e - b);
                    const int t = halgpu_genome_id(ctx, name.c_str());
                    if (t < 0) throw runtime_error("Target genome, " + name + ", not found in alignment");
                    targets.push_back(t);
                }
                b = e + 1;
            }
        }
        int ref = rootId;
        if (o.refGenome != "\"\"") {
            ref = halgpu_genome_id(ctx, o.refGenome.c_str());
            if (ref < 0) throw runtime_error("Reference genome, " + o.refGenome + ", not found in alignment");
        }
        const halgpu_seq *seqs = nullptr;
        size_t nseq = 0;
        halgpu_sequence_table(ctx, ref, &seqs, &nseq);
        const halgpu_seq *refSeq = nullptr;
        if (o.refSequence != "\"\"") {
            for (size_t i = 0; i < nseq; ++i) if (o.refSequence == seqs[i].name) refSeq = &seqs[i];
            if (refSeq == nullptr) {
                throw runtime_error("Reference sequence, " + o.refSequence + ", not found in reference genome, " + halgpu_genome_name(ctx, ref));
            }
        }
        if (halgpu_genome_num_children(ctx, ref) != 0 && o.noAncestors) {
            throw runtime_error(string("--noAncestors cannot be used when reference genome (") + halgpu_genome_name(ctx, ref) + ") is ancetral");
        }
        ofstream ofile;
        if (o.outWiggle != "stdout") {
            ofile.open(o.outWiggle.c_str());
            if (!ofile) throw runtime_error("Error opening output file " + o.outWiggle);
        }
        ostream &os = o.outWiggle == "stdout" ? cout : ofile;
        Printer pr{ctx, ref, targets, (o.countDupes ? (uint32_t)HALGPU_COUNT_DUPES : 0u) | (o.noAncestors ? (uint32_t)HALGPU_NO_ANCESTORS : 0u),
                   o.step, os, string(), vector<int32_t>()};
        if (refSeq != nullptr) {
            pr.printSequence(*refSeq, o.start, o.length);
        } else { // printGenome (:318-346)
            const uint64_t glen = (uint64_t)halgpu_genome_length(ctx, ref);
            uint64_t start = o.start, length = o.length;
            if (start + length > glen) {
                throw runtime_error("Specified range [" + to_string(start) + "," + to_string(length) + "] isout of range for genome " +
                                    halgpu_genome_name(ctx, ref) + ", which has length " + to_string(glen));
            }
            if (length == 0) length = glen - start;
            uint64_t running = 0;
            for (size_t i = 0; i < nseq; ++i) {
                const uint64_t seqLen = (uint64_t)seqs[i].length, seqStart = (uint64_t)seqs[i].start;
                if (start + length >= seqStart && start < seqStart + seqLen && running < length) {
                    const uint64_t readStart = seqStart >= start ? 0 : start - seqStart;
                    uint64_t readLen = min(seqLen - readStart, length);
                    readLen = min(readLen, length - running);
                    pr.printSequence(seqs[i], readStart, readLen);
                    running += readLen;
                }
            }
        }
        os.flush();
    } catch (exception &e) {
        cerr << "hal exception caught: " << e.what() << endl;
        rc = 1;
    }
    halgpu_close(ctx);
    return rc;
}
