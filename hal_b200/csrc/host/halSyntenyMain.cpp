// halSynteny -- GPU build of the reference CLI (synteny/impl/halSynteny.cpp): same arguments and options
// (--queryGenome, --targetGenome, --queryChromosome, --minBlockSize, --maxAnchorDistance, --alignmentIsPsl), same output;
// `--device` selects the GPU.  The alignment must be a HAL-MMAP file (or a PSL file with --alignmentIsPsl, which needs no GPU
// work at all: blocks are read, chained by dag_merge and written back).
#include "synteny.hpp"
#include <chrono>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace std;

static void usage(ostream &os, const char *prog) {
    os << prog << " v-b200: Convert alignments into synteny blocks\n\n"
       << "USAGE:\n" << prog << " [Options] <alignment> <outPslPath>\n\n"
       << "ARGUMENTS:\nalignment:   input file in HAL (mmap) or PSL format (PSL must specify --alignmentIsPsl)\n"
       << "outPslPath:  output psl file ffor synteny blocks\n\n"
       << "OPTIONS:\n--alignmentIsPsl:           alignment is in PSL format [default = 0]\n"
       << "--device <value>:           CUDA device index [default = 0]\n"
       << "--maxAnchorDistance <value>: upper bound on distance for syntenic psl blocks [default = 5000]\n"
       << "--minBlockSize <value>:     lower bound on synteny block length [default = 5000]\n"
       << "--queryChromosome <value>:  chromosome to infer synteny (default is whole genome) [default = \"\"]\n"
       << "--queryGenome <value>:      source genome [default = \"\"]\n--targetGenome <value>:     reference genome name [default = \"\"]\n";
}

int main(int argc, char **argv) {
    vector<string> pos;
    map<string, string> opt = {{"queryGenome", "\"\""}, {"targetGenome", "\"\""}, {"queryChromosome", "\"\""}, {"minBlockSize", "5000"},
                               {"maxAnchorDistance", "5000"}, {"device", "0"}};
    bool alignmentIsPsl = false;
    const vector<string> ignoredValued = {"format", "cacheMDC", "cacheRDC", "cacheBytes", "cacheW0", "chunk", "deflate", "mmapFileSize",
                                          "mmapSizeIncrease", "udcCacheDir"};
    try {
        for (int i = 1; i < argc; ++i) {
            string a = argv[i];
            if (a.rfind("--", 0) == 0) {
                string name = a.substr(2);
                if (name == "alignmentIsPsl") alignmentIsPsl = true;
                else if (name == "help") { usage(cerr, argv[0]); return 1; }
                else if (name == "inMemory" || name == "udcVerbose") continue;
                else if (opt.count(name) || find(ignoredValued.begin(), ignoredValued.end(), name) != ignoredValued.end()) {
                    if (i + 1 >= argc) throw runtime_error("Option " + a + " requires a value");
                    opt[name] = argv[++i];
                } else throw runtime_error("Unrecognized option: " + a);
            } else pos.push_back(a);
        }
        if (pos.size() != 2) throw runtime_error(pos.size() < 2 ? "Too few (required positional) arguments" : "Too many (required positional) arguments");
    } catch (exception &e) {
        cerr << e.what() << endl;
        usage(cerr, argv[0]);
        return 1;
    }
    const string qName = opt["queryGenome"], tName = opt["targetGenome"], qChrom = opt["queryChromosome"];
    // validateInputOrThrow (synteny/impl/halSynteny.cpp:45-57; the reference throws these outside its try block)
    if (qName == "\"\"" || tName == "\"\"") { cerr << "--queryGenome and --targetGenome and --queryChromosome must bespecified" << endl; return 1; }
    if (qName == tName) { cerr << "--queryGenome and --targetGenome must bedifferent" << endl; return 1; }
    const uint64_t minBlockSize = strtoull(opt["minBlockSize"].c_str(), nullptr, 10), maxAnchorDistance = strtoull(opt["maxAnchorDistance"].c_str(), nullptr, 10);
    halgpu_ctx *ctx = nullptr;
    int rc = 0;
    try {
        if (alignmentIsPsl) { // syntenyFromPsl
            const vector<halgpu::PslBlock> blocks = halgpu::readPslBlocks(pos[0]);
            ofstream out;
            out.exceptions(ofstream::failbit | ofstream::badbit);
            out.open(pos[1], ofstream::out);
            halgpu::writePsl(halgpu::dagMerge(blocks, minBlockSize, maxAnchorDistance), out);
            out.close();
        } else { // syntenyFromHal: one chromosome at a time
            char *err = nullptr;
            if (halgpu_open(pos[0].c_str(), atoi(opt["device"].c_str()), &ctx, &err) != 0) {
                string m = err ? err : "cannot open";
                halgpu_free_string(err);
                throw runtime_error(m);
            }
            if (halgpu_num_genomes(ctx) == 0) throw runtime_error("hal alignment is empty");
            const int tgt = halgpu_genome_id(ctx, tName.c_str());
            if (tgt < 0) throw runtime_error(string("Reference genome, ") + tName + ", not found in alignment");
            const int src = halgpu_genome_id(ctx, qName.c_str());
            if (src < 0) throw runtime_error(string("Reference genome, ") + qName + ", not found in alignment");
            vector<string> chroms;
            if (qChrom != "\"\"") {
                chroms.push_back(qChrom);
            } else {
                const halgpu_seq *seqs = nullptr;
                size_t n = 0;
                halgpu_sequence_table(ctx, src, &seqs, &n);
                for (size_t i = 0; i < n; ++i) chroms.push_back(seqs[i].name);
                sort(chroms.begin(), chroms.end());
            }
            ofstream out;
            out.exceptions(ofstream::failbit | ofstream::badbit);
            out.open(pos[1], ofstream::out);
            halgpu::GpuHal2Psl h2p(ctx);
            double mergeSeconds = 0;
            for (const string &c : chroms) {
                const vector<halgpu::PslBlock> blocks = h2p.convert2psl(src, tgt, c);
                const auto t0 = chrono::steady_clock::now();
                halgpu::writePsl(halgpu::dagMerge(blocks, minBlockSize, maxAnchorDistance), out);
                mergeSeconds += chrono::duration<double>(chrono::steady_clock::now() - t0).count();
            }
            out.close();
            if (getenv("HALGPU_TIMING")) {
                cerr << "[halSynteny] " << h2p.intervals << " GPU intervals, " << h2p.fragments << " mapped fragments, " << h2p.refined
                     << " after refinement, " << h2p.lines << " blocks; halgpu_liftover " << h2p.gpuSeconds << " s, host refine/merge "
                     << h2p.hostSeconds << " s, dag merge + write " << mergeSeconds << " s" << endl;
            }
        }
    } catch (exception &e) {
        cerr << "Exception caught: " << e.what() << endl;
        rc = 1;
    }
    halgpu_close(ctx);
    return rc;
}
