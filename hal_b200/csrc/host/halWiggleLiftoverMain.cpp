// halWiggleLiftover -- GPU build of the reference CLI (liftover/impl/halWiggleLiftoverMain.cpp): same positional
// arguments, same options (--noDupes, --append), same messages and exit codes; `--device` selects the GPU.  Storage
// options of the reference's CLParser are accepted and ignored: the input must be a HAL-MMAP file.
#include "wiggle_liftover.hpp"
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace std;

static void usage(ostream &os, const char *prog) {
    os << prog << " v-b200: Map wiggle genome annotation between two genomes on a B200 GPU.\n\n"
       << "USAGE:\n" << prog << " [Options] <halFile> <srcGenome> <srcWig> <tgtGenome> <tgtWig>\n\n"
       << "ARGUMENTS:\nhalFile:     input hal file (mmap format)\nsrcGenome:   source genome name\n"
       << "srcWig:      path of input wig file.  set as stdin to stream from standard input\n"
       << "tgtGenome:   target genome name\ntgtWig:      path of output .wig file.  set as stdout to stream to standard output.\n\n"
       << "OPTIONS:\n--append:             append/merge results into tgtWig.  Note that the entire tgtWig file will be loaded into memory"
          " then overwritten, so this data can be lost in event of a crash [default = 0]\n"
       << "--device <value>:     CUDA device index [default = 0]\n--help:               display this help page [default = 0]\n"
       << "--noDupes:            do not map between duplications in graph. [default = 0]\n";
}

int main(int argc, char **argv) {
    vector<string> pos;
    map<string, string> opt;
    map<string, bool> flag = {{"noDupes", false}, {"append", false}, {"help", false}, {"inMemory", false}, {"udcVerbose", false}};
    const vector<string> valued = {"device", "format", "cacheMDC", "cacheRDC", "cacheBytes", "cacheW0", "chunk", "deflate", "mmapFileSize",
                                   "mmapSizeIncrease", "udcCacheDir"};
    try {
        for (int i = 1; i < argc; ++i) {
            string a = argv[i];
            if (a.rfind("--", 0) == 0) {
                string name = a.substr(2);
                if (flag.count(name)) {
                    flag[name] = true;
                } else if (find(valued.begin(), valued.end(), name) != valued.end()) {
                    if (i + 1 >= argc) throw runtime_error("Option " + a + " requires a value");
                    opt[name] = argv[++i];
                } else {
                    throw runtime_error("Unrecognized option: " + a);
                }
            } else {
                pos.push_back(a);
            }
        }
        if (flag["help"]) { usage(cerr, argv[0]); return 1; }
        if (pos.size() != 5) throw runtime_error(pos.size() < 5 ? "Too few (required positional) arguments" : "Too many (required positional) arguments");
    } catch (exception &e) {
        cerr << e.what() << endl;
        usage(cerr, argv[0]);
        return 1;
    }
    halgpu_ctx *ctx = nullptr;
    int rc = 0;
    try {
        char *err = nullptr;
        if (halgpu_open(pos[0].c_str(), opt.count("device") ? atoi(opt["device"].c_str()) : 0, &ctx, &err) != 0) {
            string m = err ? err : "cannot open";
            halgpu_free_string(err);
            throw runtime_error(m);
        }
        if (halgpu_num_genomes(ctx) == 0) throw runtime_error("hal alignmnet is empty");
        const int src = halgpu_genome_id(ctx, pos[1].c_str());
        if (src < 0) throw runtime_error(string("srcGenome, ") + pos[1] + ", not found in alignment");
        const int tgt = halgpu_genome_id(ctx, pos[3].c_str());
        if (tgt < 0) throw runtime_error(string("tgtGenome, ") + pos[3] + ", not found in alignment");
        ifstream srcWig;
        istream *in = &cin;
        if (pos[2] != "stdin") {
            srcWig.open(pos[2].c_str());
            in = &srcWig;
            if (!srcWig) throw runtime_error("Error opening srcWig, " + pos[2]);
        }
        halgpu::GpuWiggleLiftover lift(ctx);
        if (flag["append"] && pos[4] != "stdout") {
            // load the wig data into memory so that it can be properly merged with the new data from the liftover
            ifstream old(pos[4].c_str());
            if (old) lift.preloadOutput(tgt, &old);
        }
        ofstream tgtWig;
        ostream *out = &cout;
        if (pos[4] != "stdout") {
            tgtWig.open(pos[4].c_str());
            out = &tgtWig;
            if (!tgtWig) throw runtime_error("Error opening tgtWig, " + pos[4]);
        }
        lift.convert(src, in, tgt, out, !flag["noDupes"], false);
        out->flush();
        if (getenv("HALGPU_TIMING")) {
            cerr << "[halWiggleLiftover] lines in " << lift.linesIn << ", source bases " << lift.basesIn << " in " << lift.runs << " runs, target bases out "
                 << lift.basesOut << "; parse " << lift.parseSeconds << " s, halgpu_wiggle_liftover " << lift.gpuSeconds << " s (mapping kernel "
                 << lift.kernelMs << " ms), write " << lift.writeSeconds << " s; " << (lift.fastParsed ? "multi-threaded" : "serial") << " scanner, "
                 << lift.textThreads << " text threads" << endl;
        }
    } catch (exception &e) {
        cerr << "hal exception caught: " << e.what() << endl;
        rc = 1;
    }
    halgpu_close(ctx);
    return rc;
}
