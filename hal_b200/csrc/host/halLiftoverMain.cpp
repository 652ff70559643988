// halLiftover -- GPU build of the reference CLI (liftover/impl/halLiftoverMain.cpp): same positional arguments,
// same options, same messages and exit codes; `--device` selects the GPU.  Storage options of the reference's
// CLParser (--format, --cacheBytes, ... api/impl/halCLParser.cpp:21-31) are accepted and ignored: the input must
// be a HAL-MMAP file (convert HDF5 files with the reference's halExtract --outputFormat mmap).
#include <unistd.h>
#include "gpu_liftover.hpp"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace std;

static void usage(ostream &os, const char *prog) {
    os << prog << " v-b200: Map BED genome interval coordinates between two genomes on a B200 GPU.\n\n"
       << "USAGE:\n" << prog << " [Options] <halFile> <srcGenome> <srcBed> <tgtGenome> <tgtBed>\n\n"
       << "ARGUMENTS:\nhalFile:     input hal file (mmap format)\nsrcGenome:   source genome name\n"
       << "srcBed:      path of input bed file.  set as stdin to stream from standard input\n"
       << "tgtGenome:   target genome name\ntgtBed:      path of output bed file.  set as stdout to stream to standard output.\n\n"
       << "OPTIONS:\n--append:             append results to tgtBed [default = 0]\n"
       << "--bedType <value>:    number of standard columns (3 to 12), columns beyond this are passed through. [default = 0]\n"
       << "--coalescenceLimit <value>: coalescence limit genome: the genome at or above which paralogs coalesce (default: the MRCA of source and target) [default = \"\"]\n"
       << "--device <value>:     CUDA device index [default = 0]\n--help:               display this help page [default = 0]\n"
       << "--noDupes:            do not map between duplications in graph. [default = 0]\n"
       << "--columnLiftover:     (extension) use hal::ColumnLiftover instead of hal::BlockLiftover [default = 0]\n"
       << "--outPSL:             write output in PSL instead of bed format [default = 0]\n"
       << "--outPSLWithName:     write output as input BED name followed by PSL line instead of bed format [default = 0]\n";
}

int main(int argc, char **argv) {
    vector<string> pos;
    map<string, string> opt;
    map<string, bool> flag = {{"noDupes", false}, {"append", false}, {"outPSL", false}, {"outPSLWithName", false}, {"help", false},
                              {"inMemory", false}, {"udcVerbose", false}, {"columnLiftover", false}};
    const vector<string> valued = {"coalescenceLimit", "bedType", "device", "format", "cacheMDC", "cacheRDC", "cacheBytes", "cacheW0",
                                   "chunk", "deflate", "mmapFileSize", "mmapSizeIncrease", "udcCacheDir"};
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
        if (opt.count("bedType")) {
            int bt = atoi(opt["bedType"].c_str());
            if (bt < 3 || bt > 12) throw runtime_error("--bedType must be between 3 and 12");
        }
    } catch (exception &e) {
        cerr << e.what() << endl;
        usage(cerr, argv[0]);
        return 1;
    }
    halgpu_ctx *ctx = nullptr;
    int rc = 0;
    const auto tMain = chrono::steady_clock::now();
    auto since = [](chrono::steady_clock::time_point t) { return chrono::duration<double>(chrono::steady_clock::now() - t).count(); };
    double openSeconds = 0;
    try {
        const int bedType = opt.count("bedType") ? atoi(opt["bedType"].c_str()) : 0;
        bool outPSL = flag["outPSL"];
        const bool outPSLWithName = flag["outPSLWithName"];
        if (outPSLWithName) outPSL = true;
        char *err = nullptr;
        if (halgpu_open(pos[0].c_str(), opt.count("device") ? atoi(opt["device"].c_str()) : 0, &ctx, &err) != 0) {
            string m = err ? err : "cannot open";
            halgpu_free_string(err);
            throw runtime_error(m);
        }
        openSeconds = since(tMain);
        const int src = halgpu_genome_id(ctx, pos[1].c_str());
        if (src < 0) throw runtime_error(string("srcGenome, ") + pos[1] + ", not found in alignment");
        const int tgt = halgpu_genome_id(ctx, pos[3].c_str());
        if (tgt < 0) throw runtime_error(string("tgtGenome, ") + pos[3] + ", not found in alignment");
        int coal = -1;
        if (opt.count("coalescenceLimit") && !opt["coalescenceLimit"].empty()) {
            coal = halgpu_genome_id(ctx, opt["coalescenceLimit"].c_str());
            if (coal < 0) throw runtime_error("coalescence limit genome " + opt["coalescenceLimit"] + " not found in alignment\n");
        }
        ifstream srcBed;
        istream *in = &cin;
        if (pos[2] != "stdin") {
            srcBed.open(pos[2].c_str());
            in = &srcBed;
            if (!srcBed) throw runtime_error("Error opening srcBed, " + pos[2]);
        }
        ofstream tgtBed;
        ostream *out = &cout;
        if (pos[4] != "stdout") {
            tgtBed.open(pos[4].c_str(), flag["append"] ? ios::out | ios::app : ios_base::out);
            out = &tgtBed;
            if (!tgtBed) throw runtime_error("Error opening tgtBed, " + pos[4]);
        }
        halgpu::GpuBlockLiftover lift(ctx);
        lift.columnLiftover = flag["columnLiftover"];
        const auto tConvert = chrono::steady_clock::now();
        lift.convert(src, in, tgt, out, bedType, !flag["noDupes"], outPSL, outPSLWithName, coal);
        out->flush();
        const double convertSeconds = since(tConvert);
        if (getenv("HALGPU_TIMING")) {
            cerr << "[halLiftover] lines in " << lift.linesIn << " (" << lift.fastLines << " on the multi-threaded text path), intervals "
                 << lift.intervalsLifted << ", lines out " << lift.linesOut << "; text " << lift.textSeconds << " s (parse " << lift.parseSeconds << "), halgpu_liftover "
                 << lift.gpuSeconds << " s, read " << lift.readSeconds << " s, write " << lift.writeSeconds << " s, " << lift.textThreads
                 << " text threads; open (CUDA context; genomes are staged by the first lift) " << openSeconds << " s, total so far " << since(tMain) << " s" << endl;
            cerr << "[halLiftover] convert() wall " << convertSeconds << " s (read / tokenise, lift and format / write overlap in a pipeline)" << endl;
        }
    } catch (exception &e) {
        cerr << "hal exception caught: " << e.what() << endl;
        rc = 1;
    }
    // Nothing is left to do but to tear down the CUDA context, the staged genomes and the pinned buffers (0.1 - 1 s of a 2 - 3 s
    // process): every output stream is flushed and closed by now, so the process ends here.  HALGPU_CLEAN_EXIT=1 keeps the
    // orderly teardown (leak checkers, tools/sanitize.sh).
    if (getenv("HALGPU_CLEAN_EXIT") == nullptr) {
        if (getenv("HALGPU_TIMING") != nullptr) cerr << "[halLiftover] main() " << since(tMain) << " s (no teardown)" << endl;
        cout.flush();
        cerr.flush();
        fflush(nullptr);
        _exit(rc);
    }
    const auto tClose = chrono::steady_clock::now();
    halgpu_close(ctx);
    if (getenv("HALGPU_TIMING")) cerr << "[halLiftover] close " << since(tClose) << " s, main() " << since(tMain) << " s" << endl;
    return rc;
}
