/* TEST INFRASTRUCTURE ONLY.  The reference compiles hal::ColumnLiftover into libHalLiftover but no CLI
 * reaches it (liftover/impl/halLiftoverMain.cpp:138 instantiates BlockLiftover).  This driver exposes it
 * with the same positional arguments so it can serve as the oracle for SURVEY.md 8(a) row A11:
 *   halColumnLiftoverCli in.hal srcGenome in.bed tgtGenome out.bed [--noDupes]                     */
#include "halColumnLiftover.h"
#include "halAlignmentInstance.h"
#include <fstream>
#include <iostream>
using namespace hal;
using namespace std;
int main(int argc, char **argv) {
    if (argc < 6) { cerr << "usage: " << argv[0] << " in.hal src in.bed tgt out.bed [--noDupes]" << endl; return 1; }
    try {
        bool dupes = !(argc > 6 && string(argv[6]) == "--noDupes");
        AlignmentConstPtr aln(openHalAlignment(argv[1], NULL));
        const Genome *src = aln->openGenome(argv[2]);
        const Genome *tgt = aln->openGenome(argv[4]);
        if (src == NULL || tgt == NULL) throw hal_exception("genome not found");
        ifstream in(argv[3]);
        ofstream out(argv[5]);
        ColumnLiftover lift;
        lift.convert(aln, src, &in, tgt, &out, 0, dupes);
    } catch (exception &e) {
        cerr << "exception: " << e.what() << endl;
        return 1;
    }
    return 0;
}
