// Test driver for hal_b200/csrc/host/hal_api.hpp: lifts single bases src -> tgt with the reference's iterator
// vocabulary only (toSite / toParent / toParseUp / toParseDown / toChild, no paralogy walk == halLiftover --noDupes
// restricted to one base) and prints "srcPos tgtSeqName tgtPos strand" or "srcPos -".  tests/test_hal_api.py compares
// the output with the CPU oracle.
#include "../../hal_b200/csrc/host/hal_api.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace hal;

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s hal src tgt n seed\n", argv[0]); return 2; }
    try {
        AlignmentConstPtr aln = openHalAlignment(argv[1]);
        const Genome *src = aln->openGenome(argv[2]), *tgt = aln->openGenome(argv[3]);
        if (!src || !tgt) throw hal_exception("genome not found");
        std::set<const Genome *> in{src, tgt};
        const Genome *mrca = getLowestCommonAncestor(in);
        std::vector<const Genome *> down;
        for (const Genome *g = tgt; g != mrca; g = g->getParent()) down.push_back(g);
        std::reverse(down.begin(), down.end());
        std::mt19937_64 rng(strtoull(argv[5], nullptr, 10));
        const long n = atol(argv[4]);
        printf("# %s %s root=%s mrca=%s newick=%s\n", src->getName().c_str(), tgt->getName().c_str(), aln->getRootName().c_str(),
               mrca->getName().c_str(), aln->getNewickTree().c_str());
        for (long i = 0; i < n; ++i) {
            const hal_index_t pos = (hal_index_t)(rng() % src->getSequenceLength());
            TopSegmentIteratorPtr top;
            BottomSegmentIteratorPtr bot;
            bool ok = true, isTop = src->getNumTopSegments() > 0;
            if (isTop) { top = src->getTopSegmentIterator(); top->toSite(pos, true); }
            else { bot = src->getBottomSegmentIterator(); bot->toSite(pos, true); }
            const Genome *cur = src;
            while (ok && cur != mrca) { // up: toParseUp (if we hold a bottom) then toParent
                if (!isTop) { top = cur->getTopSegmentIterator(); top->toParseUp(bot); isTop = true; }
                if (!top->tseg()->hasParent()) { ok = false; break; }
                bot = cur->getParent()->getBottomSegmentIterator();
                bot->toParent(top);
                isTop = false;
                cur = cur->getParent();
            }
            for (size_t d = 0; ok && d < down.size(); ++d) { // down: toParseDown (if we hold a top) then toChild
                if (isTop) { bot = cur->getBottomSegmentIterator(); bot->toParseDown(top); isTop = false; }
                const hal_index_t slot = cur->getChildIndex(down[d]);
                if (!bot->bseg()->hasChild((hal_size_t)slot)) { ok = false; break; }
                top = down[d]->getTopSegmentIterator();
                top->toChild(bot, (hal_size_t)slot);
                isTop = true;
                cur = down[d];
            }
            if (!ok) { printf("%ld -\n", (long)pos); continue; }
            const SegmentIterator *it = isTop ? (const SegmentIterator *)top.get() : (const SegmentIterator *)bot.get();
            if (it->getLength() != 1) { printf("%ld LEN%ld\n", (long)pos, (long)it->getLength()); continue; }
            const Sequence *sq = cur->getSequenceBySite((hal_size_t)it->getStartPosition());
            std::string base;
            it->getString(base);
            printf("%ld %s %ld %c %s\n", (long)pos, sq->getName().c_str(), (long)(it->getStartPosition() - sq->getStartPosition()),
                   it->getReversed() ? '-' : '+', base.c_str());
        }
    } catch (std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
