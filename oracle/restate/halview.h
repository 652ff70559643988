/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
 *
 * Read-only view of a HAL-MMAP file: just enough of the on-disk struct layout to reach the
 * top/bottom segment arrays, DNA and sequence tables.  Layout restated from
 *   api/mmap_impl/mmapFile.h:23-31        (MMapHeader: 3x char[32], nextOffset, rootOffset, dirty)
 *   api/mmap_impl/mmapAlignment.h:14-31   (MMapAlignmentData)
 *   api/mmap_impl/mmapGenome.h:19-46      (MMapGenomeData, 96 B)
 *   api/mmap_impl/mmapSequenceData.h:21-30 (MMapSequenceData, 328 B)
 *   api/mmap_impl/mmapTopSegmentData.h:40-44, mmapBottomSegmentData.h:35-52
 *   api/mmap_impl/mmapArray.h:6-11, mmapString.h (genome name = 24 B array header + chars)
 * Tree shape and child order come from the stored newick string (mmapAlignment.h:145-153,216-224).
 */
#ifndef ORACLE_HALVIEW_H
#define ORACLE_HALVIEW_H
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

struct SeqView {
    std::string name;
    int64_t start, length, topStart, botStart, numTop, numBot;
};

struct GenomeView {
    std::string name;
    int parent = -1;
    int slot = -1; /* index of this genome in parent's child list */
    std::vector<int> children;
    int64_t len = 0, numTop = 0, numBot = 0;
    int nc = 0;
    size_t bstride = 16;
    const uint8_t *top = nullptr, *bot = nullptr, *dna = nullptr;
    std::vector<SeqView> seqs;

    static int64_t rd(const uint8_t *p) { int64_t v; memcpy(&v, p, 8); return v; }
    int64_t tStart(int64_t i) const { return rd(top + 40 * i); }
    int64_t tBotParse(int64_t i) const { return rd(top + 40 * i + 8); }
    int64_t tNextPara(int64_t i) const { return rd(top + 40 * i + 16); }
    int64_t tParent(int64_t i) const { return rd(top + 40 * i + 24); }
    bool tRev(int64_t i) const { return top[40 * i + 32] != 0; }
    int64_t bStart(int64_t i) const { return rd(bot + bstride * i); }
    int64_t bTopParse(int64_t i) const { return rd(bot + bstride * i + 8); }
    int64_t bChild(int64_t i, int k) const { return rd(bot + bstride * i + 16 + 8 * k); }
    bool bChildRev(int64_t i, int k) const { return bot[bstride * i + 16 + 8 * nc + k] != 0; }
    /* index of the sequence containing genome position pos (sequences are stored in start order,
     * api/mmap_impl/mmapGenome.cpp:52-59) -- same answer as the stored BST, mmapGenomeSiteMap.cpp:99-113 */
    int seqOf(int64_t pos) const {
        int lo = 0, hi = (int)seqs.size() - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) / 2;
            if (seqs[mid].start <= pos) lo = mid; else hi = mid - 1;
        }
        return lo;
    }
    char base(int64_t pos) const {
        static const char tbl[16] = {'a', 'c', 'g', 't', 'n', '?', '?', '?', 'A', 'C', 'G', 'T', 'N', '?', '?', '?'};
        uint8_t b = dna[pos / 2];
        return tbl[(pos & 1) ? (b & 0xF) : (b >> 4)];
    }
};

struct HalView {
    void *map = nullptr;
    size_t mapLen = 0;
    std::string newick;
    std::vector<GenomeView> genomes; /* indexed in file genome-array order */
    int root = -1;

    void open(const std::string &path);
    void close();
    int genomeId(const std::string &name) const {
        for (size_t i = 0; i < genomes.size(); i++) if (genomes[i].name == name) return (int)i;
        return -1;
    }
    ~HalView() { close(); }
};

} // namespace oracle
#endif
