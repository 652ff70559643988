/* TEST INFRASTRUCTURE ONLY -- see halview.h */
#include "halview.h"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace oracle {

namespace {
struct TreeNode { std::string name; std::vector<TreeNode> kids; };
TreeNode parseNewick(const char *&p) {
    TreeNode n;
    if (*p == '(') {
        p++;
        for (;;) {
            n.kids.push_back(parseNewick(p));
            if (*p == ',') { p++; continue; }
            if (*p == ')') { p++; break; }
            throw std::runtime_error("bad newick");
        }
    }
    const char *s = p;
    while (*p && *p != ':' && *p != ',' && *p != ')' && *p != ';') p++;
    n.name.assign(s, p - s);
    if (*p == ':') { p++; while (*p && *p != ',' && *p != ')' && *p != ';') p++; }
    return n;
}
void link(HalView &v, const TreeNode &n, int parent, int slot) {
    int id = v.genomeId(n.name);
    if (id < 0) throw std::runtime_error("newick genome not in genome array: " + n.name);
    GenomeView &g = v.genomes[id];
    g.parent = parent;
    g.slot = slot;
    if (parent < 0) v.root = id;
    for (size_t i = 0; i < n.kids.size(); i++) {
        g.children.push_back(v.genomeId(n.kids[i].name));
        link(v, n.kids[i], id, (int)i);
    }
}
} // namespace

void HalView::open(const std::string &path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    fstat(fd, &st);
    mapLen = (size_t)st.st_size;
    map = mmap(nullptr, mapLen, PROT_READ, MAP_SHARED, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) { map = nullptr; throw std::runtime_error("mmap failed"); }
    const uint8_t *base = (const uint8_t *)map;
    if (mapLen < 376 || strncmp((const char *)base, "HAL-MMAP", 32) != 0) throw std::runtime_error("not a HAL-MMAP file");
    auto rd = [&](size_t off) { uint64_t x; memcpy(&x, base + off, 8); return x; };
    size_t rootOff = rd(104);
    if (base[112]) throw std::runtime_error("file is marked dirty");
    size_t numGenomes = rd(rootOff), nwOff = rd(rootOff + 8), nwLen = rd(rootOff + 16), gaOff = rd(rootOff + 24);
    newick.assign((const char *)base + nwOff, strnlen((const char *)base + nwOff, nwLen));
    genomes.resize(numGenomes);
    for (size_t i = 0; i < numGenomes; i++) {
        const size_t g0 = gaOff + 96 * i;
        GenomeView &g = genomes[i];
        g.len = (int64_t)rd(g0);
        size_t nseq = rd(g0 + 8);
        g.numTop = (int64_t)rd(g0 + 16);
        g.numBot = (int64_t)rd(g0 + 24);
        size_t nameOff = rd(g0 + 32), seqOff = rd(g0 + 56), dnaOff = rd(g0 + 72), topOff = rd(g0 + 80), botOff = rd(g0 + 88);
        g.name = (const char *)base + nameOff + 24;
        g.dna = base + dnaOff;
        g.top = base + topOff;
        g.bot = base + botOff;
        g.seqs.resize(nseq);
        for (size_t s = 0; s < nseq; s++) {
            size_t s0 = seqOff + 328 * s;
            SeqView &q = g.seqs[s];
            q.start = (int64_t)rd(s0);
            q.length = (int64_t)rd(s0 + 16);
            q.topStart = (int64_t)rd(s0 + 24);
            q.botStart = (int64_t)rd(s0 + 32);
            q.numTop = (int64_t)rd(s0 + 40);
            q.numBot = (int64_t)rd(s0 + 48);
            size_t nl = rd(s0 + 56), no = rd(s0 + 64);
            q.name.assign((const char *)base + no, strnlen((const char *)base + no, nl));
        }
    }
    const char *p = newick.c_str();
    TreeNode t = parseNewick(p);
    link(*this, t, -1, -1);
    for (auto &g : genomes) {
        g.nc = (int)g.children.size();
        g.bstride = 8 * (2 + g.nc) + ((g.nc + 7) / 8) * 8;
    }
}

void HalView::close() {
    if (map) munmap(map, mapLen);
    map = nullptr;
}

} // namespace oracle
