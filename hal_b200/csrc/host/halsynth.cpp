// halSynth -- synthetic HAL-MMAP writer for benchmarks and tests ("halRandGen of a named shape").
//
// halRandGen cannot be steered to an exact (genomes, levels) tree (api/tests/halRandomData.cpp:107-113) and
// generates ~12 Mbp/s; BASELINE.json's configs name exact shapes of up to 64 x 100 Mbp.  This tool writes a
// complete, reference-readable HAL-MMAP file (header, genome array, sequence tables, name perfect hashes,
// site maps, DNA, top/bottom segment arrays -- layout per api/mmap_impl/*.h) for an explicit newick tree
// with the structural rules halRandGen's generator applies per segment (api/tests/halRandomData.cpp:268-346):
// child top segment i maps to parent bottom segment i; with probability 1-exp(-branch) it is transposed to
// a random parent segment (creating paralogy rings), with the square of that it is an insertion, aligned
// segments are inverted with the same probability, and the last segment of every genome stays unaligned.
// With --branch 0 the structure is exactly what the reference generator produces for equal dimensions
// (checked by tests/test_synth.py against oracle/_ref/halTreeGen).  One sequence "<genome>_seq" per genome,
// all segments --segLen long.  The file is validated by the reference's own halValidate in the tests.
//
// usage: halSynth --newick T --segs N --segLen L [--branch b] [--seed s] out.hal
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>

namespace {

struct Node {
    std::string name;
    int parent = -1, slot = -1;
    std::vector<int> kids;
};

int parseNewick(const std::string &s, size_t &i, std::vector<Node> &nodes, int parent) {
    const int id = (int)nodes.size();
    nodes.push_back(Node());
    nodes[id].parent = parent;
    if (i < s.size() && s[i] == '(') {
        ++i;
        while (true) {
            int k = parseNewick(s, i, nodes, id);
            nodes[k].slot = (int)nodes[id].kids.size();
            nodes[id].kids.push_back(k);
            if (i < s.size() && s[i] == ',') { ++i; continue; }
            if (i < s.size() && s[i] == ')') { ++i; break; }
            fprintf(stderr, "halSynth: malformed newick\n");
            exit(1);
        }
    }
    size_t b = i;
    while (i < s.size() && s[i] != ':' && s[i] != ',' && s[i] != ')' && s[i] != ';') ++i;
    nodes[id].name = s.substr(b, i - b);
    if (i < s.size() && s[i] == ':') while (i < s.size() && s[i] != ',' && s[i] != ')' && s[i] != ';') ++i;
    return id;
}

void printNewick(const std::vector<Node> &nodes, int id, std::string &out) { // stTree_getNewickTreeString with ":0" lengths
    if (!nodes[id].kids.empty()) {
        out += '(';
        for (size_t k = 0; k < nodes[id].kids.size(); ++k) {
            if (k) out += ',';
            printNewick(nodes, nodes[id].kids[k], out);
        }
        out += ')';
    }
    out += nodes[id].name;
    if (nodes[id].parent >= 0) out += ":0";
}

struct Rng { // splitmix64
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double u01() { return (double)(next() >> 11) / 9007199254740992.0; }
};

// ---- the reference's string perfect hash (wahern/phf as vendored in api/mmap_impl/mmapPhf.cpp:311-402):
//      MurmurHash3-style rounds over big-endian 4-byte words, displacement d mixed in first ----
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
inline uint32_t round32(uint32_t k1, uint32_t h1) {
    k1 *= 0xcc9e2d51u; k1 = rotl(k1, 15); k1 *= 0x1b873593u;
    h1 ^= k1; h1 = rotl(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
    return h1;
}
inline uint32_t roundStr(const std::string &k, uint32_t h1) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(k.data());
    size_t n = k.size();
    while (n >= 4) {
        h1 = round32(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3], h1);
        p += 4; n -= 4;
    }
    uint32_t k1 = 0;
    switch (n & 3) {
    case 3: k1 |= (uint32_t)p[2] << 8; // fallthrough
    case 2: k1 |= (uint32_t)p[1] << 16; // fallthrough
    case 1: k1 |= (uint32_t)p[0] << 24; h1 = round32(k1, h1);
    }
    return h1;
}
inline uint32_t mix32(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }
inline uint32_t phfG(const std::string &k, uint32_t seed) { return mix32(roundStr(k, seed)); }
inline uint32_t phfF(uint32_t d, const std::string &k, uint32_t seed) { return mix32(roundStr(k, round32(d, seed))); }

struct Pht { // a built table: r displacements (uint32), m slots
    size_t r = 1, m = 1;
    uint32_t dmax = 0;
    std::vector<uint32_t> g;
    std::vector<int64_t> table;
};

Pht buildPht(const std::vector<std::string> &keys) { // compress-hash-displace, buckets largest first
    Pht t;
    const size_t n = keys.size();
    while (t.r < n) t.r <<= 1;
    while (t.m < n + n / 4 + 1) t.m <<= 1;
    for (;; t.m <<= 1) {
        std::vector<std::vector<size_t>> buckets(t.r);
        for (size_t i = 0; i < n; ++i) buckets[phfG(keys[i], 0) & (t.r - 1)].push_back(i);
        std::vector<size_t> order(t.r);
        for (size_t i = 0; i < t.r; ++i) order[i] = i;
        for (size_t i = 1; i < t.r; ++i) // insertion sort by size, descending (r is small)
            for (size_t j = i; j > 0 && buckets[order[j]].size() > buckets[order[j - 1]].size(); --j) std::swap(order[j], order[j - 1]);
        t.g.assign(t.r, 0);
        t.table.assign(t.m, -1);
        std::vector<char> used(t.m, 0);
        bool ok = true;
        t.dmax = 0;
        for (size_t oi = 0; oi < t.r && ok; ++oi) {
            const std::vector<size_t> &b = buckets[order[oi]];
            if (b.empty()) break;
            uint32_t d = 1;
            for (; d < 100000; ++d) {
                std::vector<size_t> slots;
                bool clash = false;
                for (size_t i : b) {
                    size_t s = phfF(d, keys[i], 0) & (t.m - 1);
                    for (size_t q : slots) clash |= q == s;
                    if (used[s] || clash) { clash = true; break; }
                    slots.push_back(s);
                }
                if (!clash) {
                    for (size_t s : slots) used[s] = 1;
                    break;
                }
            }
            if (d >= 100000) { ok = false; break; }
            t.g[order[oi]] = d;
            if (d > t.dmax) t.dmax = d;
        }
        if (ok) break;
    }
    for (size_t i = 0; i < n; ++i) t.table[phfF(t.g[phfG(keys[i], 0) & (t.r - 1)], keys[i], 0) & (t.m - 1)] = (int64_t)i;
    return t;
}

inline size_t al8(size_t n) { return (n + 7) & ~size_t(7); }
inline size_t phtBytes(const Pht &t) { return 64 + al8(t.r * 4) + t.m * 8; }

struct Writer {
    uint8_t *base = nullptr;
    size_t next = 0;
    size_t alloc(size_t n) { size_t o = next; next += al8(n); return o; }
    void u64(size_t off, uint64_t v) { std::memcpy(base + off, &v, 8); }
    void i64(size_t off, int64_t v) { std::memcpy(base + off, &v, 8); }
};

size_t writePht(Writer &w, const Pht &t) { // PerfectHashTableData, api/mmap_impl/mmapPerfectHashTable.h:14-33
    const size_t off = w.alloc(phtBytes(t));
    w.base[off] = 1;                           // nodiv
    uint32_t seed = 0, gop = 6;                // PHF_G_UINT32_BAND_R
    std::memcpy(w.base + off + 4, &seed, 4);
    w.u64(off + 8, t.r); w.u64(off + 16, t.m); w.u64(off + 24, t.dmax);
    std::memcpy(w.base + off + 32, &gop, 4);
    const size_t gOff = off + 64, hOff = gOff + al8(t.r * 4);
    w.u64(off + 40, gOff); w.u64(off + 48, hOff); w.u64(off + 56, phtBytes(t));
    std::memcpy(w.base + gOff, t.g.data(), t.r * 4);
    std::memcpy(w.base + hOff, t.table.data(), t.m * 8);
    return off;
}

} // namespace

int main(int argc, char **argv) {
    std::string newick = "((G3)G1,G2)G0;", out;
    long segs = 1000, segLen = 32;
    double branch = 0;
    uint64_t seed = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--", 0) == 0 && i + 1 < argc) {
            std::string v = argv[++i];
            if (a == "--newick") newick = v; else if (a == "--segs") segs = atol(v.c_str());
            else if (a == "--segLen") segLen = atol(v.c_str()); else if (a == "--branch") branch = atof(v.c_str());
            else if (a == "--seed") seed = strtoull(v.c_str(), nullptr, 10);
            else { fprintf(stderr, "halSynth: unknown option %s\n", a.c_str()); return 1; }
        } else out = a;
    }
    if (out.empty() || segs < 2 || segLen < 1) { fprintf(stderr, "usage: halSynth --newick T --segs N --segLen L [--branch b] [--seed s] out.hal\n"); return 1; }
    std::vector<Node> nodes;
    size_t pi = 0;
    parseNewick(newick, pi, nodes, -1);
    const int G = (int)nodes.size();
    const int64_t N = segs, L = segLen, len = N * L;
    const double pEvent = 1.0 - std::exp(-branch);
    Rng rng{seed};
    std::string tree;
    printNewick(nodes, 0, tree);
    tree += ';';

    // ---- size the file ----
    std::vector<std::string> gnames;
    for (auto &n : nodes) gnames.push_back(n.name);
    Pht gpht = buildPht(gnames);
    std::vector<Pht> spht(G);
    size_t total = 376 + 312 + al8(tree.size() + 1) + (size_t)G * 96 + phtBytes(gpht);
    for (int g = 0; g < G; ++g) {
        spht[g] = buildPht({nodes[g].name + "_seq"});
        const size_t nc = nodes[g].kids.size();
        const size_t bstride = 8 * (2 + nc) + ((nc + 7) / 8) * 8;
        const int64_t nTop = nodes[g].parent >= 0 ? N : 0, nBot = nc > 0 ? N : 0;
        total += al8(24 + nodes[g].name.size() + 1) + al8(328 + 1) + al8(nodes[g].name.size() + 5) + al8((len + 1) / 2) +
                 al8((nTop + 1) * 40) + al8((nBot + 1) * bstride) + phtBytes(spht[g]) + al8(8 + 40) + al8(272) + 2 * 24;
    }
    int fd = open(out.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0 || ftruncate(fd, (off_t)total) != 0) { perror("halSynth: output"); return 1; }
    Writer w;
    w.base = static_cast<uint8_t *>(mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
    if (w.base == MAP_FAILED) { perror("halSynth: mmap"); return 1; }

    // ---- header + root (api/mmap_impl/mmapFile.h:23-31, mmapAlignment.h:14-31) ----
    w.alloc(376);
    std::strcpy(reinterpret_cast<char *>(w.base), "HAL-MMAP");
    std::strcpy(reinterpret_cast<char *>(w.base + 32), "1.1");
    std::strcpy(reinterpret_cast<char *>(w.base + 64), "2.2");
    const size_t root = w.alloc(312);
    w.u64(104, root);
    const size_t nwOff = w.alloc(tree.size() + 1);
    std::memcpy(w.base + nwOff, tree.c_str(), tree.size() + 1);
    const size_t gaOff = w.alloc((size_t)G * 96);
    w.u64(root, G); w.u64(root + 8, nwOff); w.u64(root + 16, tree.size() + 1); w.u64(root + 24, gaOff);

    std::vector<std::string> dna(G);
    std::vector<size_t> topOff(G), botOff(G), bstrideOf(G);
    // BFS order == newick pre-order here is fine: parents precede children in `nodes`
    for (int g = 0; g < G; ++g) {
        const Node &nd = nodes[g];
        const size_t nc = nd.kids.size();
        const size_t bstride = 8 * (2 + nc) + ((nc + 7) / 8) * 8;
        bstrideOf[g] = bstride;
        const int64_t nTop = nd.parent >= 0 ? N : 0, nBot = nc > 0 ? N : 0;
        const size_t gb = gaOff + (size_t)g * 96;
        // name (MMapString: 24-byte array header + chars)
        const size_t nameOff = w.alloc(24 + nd.name.size() + 1);
        w.u64(nameOff, 1); w.u64(nameOff + 8, nd.name.size() + 1); w.u64(nameOff + 16, nd.name.size() + 1);
        std::memcpy(w.base + nameOff + 24, nd.name.c_str(), nd.name.size() + 1);
        // one sequence
        const std::string sname = nd.name + "_seq";
        const size_t seqOff = w.alloc(328 + 1), snOff = w.alloc(sname.size() + 1);
        std::memcpy(w.base + snOff, sname.c_str(), sname.size() + 1);
        w.i64(seqOff, 0); w.i64(seqOff + 8, 0); w.u64(seqOff + 16, len); w.i64(seqOff + 24, 0); w.i64(seqOff + 32, 0);
        w.u64(seqOff + 40, nTop); w.u64(seqOff + 48, nBot); w.u64(seqOff + 56, sname.size() + 1); w.u64(seqOff + 64, snOff);
        const size_t dnaOff = w.alloc((len + 1) / 2);
        topOff[g] = w.alloc((nTop + 1) * 40);
        botOff[g] = w.alloc((nBot + 1) * bstride);
        const size_t shOff = writePht(w, spht[g]);
        // site map: one node (api/mmap_impl/mmapGenomeSiteMap.h:20-60)
        const size_t smOff = w.alloc(8 + 40);
        w.u64(smOff, 1); w.u64(smOff + 8, 0); w.u64(smOff + 16, len); w.i64(smOff + 24, 0); w.i64(smOff + 32, -1); w.i64(smOff + 40, -1);
        // empty metadata (api/mmap_impl/mmapMetaData.h:9-15)
        const size_t mdOff = w.alloc(272), kOff = w.alloc(24), vOff = w.alloc(24);
        w.u64(mdOff, kOff); w.u64(mdOff + 8, vOff);
        w.u64(kOff, 8); w.u64(vOff, 8);
        w.u64(gb, len); w.u64(gb + 8, 1); w.u64(gb + 16, nTop); w.u64(gb + 24, nBot);
        w.u64(gb + 32, nameOff); w.u64(gb + 40, shOff); w.u64(gb + 48, smOff); w.u64(gb + 56, seqOff); w.u64(gb + 64, mdOff);
        w.u64(gb + 72, dnaOff); w.u64(gb + 80, topOff[g]); w.u64(gb + 88, botOff[g]);

        // bottom segments: start, topParseIndex, children NULL for now
        uint8_t *bot = w.base + botOff[g];
        for (int64_t i = 0; i <= nBot; ++i) {
            uint8_t *r = bot + (size_t)i * bstride;
            int64_t v = i * L;
            std::memcpy(r, &v, 8);
            v = (i < nBot && nTop > 0) ? i : -1;
            std::memcpy(r + 8, &v, 8);
            v = -1;
            for (size_t k = 0; k < nc; ++k) std::memcpy(r + 16 + 8 * k, &v, 8);
        }
        // top segments + DNA
        std::string &seq = dna[g];
        seq.resize(len);
        static const char ACGT[4] = {'A', 'C', 'G', 'T'};
        auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; };
        if (nd.parent < 0) {
            for (int64_t i = 0; i < len; ++i) seq[i] = ACGT[rng.next() & 3];
        } else {
            const int p = nd.parent;
            uint8_t *top = w.base + topOff[g];
            uint8_t *pbot = w.base + botOff[p];
            const size_t pstride = bstrideOf[p], pnc = nodes[p].kids.size();
            std::map<int64_t, std::vector<int64_t>> multi; // parent -> tops, only where shared
            std::vector<int64_t> lastTop(N, -1);
            for (int64_t i = 0; i <= nTop; ++i) {
                uint8_t *r = top + (size_t)i * 40;
                int64_t start = i * L, parse = (i < nTop && nBot > 0) ? i : -1, para = -1, par = -1;
                bool rev = false;
                if (i < nTop) {
                    par = i;
                    if (branch > 0 && rng.u01() <= pEvent) par = (int64_t)(rng.next() % (uint64_t)N);
                    else if (branch > 0 && rng.u01() <= pEvent && rng.u01() <= pEvent) par = -1;
                    if (par == N - 1 || i == nTop - 1) par = -1;
                    if (par >= 0) {
                        rev = branch > 0 && rng.u01() <= pEvent;
                        uint8_t *pr = pbot + (size_t)par * pstride;
                        std::memcpy(pr + 16 + 8 * nd.slot, &i, 8); // latest child wins (canonical), as in the reference generator
                        pr[16 + 8 * pnc + nd.slot] = rev ? 1 : 0;
                        if (lastTop[par] >= 0) {
                            auto &v = multi[par];
                            if (v.empty()) v.push_back(lastTop[par]);
                            v.push_back(i);
                        }
                        lastTop[par] = i;
                        const std::string &ps = dna[p];
                        for (int64_t k = 0; k < L; ++k) {
                            char c = rev ? comp(ps[par * L + L - 1 - k]) : ps[par * L + k];
                            if (branch > 0 && rng.u01() <= pEvent) c = ACGT[rng.next() & 3];
                            seq[start + k] = c;
                        }
                    } else {
                        for (int64_t k = 0; k < L; ++k) seq[start + k] = ACGT[rng.next() & 3];
                    }
                }
                std::memcpy(r, &start, 8); std::memcpy(r + 8, &parse, 8); std::memcpy(r + 16, &para, 8); std::memcpy(r + 24, &par, 8);
                r[32] = rev ? 1 : 0;
            }
            for (auto &kv : multi) { // circular paralogy ring in index order
                const auto &v = kv.second;
                for (size_t k = 0; k < v.size(); ++k) {
                    int64_t nx = v[(k + 1) % v.size()];
                    std::memcpy(top + (size_t)v[k] * 40 + 16, &nx, 8);
                }
            }
        }
        uint8_t *d = w.base + dnaOff;
        for (int64_t i = 0; i < len; ++i) {
            const char c = seq[i];
            const uint8_t nib = 8 | (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3);
            if (i & 1) d[i >> 1] |= nib; else d[i >> 1] = (uint8_t)(nib << 4);
        }
        if (nd.kids.empty()) std::string().swap(dna[g]);
    }
    // genome name hash
    const size_t ghOff = writePht(w, gpht);
    w.u64(root + 32, ghOff);
    uint8_t *ht = w.base + ghOff + 64 + al8(gpht.r * 4);
    (void)ht; // table already holds the genome array indices (keys were given in array order)
    w.u64(96, w.next); // nextOffset
    if (w.next > total) { fprintf(stderr, "halSynth: internal size error\n"); return 1; }
    munmap(w.base, total);
    if (ftruncate(fd, (off_t)w.next) != 0) perror("halSynth: truncate");
    close(fd);
    return 0;
}
