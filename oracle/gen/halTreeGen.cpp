/* TEST INFRASTRUCTURE ONLY.  Explicit-tree synthetic HAL generator, linked against the oracle build
 * of the reference (oracle/_ref/libhal.a) and written against its PUBLIC write API
 * (api/inc/halAlignment.h, halGenome.h, halTopSegment.h, halBottomSegment.h).
 *
 * Why it exists: halRandGen cannot be steered to an exact (genomes, levels) shape
 * (api/tests/halRandomData.cpp:107-113), and its alignments have no inversions or non-coincident
 * parse geometry.  Two modes:
 *   --mode randgen : explicit newick + the reference's own createRandomDimensions/createRandomGenome
 *                    (api/tests/halRandomData.h:24-27) -> "halRandGen-faithful" content.
 *   --mode varlen  : own generator: variable segment lengths (all four parse geometries), several
 *                    sequences per genome, inversions, transpositions/duplications (paralogy rings
 *                    with a random canonical member), insertions and deletions.
 * usage: halTreeGen [--mode M] [--seed S] [--newick T] [--segs N] [--minLen a] [--maxLen b] [--seqs K]
 *                   [--branch x] [--pInv p] [--pDup p] [--pIns p] [--pDel p] [--fileGB g] out.hal
 */
#include "hal.h"
#include "halAlignmentInstance.h"
#include "halRandNumberGen.h"
#include "halRandomData.h"
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <vector>

using namespace hal;
using namespace std;

struct Node {
    string name;
    double branch;
    vector<Node *> kids;
    Node() : branch(0) {}
};

static Node *parseNewick(const char *&p) {
    Node *n = new Node;
    if (*p == '(') {
        p++;
        for (;;) {
            n->kids.push_back(parseNewick(p));
            if (*p == ',') { p++; continue; }
            if (*p == ')') { p++; break; }
            cerr << "bad newick near: " << p << endl;
            exit(1);
        }
    }
    const char *s = p;
    while (*p && *p != ':' && *p != ',' && *p != ')' && *p != ';') p++;
    n->name.assign(s, p - s);
    if (*p == ':') {
        char *e;
        n->branch = strtod(p + 1, &e);
        p = e;
    }
    return n;
}

struct Opts {
    string mode, newick, out;
    unsigned seed;
    long segs, minLen, maxLen, seqs;
    double branch, pInv, pDup, pIns, pDel, pMut;
    size_t fileGB;
};

static void addTree(AlignmentPtr aln, Node *n, Node *parent, const Opts &o) {
    double b = o.branch >= 0 ? o.branch : n->branch;
    if (parent == NULL) {
        aln->addRootGenome(n->name);
    } else {
        aln->addLeafGenome(n->name, parent->name, b);
    }
    for (size_t i = 0; i < n->kids.size(); i++) addTree(aln, n->kids[i], n, o);
}

/* ---- varlen mode --------------------------------------------------------------------------- */
struct GenomeModel {
    vector<long> seqLen, seqTop, seqBot;      // per sequence
    vector<long> topStart, topLen, topParent; // per top segment (global coords)
    vector<char> topRev;
    vector<long> botStart, botLen;            // per bottom segment
    string dna;
};

static const char DNA4[4] = {'A', 'C', 'G', 'T'};
static char comp(char c) {
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    }
    return c;
}

static void cutIntoSequences(mt19937_64 &rng, long nItems, long k, vector<long> &groupSize) {
    /* split nItems consecutive items into k non-empty consecutive groups */
    if (k > nItems) k = nItems;
    if (k < 1) k = 1;
    vector<long> cuts;
    while ((long)cuts.size() < k - 1) {
        long c = 1 + (long)(rng() % (unsigned long)(nItems - 1));
        bool dup = false;
        for (size_t i = 0; i < cuts.size(); i++) dup |= (cuts[i] == c);
        if (!dup) cuts.push_back(c);
    }
    sort(cuts.begin(), cuts.end());
    long prev = 0;
    groupSize.clear();
    for (size_t i = 0; i < cuts.size(); i++) { groupSize.push_back(cuts[i] - prev); prev = cuts[i]; }
    groupSize.push_back(nItems - prev);
}

static void makeBottoms(mt19937_64 &rng, GenomeModel &g, const Opts &o) {
    g.seqBot.assign(g.seqLen.size(), 0);
    long pos = 0;
    for (size_t s = 0; s < g.seqLen.size(); s++) {
        long end = pos + g.seqLen[s];
        while (pos < end) {
            long l = o.minLen + (long)(rng() % (unsigned long)(o.maxLen - o.minLen + 1));
            if (pos + l > end) l = end - pos;
            g.botStart.push_back(pos);
            g.botLen.push_back(l);
            g.seqBot[s]++;
            pos += l;
        }
    }
}

static double u01(mt19937_64 &rng) { return (double)(rng() >> 11) / 9007199254740992.0; }

static void buildVarlen(AlignmentPtr aln, Node *root, const Opts &o) {
    mt19937_64 rng(o.seed);
    map<string, GenomeModel> models;
    deque<pair<Node *, Node *> > q;
    q.push_back(make_pair(root, (Node *)NULL));
    vector<pair<Node *, Node *> > order;
    while (!q.empty()) {
        pair<Node *, Node *> cur = q.front();
        q.pop_front();
        order.push_back(cur);
        Node *n = cur.first;
        GenomeModel &g = models[n->name];
        if (cur.second == NULL) {
            /* root: bottoms only */
            vector<long> lens(o.segs);
            long total = 0;
            for (long i = 0; i < o.segs; i++) { lens[i] = o.minLen + (long)(rng() % (unsigned long)(o.maxLen - o.minLen + 1)); total += lens[i]; }
            vector<long> grp;
            cutIntoSequences(rng, o.segs, o.seqs, grp);
            long idx = 0, pos = 0;
            for (size_t s = 0; s < grp.size(); s++) {
                long sl = 0;
                for (long j = 0; j < grp[s]; j++, idx++) { g.botStart.push_back(pos); g.botLen.push_back(lens[idx]); pos += lens[idx]; sl += lens[idx]; }
                g.seqLen.push_back(sl);
                g.seqBot.push_back(grp[s]);
                g.seqTop.push_back(0);
            }
            g.dna.resize(total);
            for (long i = 0; i < total; i++) g.dna[i] = DNA4[rng() & 3];
        } else {
            GenomeModel &p = models[cur.second->name];
            long nb = (long)p.botLen.size();
            long ntop = (long)(nb * (0.85 + 0.3 * u01(rng)));
            if (ntop < 2) ntop = 2;
            long cursor = 0, pos = 0;
            for (long i = 0; i < ntop; i++) {
                long par = -1, len;
                double r = u01(rng);
                if (r < o.pIns) {
                    par = -1;
                } else if (r < o.pIns + o.pDup || cursor >= nb) {
                    par = (long)(rng() % (unsigned long)nb);
                } else {
                    while (cursor < nb - 1 && u01(rng) < o.pDel) cursor++;
                    par = cursor++;
                }
                len = par >= 0 ? p.botLen[par] : o.minLen + (long)(rng() % (unsigned long)(o.maxLen - o.minLen + 1));
                g.topStart.push_back(pos);
                g.topLen.push_back(len);
                g.topParent.push_back(par);
                g.topRev.push_back(par >= 0 && u01(rng) < o.pInv);
                pos += len;
            }
            vector<long> grp;
            cutIntoSequences(rng, ntop, o.seqs, grp);
            long idx = 0;
            for (size_t s = 0; s < grp.size(); s++) {
                long sl = 0;
                for (long j = 0; j < grp[s]; j++, idx++) sl += g.topLen[idx];
                g.seqLen.push_back(sl);
                g.seqTop.push_back(grp[s]);
            }
            g.seqBot.assign(g.seqLen.size(), 0);
            if (!n->kids.empty()) makeBottoms(rng, g, o);
            /* dna */
            g.dna.resize(pos);
            for (long i = 0; i < ntop; i++) {
                long st = g.topStart[i], len = g.topLen[i], par = g.topParent[i];
                for (long k = 0; k < len; k++) {
                    char c;
                    if (par < 0 || u01(rng) < o.pMut) c = DNA4[rng() & 3];
                    else if (!g.topRev[i]) c = p.dna[p.botStart[par] + k];
                    else c = comp(p.dna[p.botStart[par] + len - 1 - k]);
                    g.dna[st + k] = c;
                }
            }
        }
        for (size_t i = 0; i < n->kids.size(); i++) q.push_back(make_pair(n->kids[i], n));
    }
    /* dimensions first (all genomes), then links */
    for (size_t gi = 0; gi < order.size(); gi++) {
        Node *n = order[gi].first;
        GenomeModel &g = models[n->name];
        Genome *genome = aln->openGenome(n->name);
        vector<Sequence::Info> dims;
        for (size_t s = 0; s < g.seqLen.size(); s++) {
            dims.push_back(Sequence::Info(n->name + "_s" + to_string(s), g.seqLen[s], g.seqTop[s], g.seqBot[s]));
        }
        genome->setDimensions(dims);
    }
    for (size_t gi = 0; gi < order.size(); gi++) {
        Node *n = order[gi].first;
        GenomeModel &g = models[n->name];
        Genome *genome = aln->openGenome(n->name);
        hal_size_t nc = aln->getChildNames(n->name).size();
        if (!g.botLen.empty()) {
            BottomSegmentIteratorPtr bi = genome->getBottomSegmentIterator();
            for (size_t i = 0; i < g.botLen.size(); i++) {
                bi->setCoordinates(g.botStart[i], g.botLen[i]);
                for (hal_size_t c = 0; c < nc; c++) { bi->bseg()->setChildIndex(c, NULL_INDEX); bi->bseg()->setChildReversed(c, false); }
                bi->bseg()->setTopParseIndex(NULL_INDEX);
                bi->toRight();
            }
        }
        if (!g.topLen.empty()) {
            TopSegmentIteratorPtr ti = genome->getTopSegmentIterator();
            for (size_t i = 0; i < g.topLen.size(); i++) {
                ti->setCoordinates(g.topStart[i], g.topLen[i]);
                ti->tseg()->setParentIndex(g.topParent[i] >= 0 ? g.topParent[i] : NULL_INDEX);
                ti->tseg()->setParentReversed(g.topRev[i]);
                ti->tseg()->setNextParalogyIndex(NULL_INDEX);
                ti->tseg()->setBottomParseIndex(NULL_INDEX);
                ti->toRight();
            }
        }
        genome->setString(g.dna);
    }
    /* parent->child links + paralogy rings */
    for (size_t gi = 0; gi < order.size(); gi++) {
        Node *n = order[gi].first, *pn = order[gi].second;
        if (pn == NULL) continue;
        GenomeModel &g = models[n->name];
        Genome *genome = aln->openGenome(n->name);
        Genome *parent = aln->openGenome(pn->name);
        vector<string> sibs = aln->getChildNames(pn->name);
        hal_size_t slot = 0;
        for (hal_size_t i = 0; i < sibs.size(); i++) if (sibs[i] == n->name) slot = i;
        map<long, vector<long> > byParent;
        for (size_t i = 0; i < g.topParent.size(); i++) if (g.topParent[i] >= 0) byParent[g.topParent[i]].push_back((long)i);
        for (map<long, vector<long> >::iterator it = byParent.begin(); it != byParent.end(); ++it) {
            vector<long> &m = it->second;
            long canon = m[rng() % m.size()];
            BottomSegmentIteratorPtr bi = parent->getBottomSegmentIterator(it->first);
            bi->bseg()->setChildIndex(slot, canon);
            bi->bseg()->setChildReversed(slot, g.topRev[canon]);
            if (m.size() > 1) {
                for (size_t k = 0; k < m.size(); k++) {
                    TopSegmentIteratorPtr ti = genome->getTopSegmentIterator(m[k]);
                    ti->tseg()->setNextParalogyIndex(m[(k + 1) % m.size()]);
                }
            }
        }
    }
    for (size_t gi = 0; gi < order.size(); gi++) {
        aln->openGenome(order[gi].first->name)->fixParseInfo();
    }
}

/* ---- randgen mode: the reference's own content generator over an explicit tree --------------- */
static void buildRandgen(AlignmentPtr aln, const Opts &o) {
    RandNumberGen rng(false, (int)o.seed);
    createRandomDimensions(rng, aln, o.minLen, o.maxLen, o.segs, o.segs);
    deque<string> q;
    q.push_front(aln->getRootName());
    while (!q.empty()) {
        Genome *genome = aln->openGenome(q.back());
        q.pop_back();
        createRandomGenome(rng, aln, genome);
        vector<string> kids = aln->getChildNames(genome->getName());
        for (size_t i = 0; i < kids.size(); i++) q.push_front(kids[i]);
    }
}

int main(int argc, char **argv) {
    Opts o;
    o.mode = "varlen"; o.newick = "((L0,L1)A0,(L2)A1)R;"; o.seed = 1; o.segs = 200; o.minLen = 5; o.maxLen = 40;
    o.seqs = 1; o.branch = -1; o.pInv = 0.2; o.pDup = 0.1; o.pIns = 0.05; o.pDel = 0.05; o.pMut = 0.1; o.fileGB = 1;
    vector<string> meta;
    for (int i = 1; i < argc; i++) {
        string a = argv[i];
        if (a.compare(0, 2, "--") == 0 && i + 1 < argc) {
            string v = argv[++i];
            if (a == "--mode") o.mode = v; else if (a == "--newick") o.newick = v; else if (a == "--seed") o.seed = atoi(v.c_str());
            else if (a == "--segs") o.segs = atol(v.c_str()); else if (a == "--minLen") o.minLen = atol(v.c_str());
            else if (a == "--maxLen") o.maxLen = atol(v.c_str()); else if (a == "--seqs") o.seqs = atol(v.c_str());
            else if (a == "--branch") o.branch = atof(v.c_str()); else if (a == "--pInv") o.pInv = atof(v.c_str());
            else if (a == "--pDup") o.pDup = atof(v.c_str()); else if (a == "--pIns") o.pIns = atof(v.c_str());
            else if (a == "--pDel") o.pDel = atof(v.c_str()); else if (a == "--pMut") o.pMut = atof(v.c_str());
            else if (a == "--fileGB") o.fileGB = atol(v.c_str());
            else if (a == "--meta") meta.push_back(v); /* genome:key=value, written through Genome::getMetaData()->set */
            else { cerr << "unknown option " << a << endl; return 1; }
        } else {
            o.out = a;
        }
    }
    if (o.out.empty()) { cerr << "usage: halTreeGen [options] out.hal" << endl; return 1; }
    try {
        const char *p = o.newick.c_str();
        Node *root = parseNewick(p);
        AlignmentPtr aln(mmapAlignmentInstance(o.out, CREATE_ACCESS, o.fileGB << 30));
        addTree(aln, root, NULL, o);
        if (o.mode == "randgen") buildRandgen(aln, o);
        else buildVarlen(aln, root, o);
        for (size_t i = 0; i < meta.size(); i++) {
            const size_t c = meta[i].find(':'), e = meta[i].find('=');
            Genome *g = aln->openGenome(meta[i].substr(0, c));
            if (g == NULL || c == string::npos || e == string::npos) throw hal_exception("bad --meta " + meta[i]);
            g->getMetaData()->set(meta[i].substr(c + 1, e - c - 1), meta[i].substr(e + 1));
        }
        aln->close();
    } catch (exception &e) {
        cerr << "halTreeGen: " << e.what() << endl;
        return 1;
    }
    return 0;
}
