#include "engine.hpp"
#include "columns_kernel.cuh"
#include "liftover_kernel.cuh"
#include "stage_kernels.cuh"
#include "wiggle_kernels.cuh"
#include "maf_kernels.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <fcntl.h>
#include <unistd.h>

namespace halgpu {
namespace rt {
unsigned long long g_launches = 0;
#if !defined(HALGPU_SIMT_EMUL)
namespace {
struct PinnedCache {
    std::mutex m;
    std::multimap<size_t, void *> idle;
    std::map<void *, size_t> sizeOf;
    size_t idleBytes = 0;
};
PinnedCache &pinned() {
    static PinnedCache *c = new PinnedCache; // leaked on purpose: outlives the CUDA context teardown order
    return *c;
}
} // namespace
void *hostAlloc(size_t n) {
    n = std::max<size_t>(n, 1);
    PinnedCache &c = pinned();
    {
        std::lock_guard<std::mutex> g(c.m);
        auto it = c.idle.lower_bound(n);
        if (it != c.idle.end() && it->first <= 2 * n + (1u << 20)) {
            void *p = it->second;
            c.idleBytes -= it->first;
            c.idle.erase(it);
            return p;
        }
    }
    const size_t cap = n + n / 8;
    void *p = nullptr;
    check(cudaMallocHost(&p, cap), "cudaMallocHost");
    std::lock_guard<std::mutex> g(c.m);
    c.sizeOf[p] = cap;
    return p;
}
void hostFree(void *p) {
    if (p == nullptr) return;
    PinnedCache &c = pinned();
    std::lock_guard<std::mutex> g(c.m);
    auto it = c.sizeOf.find(p);
    if (it == c.sizeOf.end()) { cudaFreeHost(p); return; }
    if (c.idleBytes + it->second > (8ull << 30)) { // keep at most 8 GB parked
        c.sizeOf.erase(it);
        cudaFreeHost(p);
        return;
    }
    c.idle.emplace(it->second, p);
    c.idleBytes += it->second;
}
#endif
}

namespace {
struct DevBuf { // RAII for per-batch device scratch (stream-ordered)
    void *p = nullptr;
    rt::Stream s;
    static rt::Stream &current() { static thread_local rt::Stream cur{}; return cur; }
    explicit DevBuf(size_t n) : p(rt::dmallocAsync(n, current())), s(current()) {}
    ~DevBuf() { rt::dfreeAsync(p, s); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <class T> T *as() const { return static_cast<T *>(p); }
    void *release() { void *q = p; p = nullptr; return q; }
};
// buffers of one batch, handed back to the context's cache when the batch ends
struct Lease {
    DeviceCache &cache;
    std::vector<void *> held;
    explicit Lease(DeviceCache &c) : cache(c) {}
    ~Lease() { for (void *p : held) cache.give(p); }
    Lease(const Lease &) = delete;
    Lease &operator=(const Lease &) = delete;
    void *take(size_t bytes) { void *p = cache.take(bytes); held.push_back(p); return p; }
    template <class T> T *as(size_t count) { return static_cast<T *>(take(count * sizeof(T))); }
    void giveNow(void *p) {
        for (size_t i = 0; i < held.size(); ++i) if (held[i] == p) { held.erase(held.begin() + (long)i); cache.give(p); return; }
    }
    void *detach(void *p) { // ownership leaves with the result
        for (size_t i = 0; i < held.size(); ++i) if (held[i] == p) { held.erase(held.begin() + (long)i); return p; }
        return p;
    }
};
unsigned gridFor(int64_t n, unsigned block, int sms) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)sms * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}
} // namespace

DeviceCache::~DeviceCache() {
    for (auto &kv : _sizeOf) rt::dfree(kv.first);
}
void *DeviceCache::take(size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    {
        std::lock_guard<std::mutex> g(_m);
        auto it = _idle.lower_bound(bytes);
        if (it != _idle.end() && it->first <= 2 * bytes + (1u << 20)) {
            void *p = it->second;
            _idle.erase(it);
            return p;
        }
    }
    const size_t cap = (bytes + bytes / 8 + 255) & ~(size_t)255;
    void *p = rt::dmalloc(cap);
    std::lock_guard<std::mutex> g(_m);
    _sizeOf[p] = cap;
    _held += cap;
    return p;
}
void DeviceCache::give(void *p) {
    if (p == nullptr) return;
    std::lock_guard<std::mutex> g(_m);
    auto it = _sizeOf.find(p);
    if (it == _sizeOf.end()) { rt::dfree(p); return; } // not ours (defensive)
    _idle.emplace(it->second, p);
}

void *Context::alloc(size_t bytes) {
    void *p = rt::dmalloc(bytes);
    _owned.push_back(p);
    _staged += bytes;
    return p;
}

// ---- staging: file -> pinned ring (several reader threads) -> HBM -> re-pack kernels -------------------------------------
// The arrays are read with pread() straight into page-locked chunks by a few threads (no page faults on a mapping, and one
// thread cannot keep PCIe busy), each chunk travels with an asynchronous copy while the next one is being read, and the
// re-pack kernels run behind the copies on the same stream.  Genomes are staged on first use (a liftover touches the genomes
// on its path only; the column sweeps and blockViz ask for all of them).
struct Context::Stager {
    static constexpr size_t CHUNK = 32u << 20;
    int fd = -1;
    uint8_t *pin[2] = {nullptr, nullptr};
    std::unique_ptr<rt::Event> done[2];
    bool used[2] = {false, false};
    unsigned turn = 0, threads = 4;
    explicit Stager(const std::string &path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw HalError(path + ": can't open HAL file for staging");
        for (int i = 0; i < 2; ++i) { pin[i] = static_cast<uint8_t *>(rt::hostAlloc(CHUNK)); done[i].reset(new rt::Event); }
        const unsigned hw = std::thread::hardware_concurrency();
        threads = std::max(1u, std::min(hw ? hw : 1u, 6u));
    }
    ~Stager() {
        if (fd >= 0) ::close(fd);
        for (int i = 0; i < 2; ++i) rt::hostFree(pin[i]);
    }
    static void readAll(int fd, uint8_t *dst, uint64_t off, size_t n) {
        while (n > 0) {
            const ssize_t r = ::pread(fd, dst, n, (off_t)off);
            if (r <= 0) throw HalError("short read while staging the HAL file");
            dst += r; off += (uint64_t)r; n -= (size_t)r;
        }
    }
    void copy(void *dst, uint64_t fileOff, size_t bytes, rt::Stream s) {
        for (size_t at = 0; at < bytes; at += CHUNK) {
            const size_t n = std::min(CHUNK, bytes - at);
            const unsigned b = turn++ & 1u;
            if (used[b]) done[b]->hostWait(); // the copy that last used this chunk has left it
            const unsigned T = n >= (4u << 20) ? threads : 1u;
            if (T == 1) {
                readAll(fd, pin[b], fileOff + at, n);
            } else {
                std::vector<std::thread> th;
                std::string failure;
                std::mutex fm;
                for (unsigned t = 0; t < T; ++t) {
                    const size_t lo = n * t / T, hi = n * (t + 1) / T;
                    th.emplace_back([&, lo, hi] {
                        try { readAll(fd, pin[b] + lo, fileOff + at + lo, hi - lo); }
                        catch (const std::exception &e) { std::lock_guard<std::mutex> g(fm); failure = e.what(); }
                    });
                }
                for (auto &x : th) x.join();
                if (!failure.empty()) throw HalError(failure);
            }
            rt::h2d(static_cast<uint8_t *>(dst) + at, pin[b], n, s);
            done[b]->record(s);
            used[b] = true;
        }
    }
};

Context::Context(const std::string &path, int device) : _file(new HalFile(path)), _device(device) {
    rt::setDevice(device);
    rt::retainPool(device);
    _stream = rt::createStream();
    _copy = rt::createStream();
    _copyBack = rt::createStream();
    _aux = rt::createStream();
    _sms = rt::smCount();
    _g.resize(_file->genomes().size());
    try {
        _stager.reset(new Stager(path));
        _hostCtr = static_cast<unsigned long long *>(rt::hostAlloc(32 * sizeof(unsigned long long)));
        for (auto &e : _ev) e.reset(new rt::Event);
        for (auto &e : _sliceEv) e.reset(new rt::Event);
        if (std::getenv("HALGPU_EAGER_STAGE") != nullptr) ensureAll(true);
    } catch (...) {
        for (void *p : _owned) rt::dfree(p);
        rt::destroyStream(_stream);
        rt::destroyStream(_copy);
        rt::destroyStream(_copyBack);
        rt::destroyStream(_aux);
        throw;
    }
}

Context::~Context() {
    try { rt::sync(_stream); } catch (...) {}
    rt::hostFree(_hostCtr);
    for (auto &kv : _plans) rt::dfree(kv.second.dSteps);
    for (void *p : _owned) rt::dfree(p);
    rt::destroyStream(_stream);
    rt::destroyStream(_copy);
    rt::destroyStream(_copyBack);
    rt::destroyStream(_aux);
}

void Context::buildBucket(const void *arr, bool isTop, int64_t N, int64_t len, uint32_t *&table, int &shift, int64_t &nb) {
    // bucket width ~ mean segment length, so a lookup lands within a segment or two of the answer
    shift = 0;
    while (shift < 40 && (len >> (shift + 1)) >= N) ++shift;
    nb = (len >> shift) + 1;
    table = static_cast<uint32_t *>(alloc((size_t)nb * sizeof(uint32_t)));
    BucketParams bp;
    bp.arr = arr; bp.bucket = table; bp.N = N; bp.numBuckets = nb; bp.isTop = isTop ? 1 : 0; bp.shift = shift;
    rt::launch(bucketKernel, gridFor(nb, 256, _sms), 256, 0, _stream, bp);
}

void Context::ensureGenome(int gi) {
    GenomeDev &d = _g[gi];
    if (d.staged) return;
    const GenomeInfo &g = _file->genomes()[gi];
    if (g.numTop >= (int64_t)0xffffffffll || g.numBottom >= (int64_t)0xffffffffll) {
        throw HalError("genome " + g.name + " has more than 2^32 segments; not supported by the bucket index");
    }
    const int nc = (int)g.children.size();
    Lease L(_cache);
    // top records
    {
        const size_t rawBytes = (size_t)(g.numTop + 1) * 40;
        uint8_t *raw = L.as<uint8_t>(rawBytes);
        _stager->copy(raw, _file->offsetOf(g.top), rawBytes, _stream);
        d.top = static_cast<TopRec *>(alloc((size_t)(g.numTop + 1) * sizeof(TopRec)));
        PackTopParams pp;
        pp.raw = raw; pp.out = d.top; pp.n = g.numTop + 1;
        rt::launch(packTopKernel, gridFor(pp.n, 256, _sms), 256, 0, _stream, pp);
    }
    // bottom records
    {
        const size_t rawBytes = (size_t)(g.numBottom + 1) * g.bottomStride;
        uint8_t *raw = L.as<uint8_t>(rawBytes);
        _stager->copy(raw, _file->offsetOf(g.bottom), rawBytes, _stream);
        d.bot = static_cast<BotCore *>(alloc((size_t)(g.numBottom + 1) * sizeof(BotCore)));
        d.child = static_cast<int64_t *>(alloc(std::max<size_t>(8, (size_t)nc * (size_t)g.numBottom * sizeof(int64_t))));
        PackBotParams pp;
        pp.raw = raw; pp.core = d.bot; pp.child = d.child;
        pp.n = g.numBottom + 1; pp.numBot = g.numBottom; pp.nc = nc; pp.stride = (int32_t)g.bottomStride;
        rt::launch(packBotKernel, gridFor(pp.n, 256, _sms), 256, 0, _stream, pp);
    }
    // sequence start table (+ sentinel) and the child genome ids
    std::vector<int64_t> ss;
    for (const SequenceInfo &q : g.sequences) ss.push_back(q.start);
    ss.push_back(g.length);
    d.seqStart = static_cast<int64_t *>(alloc(ss.size() * sizeof(int64_t)));
    rt::h2d(d.seqStart, ss.data(), ss.size() * sizeof(int64_t), _stream);
    std::vector<int32_t> kids(g.children.begin(), g.children.end());
    kids.push_back(-1);
    d.childGenome = static_cast<int32_t *>(alloc(kids.size() * sizeof(int32_t)));
    rt::h2d(d.childGenome, kids.data(), kids.size() * sizeof(int32_t), _stream);
    if (g.numTop > 0) buildBucket(d.top, true, g.numTop, g.length, d.topBucket, d.topShift, d.topBuckets);
    if (g.numBottom > 0) buildBucket(d.bot, false, g.numBottom, g.length, d.botBucket, d.botShift, d.botBuckets);
    rt::sync(_stream); // the raw copies return to the cache, ss / kids leave scope
    d.staged = true;
}

void Context::ensureDna(int gi) {
    GenomeDev &d = _g[gi];
    if (d.dna != nullptr) return;
    const GenomeInfo &g = _file->genomes()[gi];
    const size_t bytes = (size_t)((g.length + 1) / 2);
    d.dna = static_cast<uint8_t *>(alloc(std::max<size_t>(bytes, 1)));
    if (bytes > 0) _stager->copy(d.dna, _file->offsetOf(g.dna), bytes, _stream);
    rt::sync(_stream);
}

void Context::ensureAll(bool dna) {
    for (size_t g = 0; g < _g.size(); ++g) ensureGenome((int)g);
    if (dna)
        for (size_t g = 0; g < _g.size(); ++g) ensureDna((int)g);
}

// the run and xlate fields of a vertical link array (device_index.cuh)
void Context::linkFields(int64_t *links, int64_t linkStride, const int64_t *starts, int64_t startStride, int64_t n, const TopRec *landTop,
                         const int64_t *otherStarts, int64_t otherStride, int64_t *xlateOut) {
    Lease L(_cache);
    uint32_t *mark = L.as<uint32_t>((size_t)n), *next = L.as<uint32_t>((size_t)n);
    size_t tmpBytes = 0;
    rt::suffixMinU32Tmp(nullptr, tmpBytes, mark, next, (size_t)n, _stream);
    void *tmp = L.take(tmpBytes);
    LinkRunParams lp;
    lp.links = links; lp.starts = starts; lp.linkStride = linkStride; lp.startStride = startStride; lp.n = n;
    lp.landTop = landTop; lp.mark = mark;
    rt::launch(linkBreakKernel, gridFor(n, 256, _sms), 256, 0, _stream, lp);
    rt::suffixMinU32Tmp(tmp, tmpBytes, mark, next, (size_t)n, _stream);
    lp.mark = next;
    rt::launch(linkRunKernel, gridFor(n, 256, _sms), 256, 0, _stream, lp);
    LinkXlateParams xp;
    xp.links = links; xp.starts = starts; xp.otherStarts = otherStarts; xp.linkStride = linkStride; xp.startStride = startStride;
    xp.otherStride = otherStride; xp.n = n; xp.xlate = xlateOut;
    rt::launch(linkXlateKernel, gridFor(n, 256, _sms), 256, 0, _stream, xp);
    rt::sync(_stream);
}

void Context::ensureUpLinks(int gi) { // parent links of genome gi's tops
    const GenomeInfo &g = _file->genomes()[gi];
    GenomeDev &d = _g[gi];
    if (d.topX != nullptr || g.parent < 0 || g.numTop == 0) return;
    ensureGenome(gi);
    ensureGenome(g.parent);
    const int64_t topStride = sizeof(TopRec) / 8, botStride = sizeof(BotCore) / 8;
    int64_t *x = static_cast<int64_t *>(alloc((size_t)g.numTop * sizeof(int64_t)));
    linkFields(&d.top[0].parentEnc, topStride, &d.top[0].start, topStride, g.numTop, nullptr, &_g[g.parent].bot[0].start, botStride, x);
    d.topX = x;
    d.topFast = static_cast<FastRec *>(alloc((size_t)d.topBuckets * sizeof(FastRec)));
    FastIndexParams fp;
    fp.bucket = d.topBucket; fp.links = &d.top[0].parentEnc; fp.starts = &d.top[0].start; fp.xlate = x; fp.linkStride = topStride;
    fp.startStride = topStride; fp.numBuckets = d.topBuckets; fp.out = d.topFast;
    rt::launch(fastIndexKernel, gridFor(d.topBuckets, 256, _sms), 256, 0, _stream, fp);
    rt::sync(_stream);
}

void Context::ensureDownLinks(int gi, int slot) { // child links of genome gi's bottoms, one child slot
    const GenomeInfo &g = _file->genomes()[gi];
    GenomeDev &d = _g[gi];
    ensureGenome(gi);
    if (g.numBottom == 0 || slot < 0 || slot >= (int)g.children.size()) return;
    if (d.childX == nullptr) {
        d.childX = static_cast<int64_t *>(alloc(g.children.size() * (size_t)g.numBottom * sizeof(int64_t)));
        d.childLinked.assign(g.children.size(), 0);
        d.childFast.assign(g.children.size(), nullptr);
    }
    if (d.childLinked[(size_t)slot]) return;
    const int c = g.children[(size_t)slot];
    ensureGenome(c);
    const int64_t topStride = sizeof(TopRec) / 8, botStride = sizeof(BotCore) / 8;
    int64_t *col = d.child + (size_t)slot * (size_t)g.numBottom;
    int64_t *xcol = d.childX + (size_t)slot * (size_t)g.numBottom;
    linkFields(col, 1, &d.bot[0].start, botStride, g.numBottom, _g[c].top, &_g[c].top[0].start, topStride, xcol);
    d.childFast[(size_t)slot] = static_cast<FastRec *>(alloc((size_t)d.botBuckets * sizeof(FastRec)));
    FastIndexParams fp;
    fp.bucket = d.botBucket; fp.links = col; fp.starts = &d.bot[0].start; fp.xlate = xcol; fp.linkStride = 1; fp.startStride = botStride;
    fp.numBuckets = d.botBuckets; fp.out = d.childFast[(size_t)slot];
    rt::launch(fastIndexKernel, gridFor(d.botBuckets, 256, _sms), 256, 0, _stream, fp);
    rt::sync(_stream);
    d.childLinked[(size_t)slot] = 1;
}

const Plan &Context::plan(int src, int tgt, int coal) {
    const auto &G = _file->genomes();
    Plan p;
    p.src = src; p.tgt = tgt; p.mrca = _file->mrca(src, tgt);
    if (p.mrca < 0) throw HalError("source and target genomes share no ancestor");
    if (coal == p.mrca) coal = -1; // the default (liftover/impl/halBlockLiftover.cpp:36-38)
    auto key = std::make_tuple(src, tgt, coal);
    auto it = _plans.find(key);
    if (it != _plans.end()) return it->second;
    // genomes whose paralogy rings mapRecursiveParalogies walks: the MRCA ... the child of the limit
    // (api/impl/halSegmentMapper.cpp:525-576); the limit must be an ancestor of the MRCA
    std::vector<int> para;
    if (coal >= 0) {
        int g = p.mrca;
        while (g >= 0 && g != coal) { para.push_back(g); g = G[g].parent; }
        if (g < 0) throw HalError("Hit root genome when attempting to map paralogies");
    }
    p.coal = coal;
    std::vector<int> up, down;
    for (int g = src; g != p.mrca; g = G[g].parent) up.push_back(g);
    for (int g = tgt; g != p.mrca; g = G[g].parent) down.push_back(g);
    std::reverse(down.begin(), down.end());
    p.upSteps = (int)up.size();
    // path positions: [up genomes] [para levels 0..K, upward] [para levels K..1, downward] [MRCA] [down genomes]
    struct Entry { int g; int up; int flags; int jump; };
    std::vector<Entry> ent;
    for (int g : up) ent.push_back(Entry{g, 1, 0, 0});
    const int K = (int)para.size() - 1;
    const int firstPara = (int)ent.size();
    const int mrcaPos = firstPara + (K >= 0 ? 2 * K + 1 : 0); // position of the MRCA entry the downward part starts from
    for (int k = 0; k <= K; ++k) {
        const int backDown = k == 0 ? mrcaPos : firstPara + (K + 1) + (K - k); // where level k's paralogs go down from
        ent.push_back(Entry{para[k], 1, STEP_PARA | (k == K ? STEP_PARA_LAST : 0), backDown});
    }
    for (int k = K; k >= 1; --k) ent.push_back(Entry{para[k], 0, STEP_NODUPES, 0});
    ent.push_back(Entry{p.mrca, 0, 0, 0});
    for (int g : down) ent.push_back(Entry{g, 0, 0, 0});
    p.path.clear();
    for (const Entry &e : ent) p.path.push_back(e.g);
    for (const Entry &e : ent) ensureGenome(e.g); // stage what this path touches (first use of a genome)
    for (size_t i = 0; i + 1 < ent.size(); ++i) {
        if (ent[i].up) { if (G[ent[i].g].parent == ent[i + 1].g) ensureUpLinks(ent[i].g); }
        else ensureDownLinks(ent[i].g, G[ent[i + 1].g].slotInParent);
    }
    std::vector<PathStep> steps(ent.size());
    for (size_t i = 0; i < ent.size(); ++i) {
        const int g = ent[i].g;
        PathStep &s = steps[i];
        s.top = _g[g].top; s.bot = _g[g].bot; s.child = nullptr;
        s.numTop = G[g].numTop; s.numBot = G[g].numBottom;
        s.up = ent[i].up; s.flags = ent[i].flags; s.jump = ent[i].jump; s.pad = 0;
        s.topBucket = _g[g].topBucket; s.botBucket = _g[g].botBucket; s.topShift = _g[g].topShift; s.botShift = _g[g].botShift;
        if (i + 1 >= ent.size()) continue;
        const int nx = ent[i + 1].g;
        s.xlate = nullptr; s.fast = nullptr;
        if (!s.up) { // this genome's childEnc column for the slot of the next genome down
            s.child = _g[g].child + (size_t)G[nx].slotInParent * (size_t)G[g].numBottom;
            s.xlate = _g[g].childX ? _g[g].childX + (size_t)G[nx].slotInParent * (size_t)G[g].numBottom : nullptr;
            s.fast = _g[g].childFast.empty() ? nullptr : _g[g].childFast[(size_t)G[nx].slotInParent];
        } else if (G[g].parent >= 0 && G[g].parent == nx) { // the parent's column for this genome's slot: canonical-paralog test
            s.child = _g[nx].child + (size_t)G[g].slotInParent * (size_t)G[nx].numBottom;
            s.xlate = _g[g].topX;
            s.fast = _g[g].topFast;
        }
    }
    p.fastOk = coal < 0;
    for (size_t i = 0; i + 1 < steps.size(); ++i) p.fastOk = p.fastOk && steps[i].fast != nullptr;
    p.dSteps = static_cast<PathStep *>(rt::dmalloc(steps.size() * sizeof(PathStep)));
    rt::h2d(p.dSteps, steps.data(), steps.size() * sizeof(PathStep), _stream);
    rt::sync(_stream);
    return _plans.emplace(key, p).first->second;
}

void Context::buildGenomeTab(int ref, const std::vector<int> &targets, std::vector<GenomeTab> &tab) {
    const auto &G = _file->genomes();
    const int ng = (int)G.size();
    // scope = spanning tree of targets + reference (api/impl/halColumnIterator.cpp:47-51); rows reported for targets only
    std::vector<char> inScope(ng, targets.empty() ? 1 : 0), isTarget(ng, targets.empty() ? 1 : 0);
    if (!targets.empty()) {
        std::vector<int> all(targets);
        all.push_back(ref);
        int m = all[0];
        for (int t : all) {
            if (t < 0 || t >= ng) throw HalError("target genome index out of range");
            isTarget[t] = 1;
            m = _file->mrca(m, t);
        }
        for (int t : all)
            for (int g = t;; g = G[g].parent) { inScope[g] = 1; if (g == m) break; }
    }
    std::vector<int> order(ng);
    for (int g = 0; g < ng; ++g) order[g] = g;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return G[a].name < G[b].name; });
    std::vector<int> rank(ng);
    for (int i = 0; i < ng; ++i) rank[order[i]] = i;
    for (int g = 0; g < ng; ++g) {
        if (!inScope[g]) continue;
        ensureGenome(g); // (the walk tests a genome's scope flag before it touches its arrays)
    }
    tab.resize(ng);
    for (int g = 0; g < ng; ++g) {
        GenomeTab &t = tab[g];
        std::memset(&t, 0, sizeof(t));
        t.top = _g[g].top; t.bot = _g[g].bot; t.child = _g[g].child; t.childGenome = _g[g].childGenome;
        t.topBucket = _g[g].topBucket; t.botBucket = _g[g].botBucket;
        t.numTop = G[g].numTop; t.numBot = G[g].numBottom;
        t.nc = (int32_t)G[g].children.size(); t.parent = G[g].parent; t.slot = G[g].slotInParent;
        t.topShift = _g[g].topShift; t.botShift = _g[g].botShift;
        t.inScope = (uint8_t)inScope[g]; t.isTarget = (uint8_t)isTarget[g];
        t.seqStart = _g[g].seqStart; t.numSeq = (int32_t)G[g].sequences.size(); t.nameRank = rank[g];
    }
}

// reference segments (top array if the genome has one, else bottom) holding the positions first..last: binary search
// over the start fields of the mapped file
static void refSegRange(const GenomeInfo &g, int64_t first, int64_t last, int64_t &seg0, int64_t &nSegs) {
    const bool top = g.numTop > 0;
    const int64_t N = top ? g.numTop : g.numBottom;
    const uint8_t *base = top ? g.top : g.bottom;
    const size_t stride = top ? 40 : g.bottomStride;
    if (N <= 0) throw HalError("genome " + g.name + " has no segments");
    auto startOf = [&](int64_t i) { int64_t v; std::memcpy(&v, base + (size_t)i * stride, 8); return v; };
    auto find = [&](int64_t pos) { // largest i with start(i) <= pos
        int64_t lo = 0, hi = N;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (startOf(mid) <= pos) lo = mid; else hi = mid; }
        return lo;
    };
    seg0 = find(first);
    nSegs = find(last) - seg0 + 1;
}

void Context::columnRuns(int ref, int64_t first, int64_t last, const std::vector<int> &targets, uint32_t flags, halgpu_col_runs &out,
                         int64_t windowFirst) {
    const auto &G = _file->genomes();
    const int ng = (int)G.size();
    if (ref < 0 || ref >= ng) throw HalError("genome index out of range");
    if (ng > 32767) throw HalError("too many genomes for the column row record");
    for (const GenomeInfo &g : G) { // WalkStack packs the child slot into 12 bits (columns_kernel.cuh)
        if (g.children.size() > 4095) throw HalError("genome " + g.name + " has more than 4095 children; not supported by the column walk");
    }
    if (first < 0 || last < first || last >= G[ref].length) throw HalError("column range out of bounds for genome " + G[ref].name);
    const bool unique = (flags & COL_UNIQUE) != 0;
    if (windowFirst < 0) windowFirst = first;
    if (unique && windowFirst > first) throw HalError("the sweep start of a unique column range lies right of the range");
    std::vector<GenomeTab> tab;
    buildGenomeTab(ref, targets, tab);
    const int64_t n = last - first + 1;
    int64_t seg0 = 0, nSegs = 0;
    refSegRange(G[ref], first, last, seg0, nSegs);
    Lease L(_cache);
    try {
        GenomeTab *dTab = L.as<GenomeTab>(tab.size());
        uint32_t *dErr = L.as<uint32_t>(1);
        rt::h2d(dTab, tab.data(), tab.size() * sizeof(GenomeTab), _stream);
        rt::dmemset(dErr, 0, sizeof(uint32_t), _stream);
        uint32_t *segPieces = L.as<uint32_t>((size_t)nSegs + 1), *segRows = L.as<uint32_t>((size_t)nSegs + 1);
        uint64_t *pieceOff = L.as<uint64_t>((size_t)nSegs + 2), *rowOff = L.as<uint64_t>((size_t)nSegs + 2);
        ColPieceParams cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.genomes = dTab; cp.ref = ref; cp.flags = flags; cp.first = first; cp.n = n; cp.seg0 = seg0; cp.nSegs = nSegs;
        cp.window = windowFirst; cp.segPieces = segPieces; cp.segRows = segRows; cp.error = dErr;
        _ev[0]->record(_stream);
        rt::launch(colPieceKernel<false>, gridFor(nSegs, 128, _sms), 128, 0, _stream, cp);
        rt::exclusiveScanU32(segPieces, pieceOff, (size_t)nSegs, _stream);
        rt::exclusiveScanU32(segRows, rowOff, (size_t)nSegs, _stream);
        uint64_t totals[2] = {0, 0};
        uint32_t err = 0;
        rt::d2h(&totals[0], pieceOff + nSegs, 8, _stream);
        rt::d2h(&totals[1], rowOff + nSegs, 8, _stream);
        rt::d2h(&err, dErr, sizeof(err), _stream);
        rt::sync(_stream);
        if (err) throw HalError("column walk exceeded its stack or " + std::to_string(HG_MAX_ROWS) + " rows per column");
        const uint64_t nPieces = totals[0], nPieceRows = totals[1];
        int64_t *pieceCol = L.as<int64_t>(nPieces + 1);
        uint64_t *pieceRowOff = L.as<uint64_t>(nPieces + 1);
        uint32_t *pieceRows = L.as<uint32_t>(nPieces + 1);
        uint8_t *pieceClass = L.as<uint8_t>(nPieces + 1);
        ColRowRec *rows = L.as<ColRowRec>(std::max<uint64_t>(nPieceRows, 1));
        cp.pieceOff = pieceOff; cp.rowOff = rowOff; cp.pieceCol = pieceCol; cp.pieceRowOff = pieceRowOff; cp.pieceRows = pieceRows;
        cp.pieceClass = pieceClass; cp.rows = rows;
        rt::launch(colPieceKernel<true>, gridFor(nSegs, 128, _sms), 128, 0, _stream, cp);
        // join pieces that continue each other
        uint32_t *isStart = L.as<uint32_t>(nPieces + 1), *startRows = L.as<uint32_t>(nPieces + 1);
        uint64_t *runIndex = L.as<uint64_t>(nPieces + 2), *runRowOffset = L.as<uint64_t>(nPieces + 2);
        PieceMergeParams mp;
        mp.pieceCol = pieceCol; mp.pieceRowOff = pieceRowOff; mp.pieceRows = pieceRows; mp.pieceClass = pieceClass; mp.rows = rows;
        mp.isStart = isStart; mp.startRows = startRows; mp.nPieces = (int64_t)nPieces;
        rt::launch(pieceMergeKernel, gridFor((int64_t)nPieces + 1, 256, _sms), 256, 0, _stream, mp);
        rt::exclusiveScanU32(isStart, runIndex, (size_t)nPieces, _stream);
        rt::exclusiveScanU32(startRows, runRowOffset, (size_t)nPieces, _stream);
        rt::d2h(&totals[0], runIndex + nPieces, 8, _stream);
        rt::d2h(&totals[1], runRowOffset + nPieces, 8, _stream);
        rt::sync(_stream);
        const uint64_t nRuns = totals[0], nRows = totals[1];
        int64_t *runCol = L.as<int64_t>(nRuns + 1);
        uint64_t *runRowOff = L.as<uint64_t>(nRuns + 1);
        uint8_t *runClass = L.as<uint8_t>(nRuns + 1);
        ColRowRec *runRows = L.as<ColRowRec>(std::max<uint64_t>(nRows, 1));
        RunScatterParams rp;
        rp.isStart = isStart; rp.runIndex = runIndex; rp.runRowOffset = runRowOffset; rp.pieceCol = pieceCol; rp.pieceRowOff = pieceRowOff;
        rp.pieceRows = pieceRows; rp.pieceClass = pieceClass; rp.rows = rows; rp.runCol = runCol; rp.runRowOff = runRowOff;
        rp.runClass = unique ? runClass : nullptr; rp.runRows = runRows; rp.nPieces = (int64_t)nPieces; rp.nCols = n;
        rt::launch(runScatterKernel, gridFor((int64_t)nPieces + 1, 256, _sms), 256, 0, _stream, rp);
        _ev[1]->record(_stream);
        out.n_cols = (size_t)n; out.n_runs = (size_t)nRuns; out.n_rows = (size_t)nRows;
        out.run_col = static_cast<int64_t *>(rt::hostAlloc((nRuns + 1) * 8));
        out.row_offset = static_cast<uint64_t *>(rt::hostAlloc((nRuns + 1) * 8));
        out.rows = static_cast<halgpu_col_row *>(rt::hostAlloc(std::max<uint64_t>(nRows, 1) * sizeof(halgpu_col_row)));
        rt::d2h(out.run_col, runCol, (nRuns + 1) * 8, _stream);
        rt::d2h(out.row_offset, runRowOff, (nRuns + 1) * 8, _stream);
        rt::d2h(out.rows, runRows, nRows * sizeof(halgpu_col_row), _stream);
        out.run_class = nullptr;
        if (unique) {
            out.run_class = static_cast<uint8_t *>(rt::hostAlloc(std::max<uint64_t>(nRuns, 1)));
            rt::d2h(out.run_class, runClass, nRuns, _stream);
        }
        rt::sync(_stream);
        out.kernel_ms = rt::Event::elapsedMs(*_ev[0], *_ev[1]); // both walks, the scans and the merge
    } catch (...) {
        try { rt::sync(_stream); } catch (...) {}
        throw;
    }
}

void Context::depth(int ref, int64_t first, int64_t last, int64_t step, const std::vector<int> &targets, uint32_t flags,
                    int32_t *dOut, float *kernelMs) {
    const auto &G = _file->genomes();
    const int ng = (int)G.size();
    if (ref < 0 || ref >= ng) throw HalError("genome index out of range");
    if (ng > 256) throw HalError("alignment depth supports at most 256 genomes");
    for (const GenomeInfo &g : G) {
        if (g.children.size() > 4095) throw HalError("genome " + g.name + " has more than 4095 children; not supported by the column walk");
    }
    if (step < 1 || first < 0 || last < first || last >= G[ref].length) throw HalError("column range out of bounds for genome " + G[ref].name);
    std::vector<GenomeTab> tab;
    buildGenomeTab(ref, targets, tab);
    Lease L(_cache);
    try {
        GenomeTab *dTab = L.as<GenomeTab>(tab.size());
        uint32_t *dErr = L.as<uint32_t>(1);
        rt::h2d(dTab, tab.data(), tab.size() * sizeof(GenomeTab), _stream);
        rt::dmemset(dErr, 0, sizeof(uint32_t), _stream);
        DepthParams P;
        P.genomes = dTab; P.numGenomes = ng; P.ref = ref;
        P.first = first; P.step = step; P.n = (last - first) / step + 1;
        refSegRange(G[ref], first, first + (P.n - 1) * step, P.seg0, P.nSegs);
        P.flags = flags; P.depth = dOut; P.error = dErr;
        _ev[0]->record(_stream);
        rt::launch(depthKernel, gridFor(P.nSegs, 128, _sms), 128, 0, _stream, P);
        _ev[1]->record(_stream);
        uint32_t err = 0;
        rt::d2h(&err, dErr, sizeof(err), _stream);
        rt::sync(_stream); // (tab is pageable host memory: the copy above has completed by now)
        if (kernelMs) *kernelMs = rt::Event::elapsedMs(*_ev[0], *_ev[1]);
        if (err) throw HalError("column walk exceeded its stack (more than " + std::to_string(HG_WALK_STACK) + " pending branches)");
    } catch (...) {
        try { rt::sync(_stream); } catch (...) {}
        throw;
    }
}

void Context::mafText(size_t nRows, const halgpu_maf_row *rows, size_t nPieces, const halgpu_maf_piece *pieces, const char *prefix,
                      size_t prefixBytes, size_t outBytes, char *out, float *kernelMs) {
    const auto &G = _file->genomes();
    std::vector<char> needDna(G.size(), 0);
    for (size_t r = 0; r < nRows; ++r) { // nothing the kernel derives an address from goes unchecked
        const halgpu_maf_row &w = rows[r];
        if (w.genome < 0 || w.genome >= (int)G.size()) throw HalError("MAF row " + std::to_string(r) + ": genome index out of range");
        if ((uint64_t)w.first_piece + w.num_pieces > nPieces || (uint64_t)w.prefix_offset + w.prefix_len > prefixBytes) {
            throw HalError("MAF row " + std::to_string(r) + " refers to pieces / prefix bytes outside the arrays");
        }
        uint64_t bytes = (uint64_t)w.prefix_len + w.tail_newlines;
        for (uint32_t k = 0; k < w.num_pieces; ++k) {
            const halgpu_maf_piece &pc = pieces[w.first_piece + k];
            const int64_t count = pc.count_kind >> 2;
            const int kind = (int)(pc.count_kind & 3);
            if (count < 0 || kind > 2) throw HalError("MAF row " + std::to_string(r) + ": malformed piece");
            if (kind != 0 && count > 0) {
                const int64_t lo = kind == 1 ? pc.pos : pc.pos - count + 1, hi = kind == 1 ? pc.pos + count - 1 : pc.pos;
                if (lo < 0 || hi >= G[(size_t)w.genome].length) throw HalError("MAF row " + std::to_string(r) + ": bases outside genome " + G[(size_t)w.genome].name);
                needDna[(size_t)w.genome] = 1;
            }
            bytes += (uint64_t)count;
        }
        if (w.tail_newlines > 2 || w.out_offset > outBytes || bytes > outBytes - w.out_offset) throw HalError("MAF row " + std::to_string(r) + " does not fit the output");
    }
    std::vector<const uint8_t *> dnaTab(G.size(), nullptr);
    for (size_t g = 0; g < G.size(); ++g)
        if (needDna[g]) { ensureDna((int)g); dnaTab[g] = _g[g].dna; }
    Lease L(_cache);
    try {
        halgpu_maf_row *dRows = L.as<halgpu_maf_row>(std::max<size_t>(nRows, 1));
        halgpu_maf_piece *dPieces = L.as<halgpu_maf_piece>(std::max<size_t>(nPieces, 1));
        char *dPrefix = L.as<char>(std::max<size_t>(prefixBytes, 1));
        const uint8_t **dDna = L.as<const uint8_t *>(G.size());
        char *dOut = L.as<char>(std::max<size_t>(outBytes, 1));
        rt::h2d(dRows, rows, nRows * sizeof(halgpu_maf_row), _stream);
        rt::h2d(dPieces, pieces, nPieces * sizeof(halgpu_maf_piece), _stream);
        rt::h2d(dPrefix, prefix, prefixBytes, _stream);
        rt::h2d(dDna, dnaTab.data(), G.size() * sizeof(const uint8_t *), _stream);
        MafTextParams P;
        P.rows = dRows; P.pieces = dPieces; P.prefix = dPrefix; P.dna = dDna; P.out = dOut; P.nRows = (int64_t)nRows;
        _ev[0]->record(_stream);
        if (nRows > 0) rt::launch(mafTextKernel, gridFor((int64_t)nRows * 32, 256, _sms), 256, 0, _stream, P);
        _ev[1]->record(_stream);
        rt::d2h(out, dOut, outBytes, _stream);
        rt::sync(_stream);
        if (kernelMs) *kernelMs = rt::Event::elapsedMs(*_ev[0], *_ev[1]);
    } catch (...) {
        try { rt::sync(_stream); } catch (...) {}
        throw;
    }
}

namespace {
struct PhaseTimer { // HALGPU_TIMING=1: host wall-clock of each phase of a batch (diagnostics only)
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string log;
    PhaseTimer() : on(std::getenv("HALGPU_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        auto t = std::chrono::steady_clock::now();
        log += std::string(what) + "=" + std::to_string(std::chrono::duration<double, std::milli>(t - t0).count()) + "ms ";
        t0 = t;
    }
    ~PhaseTimer() { if (on) fprintf(stderr, "[halgpu timing] %s\n", log.c_str()); }
};
} // namespace

size_t Context::poolBytesFor(int src, int tgt, size_t n, int coal) {
    const auto &G = _file->genomes();
    if (src < 0 || tgt < 0 || src >= (int)G.size() || tgt >= (int)G.size()) throw HalError("genome index out of range");
    const Plan &pl = plan(src, tgt, coal);
    return ((size_t)n + (size_t)((double)n * std::max(1.25, pl.linesPerInterval * 1.125)) + 4096) * sizeof(halgpu_lift_rec); // (liftover(): poolCap)
}

void Context::liftover(int src, int tgt, uint32_t flags, size_t n, const int64_t *dGs, const int64_t *dGe,
                       const uint8_t *dStrand, LiftOutput &out, uint64_t offsetBase, const WigScatter *wig, int coal, const ExternalPool *ext) {
    // One batch == one pass of the hot path.  Everything up to the single stream synchronisation is enqueued without
    // waiting: sort -> fastLiftKernel (one lane per interval) -> liftoverKernel over the complex list (one warp per
    // interval) -> scan -> gather -> read-back of the counters.  Buffers come from the context's cache, so a warm batch
    // allocates nothing.  Only a batch with intervals that overflowed their scratch or the record pool (counted by the
    // kernels themselves) enters the retry ladder, which synchronises per rung.
    PhaseTimer pt;
    const auto &G = _file->genomes();
    if (src < 0 || tgt < 0 || src >= (int)G.size() || tgt >= (int)G.size()) throw HalError("genome index out of range");
    if (n >= 0xffffffffull) throw HalError("batch too large (>= 2^32 intervals); split it");
    if ((flags & (HALGPU_NO_DUPES | HALGPU_COLUMN_LIFTOVER)) != 0) coal = -1; // paralogs are only sought with dupes on (halSegmentMapper.cpp:619)
    const Plan &pl = plan(src, tgt, coal);
    const bool coalPath = pl.coal >= 0;
    const GenomeInfo &S = G[src];
    // liftover/impl/halBlockLiftover.cpp:24-30; BlockMapper::map (halBlockMapper.cpp:76-83) seeds from the BOTTOM array when the
    // source genome is the MRCA: HALGPU_SEED_BOTTOM
    const bool srcIsTop = S.numTop > 0 && !((flags & HALGPU_SEED_BOTTOM) != 0 && S.numBottom > 0);
    if (!srcIsTop && S.numBottom == 0) throw HalError("source genome " + S.name + " has no segments");
    const bool raw = (flags & HALGPU_RAW_FRAGMENTS) != 0;
    const bool wantPsl = (flags & HALGPU_PSL) != 0;
    const bool columnMerge = (flags & HALGPU_COLUMN_LIFTOVER) != 0;
    if (wig && coalPath) throw HalError("the wiggle liftover has no coalescence limit (the reference's halWiggleLiftover has none either)");
    if (raw && (wig || wantPsl || columnMerge)) throw HalError("HALGPU_RAW_FRAGMENTS cannot be combined with PSL counts or ColumnLiftover mode");
    out = LiftOutput();
    out.n = n;
    const int launches0 = (int)rt::g_launches;
    // the one-lane-per-interval kernel computes BlockLiftover lines only (no PSL base counts, no raw fragments, ...)
    const bool fast = !wig && !raw && !coalPath && !wantPsl && !columnMerge && !(flags & HALGPU_NO_FAST) && n > 0 && pl.fastOk &&
                      pl.path.size() <= (size_t)HG_FAST_MAX_PATH && std::getenv("HALGPU_NO_FAST") == nullptr;

    Lease L(_cache);
    try {
        enum { C_POOL = 0, C_COLLECT = 1, C_TILE = 2, C_COMPLEX = 3, C_FAIL = 4 /* .. 8, indexed by ST_* */, C_TOTAL = 9, C_TILE_CHUNK = 10 /* .. 13 */, C_WORDS = 16 };
        unsigned long long *outLoc = L.as<unsigned long long>(n + 2);
        uint32_t *status = L.as<uint32_t>(n + 1);
        unsigned long long *ctr = L.as<unsigned long long>(C_WORDS);
        rt::dmemset(status, 0, (n + 1) * sizeof(uint32_t), _stream);
        rt::dmemset(ctr, 0, C_WORDS * sizeof(unsigned long long), _stream);
        // fastLiftKernel leaves a finished interval's line in pool slot <interval id>; outLoc is preset to say so
        if (fast) rt::dmemset(outLoc + n, 0, 2 * sizeof(unsigned long long), _stream);
        else rt::dmemset(outLoc, 0, (n + 2) * sizeof(unsigned long long), _stream);
        pt.mark("alloc0");

        // visit the batch in source order so that neighbouring lanes / warps walk neighbouring records (coalescing, L2 reuse).
        // Only the upper bits of the start decide the order: ~32 source segments per sort bucket are as good as an exact order.
        const unsigned long long *sortedGs = nullptr, *sortedVal = nullptr, *sortedKey = nullptr;
        const bool sorting = !(flags & HALGPU_NO_SORT) && n > 1;
        // HALGPU_PACKED_SORT=1: one 64-bit word per interval (start << 32 | id) instead of a (key, value) pair.  Measured (profiles/
        // r02_bench_n1_f.json): CUB's onesweep pass takes ~100 us for 10 M elements either way, while the lane kernel then has to
        // fetch every interval's end with a scattered read (+50 us): off by default.
        const bool packed = sorting && S.length < 0xffffffffll && std::getenv("HALGPU_PACKED_SORT") != nullptr;
        // HALGPU_SLICES=k (measurement switch, default 1): sort the batch in k slices on a second stream so that the lane kernel of
        // slice c runs while slice c+1 is being sorted.  Measured on 10 M intervals (profiles/r02_bench_n1_slices*.json): one
        // slice 0.776 ms per step, two 0.862 ms, four 1.052 ms -- the two kernels compete for the same L2 / DRAM queues and the
        // shorter launches lose more at their tails than the overlap wins, so the product path keeps one slice.
        int nSlices = 1;
        if (const char *fs = std::getenv("HALGPU_SLICES")) {
            if (fast && sorting && !packed) nSlices = std::max(1, std::min(4, std::atoi(fs)));
        }
        if ((size_t)nSlices > n) nSlices = 1;
        size_t sliceLo[5];
        for (int c = 0; c <= nSlices; ++c) sliceLo[c] = n * (size_t)c / (size_t)nSlices;
        // (A single-pass bucket sort -- count into 2^16 position buckets, scan, scatter with one atomic per interval -- was
        // measured against CUB's two onesweep passes on 10 M intervals: 467 us (count 89, scan 112, scatter 266: scattered
        // 8-byte stores) against 285 us, gpurun_out/launches_k.csv; the radix sort stays.)
        if (sorting || fast) {
            const rt::Stream ss = nSlices > 1 ? _aux : _stream;
            if (nSlices > 1) { _sliceEv[4]->record(_stream); _sliceEv[4]->wait(_aux); } // behind the memsets above
            IotaParams ip;
            std::memset(&ip, 0, sizeof(ip));
            uint64_t *keysIn = nullptr, *valsIn = nullptr;
            if (sorting) { keysIn = L.as<uint64_t>(n); if (!packed) valsIn = L.as<uint64_t>(n); }
            ip.vals = valsIn; ip.keys = packed ? nullptr : keysIn; ip.packed = packed ? keysIn : nullptr;
            ip.gs = dGs; ip.ge = dGe; ip.directLoc = fast ? outLoc : nullptr; ip.n = (int64_t)n;
            rt::launch(iotaKeysKernel, gridFor((int64_t)n, 256, _sms), 256, 0, ss, ip);
            if (sorting) {
                uint64_t *keysOut = L.as<uint64_t>(n), *valsOut = packed ? nullptr : L.as<uint64_t>(n);
                int endBit = 1;
                while (endBit < 64 && (S.length >> endBit) != 0) ++endBit;
                const int coarse = (srcIsTop ? _g[src].topShift : _g[src].botShift) + 5;
                int bits = ((endBit - coarse) / 8) * 8;
                if (const char *sb = std::getenv("HALGPU_SORT_BITS")) bits = std::atoi(sb); // measurement switch
                if (bits < 8) bits = std::min(8, endBit);
                if (bits > endBit) bits = endBit;
                const int beginBit = std::max(0, endBit - bits);
                size_t tmpBytes = 0;
                if (packed) {
                    rt::sortKeysU64Tmp(nullptr, tmpBytes, keysIn, keysOut, n, 32 + beginBit, 32 + endBit, ss);
                    void *tmp = L.take(tmpBytes);
                    rt::sortKeysU64Tmp(tmp, tmpBytes, keysIn, keysOut, n, 32 + beginBit, 32 + endBit, ss);
                    sortedKey = reinterpret_cast<const unsigned long long *>(keysOut);
                } else {
                    const size_t widest = sliceLo[1] - sliceLo[0] + 1;
                    rt::sortPairsU64U64Tmp(nullptr, tmpBytes, keysIn, keysOut, valsIn, valsOut, widest, beginBit, endBit, ss);
                    void *tmp = L.take(tmpBytes);
                    for (int c = 0; c < nSlices; ++c) { // (the slices' sorts run one after the other: one temporary buffer)
                        const size_t lo = sliceLo[c], cnt = sliceLo[c + 1] - lo;
                        size_t tb = tmpBytes;
                        rt::sortPairsU64U64Tmp(tmp, tb, keysIn + lo, keysOut + lo, valsIn + lo, valsOut + lo, cnt, beginBit, endBit, ss);
                        if (nSlices > 1) _sliceEv[c]->record(ss);
                    }
                    sortedGs = reinterpret_cast<const unsigned long long *>(keysOut);
                    sortedVal = reinterpret_cast<const unsigned long long *>(valsOut);
                }
            }
        }
        if (pt.on) rt::sync(_stream);
        pt.mark("sort");

        // record pool: slots [0, n) are the direct slots of fastLiftKernel, the walk allocates behind them.  Sized from the
        // lines-per-interval ratio the last batches on this path produced, so a steady stream of batches does not re-walk
        // intervals that found the pool full.
        const uint64_t direct = fast ? (uint64_t)n : 0;
        uint64_t poolCap = wig ? 1 : direct + (uint64_t)((double)n * std::max(1.25, pl.linesPerInterval * 1.125)) + 4096; // (the wiggle mode emits no records)
        const bool poolExternal = ext != nullptr && ext->buf != nullptr && ext->bytes >= poolCap * sizeof(halgpu_lift_rec) && !wig;
        halgpu_lift_rec *pool = poolExternal ? static_cast<halgpu_lift_rec *>(ext->buf) : L.as<halgpu_lift_rec>(poolCap);
        if (direct) rt::h2d(ctr + C_POOL, &direct, sizeof(direct), _stream); // (pageable 8 bytes: copied before the call returns)
        uint32_t *pslPool = nullptr;
        if (wantPsl) {
            pslPool = L.as<uint32_t>(poolCap * 4);
            rt::dmemset(pslPool, 0, poolCap * 16, _stream);
        }

        LiftParams P;
        std::memset(&P, 0, sizeof(P));
        P.steps = pl.dSteps; P.P = (int32_t)pl.path.size(); P.dupes = (flags & HALGPU_NO_DUPES) ? 0 : 1;
        P.columnMerge = columnMerge ? 1 : 0;
        P.upCanonicalOnly = (P.columnMerge && !P.dupes) ? 1 : 0;
        P.srcIsTop = srcIsTop ? 1 : 0;
        P.srcShift = srcIsTop ? _g[src].topShift : _g[src].botShift;
        P.srcN = srcIsTop ? S.numTop : S.numBottom;
        P.srcLen = S.length;
        P.srcBucket = srcIsTop ? _g[src].topBucket : _g[src].botBucket;
        P.srcNumBuckets = srcIsTop ? _g[src].topBuckets : _g[src].botBuckets;
        P.tgtSeqStart = _g[tgt].seqStart; P.tgtNumSeq = (int32_t)G[tgt].sequences.size();
        P.gs = dGs; P.ge = dGe; P.strand = dStrand;
        P.outLoc = outLoc; P.status = status; P.failCount = ctr + C_FAIL;
        P.pool = pool; P.poolCursor = ctr + C_POOL; P.poolCap = poolCap;
        P.pslPool = pslPool;
        if (wantPsl) { ensureDna(src); ensureDna(tgt); }
        P.srcDna = _g[src].dna; P.tgtDna = _g[tgt].dna;
        if (wig) { P.wigKeys = wig->keys; P.wigValOff = wig->valOff; P.wigVals = wig->vals; }

        // HALGPU_SEED_TILE=1 (measurement switch): TMA-staged seed tiles, see liftover_kernel.cuh and DESIGN.md
        const char *tileEnv = std::getenv("HALGPU_SEED_TILE");
        const bool seedTile = !wig && !raw && !coalPath && srcIsTop && tileEnv != nullptr && tileEnv[0] == '1';
        // HALGPU_FUSE=1 (measurement switch): the fused walk (whole collinear runs per fragment, liftover_kernel.cuh) runs the first
        // pass over plain BED batches; what it flags ST_REDO_EXACT, and every retry rung, is walked piece by piece by the plain
        // instantiation.  Measured on the divergent C2 (gpurun_out/bench_j_default.json): 87 -> 138 ms per 9.94 M intervals although only
        // ~1 % of the intervals are handed back (and the scratch-overflow retries all but vanish): the plain walk moves the 32 equally
        // cut pieces of a synthetic interval in lock step, one hop per iteration, while a fused fragment is cut at every run end into a
        // chain of remainders that are pushed, popped and re-located (bucket search) one after the other with few lanes busy -- off by
        // default; it is the better walk only where segment boundaries do not line up across genomes.
        const char *fuseEnv = std::getenv("HALGPU_FUSE");
        const bool fuse = !wig && !raw && !coalPath && !wantPsl && !columnMerge && !seedTile && !(flags & HALGPU_NO_FAST) && pl.fastOk &&
                          fuseEnv != nullptr && fuseEnv[0] == '1';
        void (*const firstKernel)(const LiftParams) = fuse ? liftoverKernel<LIFT_BED | LIFT_FUSE> : nullptr;
        void (*const mapKernel)(const LiftParams) =
            seedTile ? liftoverKernel<LIFT_BED | LIFT_TILE> : wig ? liftoverKernel<LIFT_WIG>
                : (raw ? (coalPath ? liftoverKernel<LIFT_RAW_COAL> : liftoverKernel<LIFT_RAW>) : (coalPath ? liftoverKernel<LIFT_COAL> : liftoverKernel<LIFT_BED>));
        const unsigned block = 128, warpsPerBlock = block / 32;
        uint32_t *complexList = nullptr;
        // rung 1: the whole batch; the walk keeps its lists in shared memory
        _ev[0]->record(_stream);
        if (fast) {
            complexList = L.as<uint32_t>(n + 1);
            FastParams F;
            std::memset(&F, 0, sizeof(F));
            F.steps = pl.dSteps; F.P = P.P; F.srcIsTop = P.srcIsTop; F.srcLen = S.length;
            F.tgtSeqStart = P.tgtSeqStart; F.tgtNumSeq = P.tgtNumSeq;
            F.n = (int64_t)n; F.gs = dGs; F.ge = dGe; F.strand = dStrand;
            F.sortedGs = sortedGs; F.sortedVal = sortedVal; F.sortedKey = sortedKey;
            F.tileCursor = ctr + C_TILE; F.pool = pool;
            // tiles of 32 work items are handed out 4 at a time from an atomic cursor, so the resident warps sweep the sorted batch
            // together.  Measured on 10 M intervals (profiles/r02_bench_n1_tilegrab*.json): 1 tile per atomicAdd 0.354 ms, 4 tiles
            // 0.302 ms, 16 tiles 0.411 ms, a fixed stride per warp (since removed) 0.426 ms.
            int tileGrab = 4;
            if (const char *tg = std::getenv("HALGPU_TILE_GRAB")) tileGrab = std::max(1, std::min(64, std::atoi(tg))); // measurement switch
            F.tileGrab = tileGrab;
            F.complexList = complexList; F.complexCount = ctr + C_COMPLEX;
            for (int c = 0; c < nSlices; ++c) {
                const size_t lo = sliceLo[c], cnt = sliceLo[c + 1] - lo;
                if (cnt == 0) continue;
                if (nSlices > 1) _sliceEv[c]->wait(_stream); // this slice is sorted
                FastParams Fc = F;
                Fc.n = (int64_t)cnt;
                if (sortedGs) { Fc.sortedGs = sortedGs + lo; Fc.sortedVal = sortedVal + lo; }
                if (sortedKey) Fc.sortedKey = sortedKey + lo;
                Fc.tileCursor = nSlices > 1 ? ctr + C_TILE_CHUNK + c : ctr + C_TILE;
                const int64_t tiles = ((int64_t)cnt + 31) / 32;
                const unsigned fgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((tiles + 7) / 8, (int64_t)_sms * 8));
                rt::launch(fastLiftKernel, fgrid, 256, 0, _stream, Fc);
            }
            _ev[1]->record(_stream);
        }
        {
            P.listCap = 64; P.frameCap = 32; P.gscratch = nullptr; P.gscratchPerWarp = 0;
            P.n = (int64_t)n;
            if (fast) { P.work = complexList; P.nDev = ctr + C_COMPLEX; }
            else P.work64 = sortedKey ? sortedKey : sortedVal; // (the interval id sits in the low 32 bits of either)
            P.seedTile = seedTile ? 1 : 0;
            const size_t smem = ((size_t)liftScratchBytes(P.listCap, P.frameCap) + (P.seedTile ? (size_t)seedTileBytes() : 0)) * warpsPerBlock;
            rt::allowSmem(mapKernel, smem);
            if (firstKernel) rt::allowSmem(firstKernel, smem);
            rt::launch(firstKernel ? firstKernel : mapKernel, gridFor((int64_t)n, warpsPerBlock, _sms * 2), block, smem, _stream, P);
            P.work = nullptr; P.work64 = nullptr; P.nDev = nullptr;
        }
        _ev[2]->record(_stream);
        if (pt.on) rt::sync(_stream);
        pt.mark("kernel");

        // CSR assembly in input order, enqueued right behind the kernels; redone only if the retry ladder had to run
        uint64_t *csr = nullptr;
        halgpu_lift_rec *recs = nullptr;
        uint32_t *psl = nullptr;
        uint64_t recCap = 0;
        uint64_t *blockSums = nullptr;
        auto assemble = [&]() {
            if (wig) return;
            if (csr == nullptr) {
                csr = L.as<uint64_t>(n + 2);
                blockSums = L.as<uint64_t>(n / (HG_SCAN_BLOCK * HG_SCAN_ITEMS) + 2);
            }
            if (recs == nullptr || recCap < P.poolCap) {
                if (recs) L.giveNow(recs);
                if (psl) L.giveNow(psl);
                recCap = P.poolCap;
                recs = L.as<halgpu_lift_rec>(recCap);
                psl = wantPsl ? L.as<uint32_t>(recCap * 4) : nullptr;
            }
            CsrScanParams sp;
            sp.outLoc = outLoc; sp.csr = csr; sp.blockSums = blockSums; sp.n = (int64_t)n;
            sp.skipIfZero = fast ? ctr + C_COMPLEX : nullptr; // all intervals finished by fastLiftKernel: offsets[i] = i
            const unsigned sgrid = gridFor(((int64_t)n + HG_SCAN_BLOCK * HG_SCAN_ITEMS - 1) / (HG_SCAN_BLOCK * HG_SCAN_ITEMS), 1, _sms);
            if (fast) rt::launch(csrIdentityKernel, gridFor((int64_t)n + 1, 256, _sms), 256, 0, _stream, sp);
            rt::launch(csrScanKernel<0>, sgrid, HG_SCAN_BLOCK, 0, _stream, sp);
            rt::launch(csrScanSumsKernel, 1, HG_SCAN_BLOCK, 0, _stream, sp);
            rt::launch(csrScanKernel<2>, sgrid, HG_SCAN_BLOCK, 0, _stream, sp);
            GatherParams gp;
            gp.pslPool = P.pslPool; gp.psl = psl;
            gp.outLoc = outLoc; gp.csr = csr; gp.pool = P.pool; gp.recs = recs; gp.n = (int64_t)n;
            gp.skipIfZero = fast ? ctr + C_COMPLEX : nullptr;
            rt::launch(gatherKernel, gridFor((int64_t)n, 256, _sms), 256, 0, _stream, gp);
            rt::d2d(ctr + C_TOTAL, csr + n, sizeof(uint64_t), _stream);
            if (offsetBase != 0) { // chunk-local offsets -> offsets of the whole batch (halgpu_liftover lifts chunk by chunk)
                AddBaseParams ab;
                ab.v = csr; ab.n = (int64_t)n + 1; ab.base = offsetBase;
                rt::launch(addBaseKernel, gridFor((int64_t)n + 1, 256, _sms), 256, 0, _stream, ab);
            }
        };
        auto readBack = [&]() { // the batch's one synchronisation
            rt::d2h(_hostCtr, ctr, C_WORDS * sizeof(unsigned long long), _stream);
            rt::sync(_stream);
        };
        assemble();
        readBack();
        out.kernelMs = rt::Event::elapsedMs(*_ev[0], *_ev[2]);
        if (fast) { out.fastMs = rt::Event::elapsedMs(*_ev[0], *_ev[1]); out.nComplex = (size_t)_hostCtr[C_COMPLEX]; }
        pt.mark("assemble");

        // retry ladder: pool growth, then larger per-warp scratch in global memory
        uint32_t *list = nullptr;
        auto collect = [&](uint32_t want) -> uint64_t {
            if (list == nullptr) list = L.as<uint32_t>(n + 1);
            rt::dmemset(ctr + C_COLLECT, 0, sizeof(unsigned long long), _stream);
            CollectParams cp;
            cp.status = status; cp.list = list; cp.count = ctr + C_COLLECT; cp.subset = nullptr; cp.n = (int64_t)n; cp.want = want;
            rt::launch(collectKernel, gridFor((int64_t)n, 256, _sms), 256, 0, _stream, cp);
            unsigned long long c = 0;
            rt::d2h(&c, ctr + C_COLLECT, sizeof(c), _stream);
            rt::sync(_stream);
            return c;
        };
        int listCap = 64, frameCap = 32;
        bool redo = false;
        for (int round = 0; round < 64; ++round) {
            const unsigned long long *fails = _hostCtr + C_FAIL;
            if (std::getenv("HALGPU_DEBUG")) fprintf(stderr, "[halgpu] round %d fails %llu %llu %llu pool %llu complex %llu total %llu\n", round, fails[1], fails[2], fails[3], _hostCtr[C_POOL], _hostCtr[C_COMPLEX], _hostCtr[C_TOTAL]);
            if (fails[ST_BAD_INPUT] > 0) {
                throw HalError(std::to_string(fails[ST_BAD_INPUT]) + " interval(s) lie outside genome " + S.name + " (length " + std::to_string(S.length) + ")");
            }
            if (fails[ST_POOL_FULL] == 0 && fails[ST_SCRATCH_OVERFLOW] == 0 && fails[ST_REDO_EXACT] == 0) break;
            redo = true;
            rt::Event r0, r1;
            // (a) intervals the fused walk handed back, and intervals that found the output pool full (grow it: old records stay
            //     valid): redo them with the plain walk
            const uint32_t want = fails[ST_REDO_EXACT] ? (uint32_t)ST_REDO_EXACT : (uint32_t)ST_POOL_FULL;
            const uint64_t nFull = (fails[ST_REDO_EXACT] || fails[ST_POOL_FULL]) ? collect(want) : 0;
            if (nFull > 0) {
                if (want == ST_REDO_EXACT) out.nRedo += nFull;
                const uint64_t used = _hostCtr[C_POOL];
                const uint64_t newCap = want == ST_REDO_EXACT ? poolCap : std::max<uint64_t>(poolCap * 2, used * 2);
                if (newCap != poolCap) {
                halgpu_lift_rec *np = L.as<halgpu_lift_rec>(newCap);
                rt::d2d(np, pool, poolCap * sizeof(halgpu_lift_rec), _stream);
                if (wantPsl) { // counters of the records already written move along; the new tail starts at zero
                    uint32_t *nq = L.as<uint32_t>(newCap * 4);
                    rt::dmemset(nq, 0, newCap * 16, _stream);
                    rt::d2d(nq, pslPool, poolCap * 16, _stream);
                    rt::sync(_stream);
                    L.giveNow(pslPool);
                    pslPool = nq;
                    P.pslPool = pslPool;
                }
                rt::sync(_stream);
                L.giveNow(pool);
                pool = np;
                poolCap = newCap;
                P.pool = pool; P.poolCap = poolCap;
                }
                uint32_t *ids = L.as<uint32_t>(nFull);
                rt::d2d(ids, list, nFull * sizeof(uint32_t), _stream);
                rt::dmemset(ctr + C_FAIL + want, 0, sizeof(unsigned long long), _stream); // all of them run again
                P.n = (int64_t)nFull; P.work = ids;
                const bool inSmem = listCap == 64;
                const uint64_t per = liftScratchBytes(listCap, frameCap);
                const int64_t warps = std::min<int64_t>((int64_t)nFull, (int64_t)_sms * 8);
                void *scratch = nullptr;
                P.listCap = listCap; P.frameCap = frameCap;
                if (inSmem) { P.gscratch = nullptr; P.gscratchPerWarp = 0; }
                else { scratch = L.take(per * (uint64_t)warps); P.gscratch = static_cast<uint8_t *>(scratch); P.gscratchPerWarp = per; }
                const unsigned grid = inSmem ? gridFor((int64_t)nFull, warpsPerBlock, _sms * 2) : (unsigned)((warps + warpsPerBlock - 1) / warpsPerBlock);
                r0.record(_stream);
                rt::launch(mapKernel, grid, block, (inSmem ? (size_t)per * warpsPerBlock : 0) + (seedTile ? (size_t)seedTileBytes() * warpsPerBlock : 0), _stream, P);
                r1.record(_stream);
                readBack();
                out.kernelMs += rt::Event::elapsedMs(r0, r1);
                L.giveNow(ids);
                if (scratch) L.giveNow(scratch);
                continue;
            }
            // (b) intervals whose fragment lists outgrew the scratch: next rung
            const uint64_t nOver = collect(ST_SCRATCH_OVERFLOW);
            if (nOver == 0) break;
            if (out.nRetry == 0) out.nRetry = nOver;
            listCap *= (listCap == 64 ? 64 : 16); // 64 -> 4096 -> 65536 -> 1M
            frameCap = listCap / 4;
            if (listCap > (1 << 24)) throw HalError("an interval maps to more than 16M fragments; not supported");
            // the per-warp refinement / merge of one interval sorts and cuts its fragments with quadratic work (fine for the tens of
            // fragments of a BED interval, seconds at 64 K): beyond that the caller has to lift pieces, or take the fragments
            // themselves (HALGPU_RAW_FRAGMENTS) and refine them with an n log n pass as halSynteny's host layer does
            if (!raw && !wig && listCap > (1 << 16)) {
                throw HalError("an interval maps to more than 65536 fragments; lift it in pieces (or use HALGPU_RAW_FRAGMENTS and refine on the caller's side, as halSynteny does)");
            }
            const uint64_t per = liftScratchBytes(listCap, frameCap);
            int64_t warps = std::min<int64_t>((int64_t)nOver, (int64_t)_sms * 8);
            const uint64_t budget = 8ull << 30; // scratch budget
            if ((uint64_t)warps * per > budget) warps = std::max<int64_t>(1, (int64_t)(budget / per));
            void *scratch = L.take(per * (uint64_t)warps);
            uint32_t *ids = L.as<uint32_t>(nOver);
            rt::d2d(ids, list, nOver * sizeof(uint32_t), _stream);
            rt::dmemset(ctr + C_FAIL + ST_SCRATCH_OVERFLOW, 0, sizeof(unsigned long long), _stream); // all of them run again
            P.n = (int64_t)nOver; P.work = ids;
            P.listCap = listCap; P.frameCap = frameCap; P.gscratch = static_cast<uint8_t *>(scratch); P.gscratchPerWarp = per;
            r0.record(_stream);
            rt::launch(mapKernel, (unsigned)((warps + warpsPerBlock - 1) / warpsPerBlock), block, seedTile ? (size_t)seedTileBytes() * warpsPerBlock : 0, _stream, P);
            r1.record(_stream);
            readBack();
            out.kernelMs += rt::Event::elapsedMs(r0, r1);
            L.giveNow(ids);
            L.giveNow(scratch);
        }
        if (redo) {
            assemble();
            readBack();
        }
        pt.mark("retries");
        if (std::getenv("HALGPU_DEBUG")) fprintf(stderr, "[halgpu] batch of %zu: complex %llu, re-walked exactly %zu, scratch retries %zu, fused walk %d\n", n, _hostCtr[C_COMPLEX], out.nRedo, out.nRetry, (int)fuse);
        if (wig) { // values went straight into wig->keys; there is no record list to assemble
            out.launches = (int)rt::g_launches - launches0;
            return;
        }
        if (!raw && !columnMerge && n > 0) { // what the pool has to hold next time (the cursor counts every request, granted or not)
            const uint64_t directLines = fast ? (uint64_t)n - (uint64_t)_hostCtr[C_COMPLEX] : 0;
            const double lpi = (double)(_hostCtr[C_TOTAL] - directLines) / (double)n;
            if (lpi > pl.linesPerInterval) pl.linesPerInterval = lpi;
        }
        out.offsets = static_cast<uint64_t *>(L.detach(csr));
        if (fast && _hostCtr[C_COMPLEX] == 0) recs = pool; // every line sits in its direct slot: the pool is the result
        out.recs = static_cast<halgpu_lift_rec *>(L.detach(recs));
        out.recsExternal = poolExternal && recs == static_cast<halgpu_lift_rec *>(ext->buf);
        out.psl = wantPsl ? static_cast<uint32_t *>(L.detach(psl)) : nullptr;
        out.nRec = (size_t)_hostCtr[C_TOTAL];
        out.launches = (int)rt::g_launches - launches0;
    } catch (...) {
        try { rt::sync(_aux); rt::sync(_stream); } catch (...) {} // nothing of this batch may still run when its buffers return to the cache
        throw;
    }
}

void Context::wiggle(int src, int tgt, uint32_t flags, size_t nRuns, const int64_t *first, const int64_t *last, const int64_t *valOff,
                     const double *vals, size_t nVals, size_t nPre, const int64_t *prePos, const double *preVal, WigOutput &out) {
    const auto &G = _file->genomes();
    if (src < 0 || tgt < 0 || src >= (int)G.size() || tgt >= (int)G.size()) throw HalError("genome index out of range");
    const int64_t tgtLen = G[tgt].length;
    for (size_t i = 0; i < nRuns; ++i) { // every value index the kernel will form must exist
        if (last[i] < first[i]) throw HalError("wiggle run " + std::to_string(i) + " is empty or reversed");
        const bool ok = valOff[i] >= 0 ? (uint64_t)valOff[i] + (uint64_t)(last[i] - first[i]) < nVals : (uint64_t)(~valOff[i]) < nVals;
        if (!ok) throw HalError("wiggle run " + std::to_string(i) + " refers to values outside the value array");
    }
    for (size_t i = 0; i < nPre; ++i) {
        if (prePos[i] < 0 || prePos[i] >= tgtLen) throw HalError("preloaded wiggle position outside genome " + G[tgt].name);
    }
    out = WigOutput();
    DevBuf::current() = _stream;
    const int launches0 = (int)rt::g_launches;
    DevBuf keys((size_t)std::max<int64_t>(tgtLen, 1) * 8);
    {
        WigFillParams fp;
        fp.keys = keys.as<unsigned long long>(); fp.n = tgtLen;
        rt::launch(wigFillKernel, gridFor(tgtLen, 256, _sms), 256, 0, _stream, fp);
    }
    if (nPre > 0) {
        DevBuf dp(nPre * 8), dv(nPre * 8);
        rt::h2d(dp.p, prePos, nPre * 8, _stream);
        rt::h2d(dv.p, preVal, nPre * 8, _stream);
        WigPreloadParams pp;
        pp.keys = keys.as<unsigned long long>(); pp.pos = dp.as<int64_t>(); pp.val = dv.as<double>(); pp.n = (int64_t)nPre; pp.genomeLen = tgtLen;
        rt::launch(wigPreloadKernel, gridFor((int64_t)nPre, 256, _sms), 256, 0, _stream, pp);
        rt::sync(_stream);
    }
    if (nRuns > 0) {
        DevBuf dF(nRuns * 8), dL(nRuns * 8), dO(nRuns * 8), dV(std::max<size_t>(nVals, 1) * 8);
        rt::h2d(dF.p, first, nRuns * 8, _stream);
        rt::h2d(dL.p, last, nRuns * 8, _stream);
        rt::h2d(dO.p, valOff, nRuns * 8, _stream);
        rt::h2d(dV.p, vals, nVals * 8, _stream);
        WigScatter ws;
        ws.keys = keys.as<unsigned long long>(); ws.valOff = dO.as<int64_t>(); ws.vals = dV.as<double>();
        LiftOutput lo;
        liftover(src, tgt, flags & HALGPU_NO_DUPES, nRuns, dF.as<int64_t>(), dL.as<int64_t>(), nullptr, lo, 0, &ws);
        out.kernelMs = lo.kernelMs;
        out.nRetry = lo.nRetry;
        DevBuf::current() = _stream;
    }
    // read-out: ascending positions of the bases that hold a value, then their values
    DevBuf flag((size_t)std::max<int64_t>(tgtLen, 1)), pos((size_t)std::max<int64_t>(tgtLen, 1) * 8), cnt(sizeof(unsigned long long));
    WigFlagParams fl;
    fl.keys = keys.as<unsigned long long>(); fl.flag = flag.as<uint8_t>(); fl.n = tgtLen;
    rt::launch(wigFlagKernel, gridFor(tgtLen, 256, _sms), 256, 0, _stream, fl);
    rt::selectFlagged(flag.as<uint8_t>(), pos.as<int64_t>(), cnt.as<unsigned long long>(), (size_t)tgtLen, _stream);
    unsigned long long nSet = 0;
    rt::d2h(&nSet, cnt.p, sizeof(nSet), _stream);
    rt::sync(_stream);
    DevBuf val(std::max<size_t>((size_t)nSet, 1) * 8);
    WigGatherParams gp;
    gp.keys = keys.as<unsigned long long>(); gp.pos = pos.as<int64_t>(); gp.val = val.as<double>(); gp.n = (int64_t)nSet;
    rt::launch(wigGatherKernel, gridFor((int64_t)nSet, 256, _sms), 256, 0, _stream, gp);
    out.pos = static_cast<int64_t *>(rt::hostAlloc(std::max<size_t>((size_t)nSet, 1) * 8));
    out.val = static_cast<double *>(rt::hostAlloc(std::max<size_t>((size_t)nSet, 1) * 8));
    rt::d2h(out.pos, pos.p, (size_t)nSet * 8, _stream);
    rt::d2h(out.val, val.p, (size_t)nSet * 8, _stream);
    rt::sync(_stream);
    out.n = (size_t)nSet;
    out.launches = (int)rt::g_launches - launches0;
}

} // namespace halgpu
